"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU and exports every symbol that
include/xvector_b200.h declares; compute entry points fail LOUDLY (no CPU fallback) when there is no sm_100 device."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "xvector_b200.h")).read()
    return sorted(set(re.findall(r"XV_API\s+[\w\s\*]+?\b(xv_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tf_kaldi_speaker_b200 import _lib as L
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 25, names
    for n in names:
        assert hasattr(lib, n), "libxvector_b200.so does not export %s" % n
    assert lib.xv_version() >= 100


def test_struct_layouts_match_header():
    """ctypes mirrors of xv_operand / xv_head_args / xv_gemm_args (sizes follow the C layout rules of the header)."""
    from tf_kaldi_speaker_b200 import _lib as L
    assert C.sizeof(L.Operand) == 48
    assert C.sizeof(L.HeadArgs) == 24 + 8 * 8 + 8 + 8       # 6 scalars, 8 pointers, inv_batch (+pad), gnorm
    assert L.GemmArgs.head.offset % 8 == 0
    assert L.GemmArgs.out.offset == 2 * 48 + 8 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from tf_kaldi_speaker_b200 import _lib as L
    from tf_kaldi_speaker_b200.runtime import Engine
    with pytest.raises(L.XvError):
        Engine()
    lib = L.load()
    info = (C.c_int32 * 3)()
    rc = lib.xv_device_info(info)
    assert rc == L.XV_ERR_CUDA
    assert len(lib.xv_last_error()) > 0
    # argument validation happens before any CUDA call and reports through xv_last_error()
    args = L.GemmArgs()
    assert lib.xv_gemm_bf16(C.byref(args), None) == L.XV_ERR_INVALID
    assert b"positive" in lib.xv_last_error()
