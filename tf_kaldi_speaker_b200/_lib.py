"""ctypes binding of libxvector_b200.so (the C ABI declared in include/xvector_b200.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point reports an
error, this module raises.  PyTorch is used only as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XV_LIB_PATH: load a differently-built library (compile-time tuning sweeps, tools/layers_bench.py); default = in-tree build
LIB_PATH = os.environ.get("XV_LIB_PATH") or os.path.join(_HERE, "libxvector_b200.so")

XV_OK, XV_ERR_INVALID, XV_ERR_UNSUPPORTED, XV_ERR_CUDA = 0, -1, -2, -3
EPI_BF16, EPI_F32, EPI_HEAD_FWD, EPI_HEAD_BWD = 0, 1, 2, 3
HEAD_SOFTMAX, HEAD_ASOFTMAX, HEAD_AM, HEAD_AAM = 0, 1, 2, 3
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_PRELU, ACT_TANH = 0, 1, 2, 3, 4
OPT_SGD, OPT_MOMENTUM, OPT_NESTEROV, OPT_ADAM = 0, 1, 2, 3


class XvError(RuntimeError):
    pass


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("ld", C.c_int64),
                ("mn_major", C.c_int32), ("div", C.c_int32), ("tap_rows", C.c_int32), ("_pad", C.c_int32)]


class HeadArgs(C.Structure):
    _fields_ = [("type", C.c_int32), ("asoftmax_m", C.c_int32), ("margin", C.c_float), ("cos_m", C.c_float),
                ("sin_m", C.c_float), ("threshold", C.c_float),
                ("sched", C.c_void_p), ("labels", C.c_void_p), ("xnorm", C.c_void_p), ("part_max", C.c_void_p),
                ("part_sum", C.c_void_p), ("target_logit", C.c_void_p), ("logits_out", C.c_void_p),
                ("lse", C.c_void_p), ("inv_batch", C.c_float), ("gnorm", C.c_void_p)]


class BnBwdArgs(C.Structure):
    _fields_ = [("y", C.c_void_p), ("ldy", C.c_int64), ("scale", C.c_void_p), ("shift", C.c_void_p), ("mean", C.c_void_p),
                ("rstd", C.c_void_p), ("neg_slope", C.c_float), ("_pad", C.c_int32)]


class GemmArgs(C.Structure):
    _fields_ = [("a", Operand), ("b", Operand), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("splits", C.c_int32), ("epilogue", C.c_int32), ("seg_len", C.c_int32), ("seg_valid", C.c_int32),
                ("accumulate", C.c_int32), ("out", C.c_void_p), ("ldc", C.c_int64), ("bias", C.c_void_p),
                ("col_sum", C.c_void_p), ("col_sumsq", C.c_void_p), ("head", HeadArgs), ("bn_bwd", BnBwdArgs),
                ("affine_scale", C.c_void_p), ("affine_shift", C.c_void_p), ("affine_neg_slope", C.c_float),
                ("_pad2", C.c_int32)]


_lib = None


def load():
    """Load the shared library once.  Raises XvError (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XvError("libxvector_b200.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                      "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.xv_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().xv_last_error().decode("utf-8", "replace")
        if rc == XV_ERR_UNSUPPORTED:
            raise NotImplementedError(msg)
        if rc == XV_ERR_INVALID:
            raise ValueError(msg)
        raise XvError("xvector_b200 error %d: %s" % (rc, msg))


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def operand(t, mn_major=False, div=0, tap_rows=0, rows=None, cols=None):
    """Describe a 2-D bf16 torch tensor (or a [rows, cols] window of it) as a GEMM operand."""
    assert t.dim() == 2 and t.stride(1) == 1, "operand must be row-major 2-D"
    return Operand(t.data_ptr(), t.shape[0] if rows is None else rows, t.shape[1] if cols is None else cols,
                   t.stride(0), 1 if mn_major else 0, div, tap_rows, 0)


def gemm(a_op, b_op, M, N, K, out, epilogue=EPI_BF16, splits=1, bias=None, col_sum=None, col_sumsq=None,
         seg_len=0, seg_valid=0, head=None, ldc=None, accumulate=False, bn_bwd=None, affine=None):
    args = GemmArgs()
    args.a, args.b = a_op, b_op
    args.M, args.N, args.K = M, N, K
    args.splits, args.epilogue = splits, epilogue
    args.seg_len, args.seg_valid = seg_len, seg_valid
    args.accumulate = 1 if accumulate else 0
    args.out = out.data_ptr()
    args.ldc = out.stride(0) if ldc is None else ldc
    args.bias = 0 if bias is None else bias.data_ptr()
    args.col_sum = 0 if col_sum is None else col_sum.data_ptr()
    args.col_sumsq = 0 if col_sumsq is None else col_sumsq.data_ptr()
    if head is not None:
        args.head = head
    if bn_bwd is not None:       # (y, scale, shift, mean, rstd, neg_slope): fused BN-backward reductions (dgrad epilogue)
        y, scale, shift, mean, rstd, neg_slope = bn_bwd
        args.bn_bwd.y, args.bn_bwd.ldy = y.data_ptr(), y.stride(0)
        args.bn_bwd.scale, args.bn_bwd.shift = scale.data_ptr(), shift.data_ptr()
        args.bn_bwd.mean, args.bn_bwd.rstd = mean.data_ptr(), rstd.data_ptr()
        args.bn_bwd.neg_slope = float(neg_slope)
    if affine is not None:       # (scale, shift, neg_slope): inference-mode BN + activation folded into the epilogue
        args.affine_scale, args.affine_shift = affine[0].data_ptr(), affine[1].data_ptr()
        args.affine_neg_slope = float(affine[2])
    check(load().xv_gemm_bf16(C.byref(args), stream_ptr()))
