"""In-tree build of libxvector_b200.so (sm_100a only): parallel ``nvcc -c`` per translation unit with a
content-hash cache under ``build/``, then one link step.  The .so is git-ignored but travels to the GPU box."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# XV_LIB_VARIANT=name: build libxvector_b200.<name>.so in its own object directory (tuning sweeps with XV_EXTRA_CFLAGS)
_VARIANT = os.environ.get("XV_LIB_VARIANT", "")
BUILD = os.path.join(ROOT, "build", "xvector_b200" + ("." + _VARIANT if _VARIANT else ""))
LIB = os.path.join(HERE, "libxvector_b200%s.so" % ("." + _VARIANT if _VARIANT else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
CFLAGS += os.environ.get("XV_EXTRA_CFLAGS", "").split()       # tuning sweeps, e.g. XV_EXTRA_CFLAGS="-DXV_RU=4"


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".h", ".cuh")):
                h.update(f.encode())
                h.update(open(os.path.join(d, f), "rb").read())
    h.update(" ".join(CFLAGS).encode())
    return h.hexdigest()


def _compile(src, hdig, verbose):
    path = os.path.join(CSRC, src)
    dig = hashlib.sha256(open(path, "rb").read() + hdig.encode()).hexdigest()[:20]
    obj = os.path.join(BUILD, "%s.%s.o" % (src[:-3], dig))
    if os.path.exists(obj):
        return obj, False
    for old in os.listdir(BUILD):
        if old.startswith(src[:-3] + ".") and old.endswith(".o"):
            os.remove(os.path.join(BUILD, old))
    cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(obj + ".tmp", obj)
    return obj, True


def build(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for f in os.listdir(BUILD):
            os.remove(os.path.join(BUILD, f))
    hdig = _headers_digest()
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, hdig, verbose), srcs))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB + ".tmp"] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                               "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
