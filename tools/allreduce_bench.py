#!/usr/bin/env python
"""Gradient-exchange microbenchmark (run under torchrun): NCCL all-reduce vs xv_dp_allreduce_multimem on the config-2
flat gradient buffer (9.75 M floats = 39 MB) and on the trunk-only prefix (24.3 MB).  CUDA events, max over ranks.
    XV_AR_CFG=0..3 python -m torch.distributed.run --nproc-per-node N ... tools/allreduce_bench.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

from tf_kaldi_speaker_b200 import _lib as L
from tf_kaldi_speaker_b200 import parallel


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / reps * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world = parallel.init_from_env("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    lib = L.load()
    info = (C.c_int32 * 3)()
    L.check(lib.xv_device_info(info))
    sms = int(info[0])
    group = dist.group.WORLD.group_name
    out = {"world": world, "cfg": os.environ.get("XV_AR_CFG", "default")}
    nmax = 9749504 + (1 << 20)
    buf = symm_mem.empty(nmax, dtype=torch.float32, device="cuda")
    hdl = symm_mem.rendezvous(buf, group)
    flags = symm_mem.empty(8192, dtype=torch.int32, device="cuda")
    hf = symm_mem.rendezvous(flags, group)
    flags.zero_()
    torch.cuda.synchronize()
    dist.barrier()
    epoch = torch.zeros(1024, dtype=torch.int32, device="cuda")
    plain = torch.zeros(nmax, dtype=torch.float32, device="cuda")
    for name, n in (("full 39.0 MB", 9749504), ("trunk 24.3 MB", 6063104), ("2 MB", 524288)):
        buf.fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        L.check(lib.xv_dp_allreduce_multimem(C.c_void_p(int(hdl.multicast_ptr)), C.c_void_p(int(hf.buffer_ptrs_dev)), L.ptr(epoch),
                                             rank, world, C.c_int64(0), C.c_int64(n), sms, L.stream_ptr()))
        torch.cuda.synchronize()
        want = float(sum(r + 1 for r in range(world)))
        ok = bool((buf[:n] == want).all().item()) and bool((buf[n:n + 1024] == float(rank + 1)).all().item())
        t_mm = timed(lambda: L.check(lib.xv_dp_allreduce_multimem(C.c_void_p(int(hdl.multicast_ptr)), C.c_void_p(int(hf.buffer_ptrs_dev)),
                                                                  L.ptr(epoch), rank, world, C.c_int64(0), C.c_int64(n), sms, L.stream_ptr())))
        buf.fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        L.check(lib.xv_dp_allreduce_p2p(C.c_void_p(int(hdl.buffer_ptrs_dev)), C.c_void_p(int(hf.buffer_ptrs_dev)), L.ptr(epoch),
                                        rank, world, C.c_int64(0), C.c_int64(n), sms, L.stream_ptr()))
        torch.cuda.synchronize()
        ok = ok and bool((buf[:n] == want).all().item()) and bool((buf[n:n + 1024] == float(rank + 1)).all().item())
        t_p2p = timed(lambda: L.check(lib.xv_dp_allreduce_p2p(C.c_void_p(int(hdl.buffer_ptrs_dev)), C.c_void_p(int(hf.buffer_ptrs_dev)),
                                                              L.ptr(epoch), rank, world, C.c_int64(0), C.c_int64(n), sms, L.stream_ptr())))
        g = plain[:n]
        t_nccl = timed(lambda: dist.all_reduce(g))
        sweep = {}
        if os.environ.get("XV_AR_SWEEP"):
            keep = os.environ.get("XV_AR_CFG")
            for cfg in os.environ["XV_AR_SWEEP"].split(","):
                os.environ["XV_AR_CFG"] = cfg
                buf.fill_(float(rank + 1))
                torch.cuda.synchronize()
                dist.barrier()
                L.check(lib.xv_dp_allreduce_multimem(C.c_void_p(int(hdl.multicast_ptr)), C.c_void_p(int(hf.buffer_ptrs_dev)), L.ptr(epoch),
                                                     rank, world, C.c_int64(0), C.c_int64(n), sms, L.stream_ptr()))
                torch.cuda.synchronize()
                good = bool((buf[:n] == want).all().item()) and bool((buf[n:n + 1024] == float(rank + 1)).all().item())
                t = timed(lambda: L.check(lib.xv_dp_allreduce_multimem(C.c_void_p(int(hdl.multicast_ptr)), C.c_void_p(int(hf.buffer_ptrs_dev)),
                                                                       L.ptr(epoch), rank, world, C.c_int64(0), C.c_int64(n), sms,
                                                                       L.stream_ptr())))
                sweep["cfg%s" % cfg] = {"us": t, "correct": good}
            if keep is None:
                os.environ.pop("XV_AR_CFG", None)
            else:
                os.environ["XV_AR_CFG"] = keep
        out[name] = {"multimem_us": t_mm, "p2p_us": t_p2p, "nccl_us": t_nccl, "correct": ok, "multimem_cfg_sweep": sweep,
                     "multimem_busbw_GBs": 2.0 * (world - 1) / world * n * 4 / t_mm / 1e3,
                     "nccl_busbw_GBs": 2.0 * (world - 1) / world * n * 4 / t_nccl / 1e3}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
