#!/usr/bin/env python
"""N-GPU NCCL check of SyncBN (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 \
        tools/dist_check_syncbn.py

SURVEY 8e equivalence: an N-GPU data-parallel step with ``sync_bn`` on per-rank batches of B rows must equal the
1-GPU step on the concatenated N*B rows (loss, parameter update, BN moving statistics) up to reduction-order rounding.
Trainer A (every rank, no communication): the global batch on one device.  Trainer B: DataParallel + sync_bn on this
rank's rows.  Trainer C: DataParallel with per-replica BN (the default) -- reported to show that the difference A-C is
real, i.e. that the check can fail.  One-step updates from identical parameters, eager call and CUDA-graph replay, with
trainer A's own eager-vs-replay error as the noise floor (see tools/dist_check_sharded.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
from tf_kaldi_speaker_b200 import parallel
from tf_kaldi_speaker_b200.misc.utils import ParamsPlain
from tf_kaldi_speaker_b200.model.trainer import Trainer
from tf_kaldi_speaker_b200.runtime import set_engine

B, T, C = 32, 100, 503
LOSS = "additive_angular_margin_softmax"


ATTENTION = "--attention" in sys.argv      # multi-head attention pooling with a non-zero penalty term (model/pooling.py:185-189):
                                           # the penalty is divided by the GLOBAL batch, so N ranks must reproduce the 1-GPU value


def run(mode, rank, world, x, y):
    pd = dict(bench.PD)
    pd["sync_bn"] = (mode == "sync")
    if ATTENTION:
        pd.update(pooling_type="self_attention", att_key_input="tdnn4_relu", att_key_num_nodes=[256, 256],
                  att_key_network_type=3, att_value_input="tdnn5_relu", att_value_num_nodes=[], att_value_network_type=0,
                  att_apply_nonlinear=False, att_use_scale=True, att_num_heads=4, att_split_key=True, att_penalty_term=0.5)
    tr = Trainer(ParamsPlain(**pd), "/tmp/xv_syncbn_%s_%d" % (mode, rank))
    tr.build("train", bench.D, LOSS, C)
    if mode != "single":
        parallel.DataParallel(tr, B)
    set_engine(tr.engine)
    st = tr.engine.store
    p0 = {k: v.copy() for k, v in st.export_tf().items()}
    out = []
    r = tr.train_step(x, y, 0.01, 20000, fetch_loss=True)
    torch.cuda.synchronize()
    out.append((r["raw_loss"], {k: v.copy() for k, v in st.export_tf().items()}, r["loss"]))
    for _ in range(3):
        tr.train_step(x, y, 0.01, 20000, fetch_loss=False)
    st.load_tf(p0)
    r = tr.train_step(x, y, 0.01, 20000, fetch_loss=True)
    torch.cuda.synchronize()
    out.append((r["raw_loss"], {k: v.copy() for k, v in st.export_tf().items()}, r["loss"]))
    assert tr._static[tuple(x.shape)]["graphs"] is not None
    return p0, out


def main():
    rank, world = parallel.init_from_env("nccl")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    g = torch.Generator().manual_seed(7)
    R = world * B
    m = torch.randn(R, 1, bench.D, generator=g)
    s = 0.5 + torch.rand(R, 1, bench.D, generator=g)
    xg = (m + s * torch.randn(R, T, bench.D, generator=g)).cuda()
    yg = torch.randint(0, C, (R,), generator=g, dtype=torch.int32).cuda()
    xl, yl = xg[rank * B:(rank + 1) * B].contiguous(), yg[rank * B:(rank + 1) * B].contiguous()
    p0, single = run("single", rank, world, xg, yg)
    _, sync = run("sync", rank, world, xl, yl)
    _, plain = run("replica", rank, world, xl, yl)

    def upd_err(pa, pb, k):
        da, db = (pa[k] - p0[k]).astype(np.float64), (pb[k] - p0[k]).astype(np.float64)
        na = np.linalg.norm(da)
        return 0.0 if na == 0 else float(np.linalg.norm(da - db) / na)

    ok, worst, ratio, replica_gap = True, {}, 0.0, 0.0
    for k in p0:
        nk = upd_err(single[0][1], single[1][1], k)
        if nk > 0.5:
            continue                      # zero true gradient: rounding noise only
        tol = max(3.0 * nk, 5e-3)
        for which in (0, 1):
            e = upd_err(single[which][1], sync[which][1], k)
            ratio = max(ratio, e / tol)
            if e > tol:
                ok = False
                worst["%s[%s]" % (k, "eager" if which == 0 else "graph")] = (e, nk)
        replica_gap = max(replica_gap, upd_err(single[0][1], plain[0][1], k) / tol)
    lerr = max(abs(a[0] - b[0]) / max(abs(a[0]), 1e-6) for a, b in zip(single, sync))
    terr = max(abs(a[2] - b[2]) / max(abs(a[2]), 1e-6) for a, b in zip(single, sync))       # total loss: includes the penalty
    ok = ok and lerr < 2e-3 and terr < 2e-3
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"check": "SyncBN data-parallel step (%d ranks x %d rows) == 1-GPU step on the %d-row batch%s"
                                   % (world, B, R, " [self-attention pooling, 4 heads, penalty 0.5]" if ATTENTION else ""),
                          "total_loss_single": [a[2] for a in single], "total_loss_sync_bn": [a[2] for a in sync],
                          "ok": bool(flag.item() > 0), "raw_loss_single": [a[0] for a in single],
                          "raw_loss_sync_bn": [a[0] for a in sync], "raw_loss_per_replica_bn": [a[0] for a in plain],
                          "max_rel_raw_loss": lerr, "violations": worst, "max_error_over_tolerance": ratio,
                          "per_replica_bn_error_over_tolerance": replica_gap,
                          "criterion": "per parameter: one-step update rel-Frobenius error <= max(3 x the single-GPU trainer's "
                                       "own eager-vs-replay error, 5e-3)"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
