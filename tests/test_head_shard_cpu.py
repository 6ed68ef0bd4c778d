"""World-size-2 gloo test of the class-sharded head protocol (parallel.HeadShard): the column partition, the row
all-gather, the per-row (max, sum, target) exchange and the dLoss/dx reduce-scatter must reproduce the unsharded
additive-margin head of the oracle (loss and gradients) on the concatenated global batch (SURVEY 8e 2')."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_arithmetic():
    sys.path.insert(0, ROOT)
    from tf_kaldi_speaker_b200.parallel import HeadShard
    for C, N in ((7200, 8), (7200, 2), (4300, 4), (1000, 8), (37, 2)):
        cover = []
        for r in range(N):
            sh = HeadShard(C, rank=r, world=N)
            assert sh.lo % 8 == 0 and sh.n_local > 0
            cover += list(range(sh.lo, sh.hi))
            lab = torch.arange(C, dtype=torch.int32)
            loc = sh.local_labels(lab)
            assert int((loc >= 0).sum()) == sh.n_local and int(loc.max()) == sh.n_local - 1
        assert cover == list(range(C))
    try:
        HeadShard(8, rank=0, world=2)        # 8 columns per shard: the second shard would be empty
        assert False
    except ValueError:
        pass


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import xvector_oracle as O
    from tf_kaldi_speaker_b200 import parallel
    parallel.init_from_env("gloo")
    torch.manual_seed(0)
    B, E, C = 6, 16, 21
    R = world * B
    x = torch.randn(R, E, dtype=torch.float64)
    y = torch.randint(0, C, (R,))
    W = torch.randn(E, C, dtype=torch.float64)
    p = O.ParamsPlain(amsoftmax_m=0.2, amsoftmax_lambda_min=0, amsoftmax_lambda_base=10, amsoftmax_lambda_gamma=1,
                      amsoftmax_lambda_power=1, global_step=3)
    # unsharded reference on the global batch
    xg, Wg = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    loss_ref, _ = O.additive_margin_softmax_head(xg, y, {"softmax/output/kernel": Wg}, p)
    gx_ref, gw_ref = torch.autograd.grad(loss_ref, (xg, Wg))

    sh = parallel.HeadShard(C)
    # forward exchange 1: rows of every rank
    x_loc = x[rank * B:(rank + 1) * B].contiguous()
    y_loc = y[rank * B:(rank + 1) * B].to(torch.int32).contiguous()
    x_all = torch.empty(R, E, dtype=torch.float64)
    y_all = torch.empty(R, dtype=torch.int32)
    sh.all_gather(x_all, x_loc)
    sh.all_gather(y_all, y_loc)
    ok = torch.equal(x_all, x) and torch.equal(y_all.long(), y)
    # this shard's columns: modified logits z' (margin on the target column), as the fused epilogue forms them
    xa = x_all.clone().requires_grad_(True)
    Wl = W[:, sh.lo:sh.hi].clone().requires_grad_(True)
    lab = sh.local_labels(y_all.long())
    lam = max(0.0, 10.0 * (1.0 + 1.0 * 3) ** (-1.0))
    fa = 1.0 / (1.0 + lam)
    fs = 1.0 - fa
    wn = Wl * torch.rsqrt(torch.clamp((Wl * Wl).sum(0, keepdim=True), min=1e-12))
    z = xa @ wn
    n = torch.clamp(xa.norm(dim=1), min=1e-12)
    own = lab >= 0
    zt = z[own, lab[own]]
    zmod = z.clone()
    zmod[own, lab[own]] = fs * zt + fa * n[own] * (torch.clamp(zt / n[own], -1, 1) - 0.2)
    part = torch.zeros(3, R, dtype=torch.float64)
    part[0] = zmod.max(1).values.detach()
    part[1] = torch.exp(zmod - part[0][:, None]).sum(1).detach()
    part[2, own] = zmod[own, lab[own]].detach()
    # forward exchange 2: per-row (max, sum, target) triples; combine = xv_head_combine_shards
    parts = torch.empty(world, 3, R, dtype=torch.float64)
    sh.all_gather(parts, part)
    gmax = parts[:, 0].max(0).values
    lse = torch.log((parts[:, 1] * torch.exp(parts[:, 0] - gmax)).sum(0)) + gmax
    loss = float((lse - parts[:, 2].sum(0)).sum() / R)
    ok = ok and abs(loss - float(loss_ref)) < 1e-10
    # backward: d = (softmax - onehot) / R on this shard's columns, through z' -> (x, W)
    d = torch.exp(zmod.detach() - lse[:, None]) / R
    d[own, lab[own]] -= 1.0 / R
    gx_part, gw_loc = torch.autograd.grad(zmod, (xa, Wl), grad_outputs=d)
    ok = ok and torch.allclose(gw_loc, gw_ref[:, sh.lo:sh.hi], rtol=1e-9, atol=1e-12)     # complete without any exchange
    gx_loc = torch.empty(B, E, dtype=torch.float64)
    sh.reduce_scatter_sum(gx_loc, gx_part.contiguous())
    ok = ok and torch.allclose(gx_loc, gx_ref[rank * B:(rank + 1) * B], rtol=1e-9, atol=1e-12)
    # checkpoint export: column shards -> full matrix
    full = sh.gather_columns(W[:, sh.lo:sh.hi].contiguous())
    ok = ok and torch.equal(full, W)
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world2_gloo_sharded_head_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29741
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1], r
