"""World-size-2 gloo test of the data-parallel host logic (batch sharding, flat-buffer all-reduce, scalar mean).
The N-GPU semantics to preserve: summing the per-rank gradients of sum_i CE_i / (N*B) equals the gradient of the
mean loss over the concatenated global batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle import xvector_oracle as O
    from tf_kaldi_speaker_b200 import parallel
    r, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    B, E, C = 8, 16, 5
    x = torch.randn(B, E, dtype=torch.float64)
    y = torch.randint(0, C, (B,))
    W = torch.randn(E, C, dtype=torch.float64)
    # global-batch reference
    Wg = W.clone().requires_grad_(True)
    p = O.ParamsPlain(amsoftmax_m=0.2, amsoftmax_lambda_min=0, amsoftmax_lambda_base=10, amsoftmax_lambda_gamma=1,
                      amsoftmax_lambda_power=1, global_step=3)
    loss, _ = O.additive_margin_softmax_head(x, y, {"softmax/output/kernel": Wg}, p)
    (gref,) = torch.autograd.grad(loss, Wg)
    # this rank's shard, loss scaled by 1/(N*B_local) = local mean / N
    xs, ys = parallel.shard_batch(x, y, rank, world)
    Wl = W.clone().requires_grad_(True)
    ll, _ = O.additive_margin_softmax_head(xs, ys, {"softmax/output/kernel": Wl}, p)
    (gl,) = torch.autograd.grad(ll / world, Wl)
    comm = parallel.FlatAllReduce()
    flat = gl.reshape(-1).clone()
    comm.allreduce_(flat)
    ok_grad = torch.allclose(flat.reshape(E, C), gref, rtol=1e-10, atol=1e-12)
    tot = comm.mean_scalar(float(ll) / world)
    ok_loss = abs(tot - float(loss)) < 1e-10
    params = torch.full((4,), float(rank))
    comm.broadcast_(params, src=0)
    ok_bcast = bool((params == 0).all())
    # DataParallel on a stand-in parameter store: ONE exchange covers the flat gradient buffer AND the step's loss scalars
    # (the first 32 floats of the zero arena behind it), so the logged loss needs no second collective
    class _Spec(object):
        def __init__(self, offset):
            self.offset = offset

    class _Store(object):
        def __init__(self):
            self.specs = {"tdnn/tdnn1_conv/kernel": _Spec(0), "tdnn/tdnn6_dense/kernel": _Spec(2048)}
            self.n = 4096
            self.params = torch.zeros(self.n)
            self.buffers = torch.zeros(32)
            self.grads_ext = torch.cat([torch.arange(self.n, dtype=torch.float32) * (rank + 1), torch.zeros(64)])
            self.grads = self.grads_ext[:self.n]
            self.arena = self.grads_ext[self.n:]
            self.arena_used = 0

        def refresh_shadows(self):
            pass

    class _Eng(object):
        _scalars = None

        @property
        def scalars(self):
            if self._scalars is None:
                self._scalars = self.store.arena[:8]
                self.store.arena_used = 32
            return self._scalars

    class _Trainer(object):
        pass
    tr = _Trainer()
    tr.engine = _Eng()
    tr.engine.store = _Store()
    tr.engine.device = "cpu"
    dp = parallel.DataParallel(tr, local_batch=4)
    ok_dp = dp.scalars_reduced and dp.reduce_numel == 4096 + 32 and abs(tr.engine.inv_global_batch - 1.0 / (world * 4)) < 1e-12
    tr.engine.scalars[0] = float(rank + 1)           # this rank's share of the global-batch mean loss
    dp.allreduce_gradients()
    expect = torch.arange(4096, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok_dp = ok_dp and torch.equal(tr.engine.store.grads, expect)
    ok_dp = ok_dp and float(tr.engine.scalars[0]) == float(sum(r + 1 for r in range(world)))
    out.put((rank, ok_grad, ok_loss, ok_bcast and ok_dp))
    dist.destroy_process_group()


def test_world2_gloo_allreduce_equals_global_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29731
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r


def _syncbn_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from tf_kaldi_speaker_b200 import parallel
    parallel.init_from_env("gloo")
    torch.manual_seed(1)
    B, C = 5, 7
    y_all = torch.randn(world * B, C, dtype=torch.float64)
    y = y_all[rank * B:(rank + 1) * B].contiguous()
    sync = parallel.SyncBN()
    # frame-level rule: all-reduced (sum, sum of squares) with the global count == statistics of the concatenated batch
    stats = torch.stack([y.sum(0), (y * y).sum(0)])
    sync.all_reduce_sum_(stats)
    n = world * B
    mean, var = stats[0] / n, stats[1] / n - (stats[0] / n) ** 2
    ok = torch.allclose(mean, y_all.mean(0)) and torch.allclose(var, y_all.var(0, unbiased=False))
    # utterance-level rule: rows all-gathered, this rank's rows sit at [rank*B, (rank+1)*B)
    rows = torch.empty(world * B, C, dtype=torch.float64)
    sync.all_gather(rows, y)
    ok = ok and torch.equal(rows, y_all) and sync.world == world and sync.rank == rank
    # gradients that every rank already holds in full are pre-scaled by 1/N so that the final sum all-reduce restores them
    g = torch.full((3,), 6.0, dtype=torch.float64) / world
    sync.all_reduce_sum_(g)
    ok = ok and torch.allclose(g, torch.full((3,), 6.0, dtype=torch.float64))
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world2_gloo_syncbn_exchanges():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, 29733, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] for r in res), res
