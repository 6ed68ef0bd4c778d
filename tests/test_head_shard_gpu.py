"""Class-sharded head on ONE GPU: S engines (one host thread and one CUDA stream each) play the S ranks; the
HeadShard exchanges are emulated through host-side barriers.  Everything else -- Engine.margin_head_sharded, the
xv_head_local_labels / xv_head_shard_partials / xv_head_combine_shards kernels, the fused GEMM epilogues on a column
slice -- is the code the NCCL path runs.  Reference: the fp64 oracle head on the concatenated global batch.
The real 2-GPU NCCL run of the same path is tools/dist_check_sharded.py (under torchrun)."""
import threading

import numpy as np
import pytest
import torch

from oracle import xvector_oracle as O

pytestmark = pytest.mark.gpu


class _ThreadShard(object):
    """parallel.HeadShard with the collectives served by shared host lists + barriers (ranks = threads)."""

    def __init__(self, base, hub):
        self.__dict__.update(base.__dict__)
        self._base, self.hub = base, hub

    def range_of(self, r):
        return self._base.range_of(r)

    def all_gather(self, out, local):
        hub = self.hub
        torch.cuda.current_stream().synchronize()
        hub["slots"][self.rank] = local.reshape(-1).clone()
        hub["bar"].wait()
        out.view(-1).copy_(torch.cat([hub["slots"][r] for r in range(self.world)]))
        torch.cuda.current_stream().synchronize()
        hub["bar"].wait()
        return out

    def reduce_scatter_sum(self, out, full):
        hub = self.hub
        torch.cuda.current_stream().synchronize()
        hub["slots"][self.rank] = full.reshape(-1).clone()
        hub["bar"].wait()
        n = out.numel()
        tot = sum(hub["slots"][r][self.rank * n:(self.rank + 1) * n] for r in range(self.world))
        out.view(-1).copy_(tot)
        torch.cuda.current_stream().synchronize()
        hub["bar"].wait()
        return out


def _rank_body(rank, S, hub, case, results):
    from tf_kaldi_speaker_b200 import _lib as L
    from tf_kaldi_speaker_b200.parallel import HeadShard
    from tf_kaldi_speaker_b200.runtime import Engine, ScaledUtt, UttAct, VarSpec, _pad_to
    try:
        torch.cuda.set_device(0)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            lt, head_type, margin, am, scaling, fa, fs = case["head"]
            x, y, w, b = case["x"], case["y"], case["w"], case["b"]
            R, E = x.shape
            Cn = w.shape[1]
            B = R // S
            eng = Engine()
            sh = _ThreadShard(HeadShard(Cn, rank=rank, world=S), hub)
            eng.head_shard = sh
            eng.inv_global_batch = 1.0 / R
            n_loc, cpad = sh.n_local, _pad_to(sh.n_local, 8)
            eng.store.declare(VarSpec("softmax/output/kernel", (E, n_loc), (E, cpad), full_shape=(E, Cn),
                                      col_range=(sh.lo, sh.hi)))
            if b is not None:
                eng.store.declare(VarSpec("softmax/output/bias", (n_loc,), (cpad,), full_shape=(Cn,),
                                          col_range=(sh.lo, sh.hi)))
            eng.store.finalize()
            eng.store.load_tf({"softmax/output/kernel": w})          # full matrix: sliced to this shard's columns
            if b is not None:
                eng.store.load_tf({"softmax/output/bias": b})
            u = UttAct(torch.from_numpy(x[rank * B:(rank + 1) * B]).cuda().contiguous(), None, "u")
            u.needs_grad = True
            feats = ScaledUtt(u, scaling) if scaling > 0 else u
            labels = torch.from_numpy(y[rank * B:(rank + 1) * B].astype(np.int32)).cuda()
            eng.begin_step(True)
            eng.set_sched(fa, fs)
            loss, _, _ = eng.margin_head_sharded(feats, labels, "softmax/output/kernel",
                                                 "softmax/output/bias" if b is not None else None, head_type, Cn, True,
                                                 margin=margin, asoftmax_m=am, scaling=scaling)
            eng.backward()
            stream.synchronize()
            g = eng.store.export_tf(grads=True)
            results[rank] = dict(loss=float(loss.item()), du=u.grad.double().cpu(), lo=sh.lo, hi=sh.hi,
                                 gw=torch.from_numpy(g["softmax/output/kernel"]).double(),
                                 gb=(torch.from_numpy(g["softmax/output/bias"]).double() if b is not None else None))
    except Exception as ex:         # release the peers, report in the main thread
        results[rank] = ex
        hub["bar"].abort()


def _run_sharded(S, case):
    hub = {"bar": threading.Barrier(S), "slots": [None] * S}
    results = [None] * S
    th = [threading.Thread(target=_rank_body, args=(r, S, hub, case, results)) for r in range(S)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    for r in results:
        if isinstance(r, Exception):
            raise r
        assert r is not None, "a rank thread did not finish"
    return results


@pytest.mark.parametrize("S", [2, 3])
@pytest.mark.parametrize("lt", ["additive_angular_margin_softmax", "asoftmax", "softmax"])
def test_sharded_head_matches_global_oracle(S, lt):
    from tf_kaldi_speaker_b200 import _lib as L
    from tf_kaldi_speaker_b200.model.loss import margin_lambda
    g = torch.Generator().manual_seed(11 + S)
    B, E, Cn = 24, 512, 1003                       # ragged: 1003 classes -> shards of 504|499 or 336|336|331 columns
    R = S * B
    x = torch.randn(R, E, generator=g)
    y = torch.randint(0, Cn, (R,), generator=g)
    y[0], y[1], y[2] = 0, Cn - 1, (Cn // S + 7) // 8 * 8      # first / last class, first column of shard 1
    w = (torch.rand(E, Cn, generator=g) - 0.5) * 0.2
    pd = dict(weight_l2_regularizer=1e-4, global_step=2000, feature_norm=(lt != "softmax"), feature_scaling_factor=20.0)
    for pre in ("asoftmax", "amsoftmax", "arcsoftmax"):
        pd.update({pre + "_lambda_min": 0, pre + "_lambda_base": 1000, pre + "_lambda_gamma": 1e-2, pre + "_lambda_power": 3})
    pd["asoftmax_m"], pd["arcsoftmax_m"] = 4, 0.2
    _, fa, fs = margin_lambda(0, 1000, 1e-2, 3, 2000)
    b = None
    if lt == "softmax":
        head = (lt, L.HEAD_SOFTMAX, 0.0, 1, 0.0, 1.0, 0.0)
        b = (torch.randn(Cn, generator=g) * 0.1).numpy()
    elif lt == "asoftmax":
        head = (lt, L.HEAD_ASOFTMAX, 0.0, 4, 20.0, fa, fs)
    else:
        head = (lt, L.HEAD_AAM, 0.2, 1, 20.0, fa, fs)
    case = dict(head=head, x=x.numpy(), y=y.numpy(), w=w.numpy(), b=b)
    res = _run_sharded(S, case)

    # fp64 oracle on the global batch
    xd = x.double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    P = {"softmax/output/kernel": wd}
    if b is not None:
        bd = torch.from_numpy(b).double().requires_grad_(True)
        P["softmax/output/bias"] = bd
        lo, _ = O.softmax_head(xd, y, P)
        gx, gw, gb = torch.autograd.grad(lo, [xd, wd, bd])
    else:
        po = O.ParamsPlain(**pd)
        lo, _ = O.loss_network(lt, O.l2_scaling(xd, 20.0), y, P, po)
        gx, gw = torch.autograd.grad(lo, [xd, wd])
    for r, out in enumerate(res):
        assert abs(out["loss"] - lo.item()) <= 2e-4 * abs(lo.item()) + 1e-5, (r, out["loss"], lo.item())   # global mean on every rank
        ex = (out["du"] - gx[r * B:(r + 1) * B]).norm() / gx[r * B:(r + 1) * B].norm()
        ew = (out["gw"] - gw[:, out["lo"]:out["hi"]]).norm() / gw[:, out["lo"]:out["hi"]].norm()
        assert ex < 2e-2 and ew < 2e-2, (r, float(ex), float(ew))
        if b is not None:
            eb = (out["gb"] - gb[out["lo"]:out["hi"]]).norm() / gb[out["lo"]:out["hi"]].norm()
            assert eb < 2e-2, (r, float(eb))
