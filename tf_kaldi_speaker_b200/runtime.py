"""Device-side runtime of the x-vector path: parameter store, workspaces and the kernel sequencer ("Engine").

The reference builds a TF1 graph and lets ``sess.run`` execute it (model/trainer.py:351-436, 491-508).  Here
the operator functions in ``model/`` call Engine methods that enqueue hand-written sm_100a kernels (through the
C ABI in libxvector_b200.so) on the current CUDA stream and push their backward closures on a tape; PyTorch
only owns device memory and streams.  There is no CPU fallback.

Data layout in HBM:
  * parameters: ONE flat fp32 buffer (tensors start at multiples of 1024 elements) + one flat fp32 gradient
    buffer of the same layout (a single NCCL all-reduce covers every gradient) + optimizer state + a flat bf16
    "shadow" buffer holding the tensor-core copies of the kernels (plain, or the [hi; lo; hi] 3-term split).
  * frame-level activations: bf16, channels-last, flat-time [B*T, C_pad]; the row stride T is constant through
    tdnn1..5 and rows with t >= valid length are kept at zero ("invalid rows").
  * utterance-level activations: fp32 [B, C] plus a bf16 [hi | hi | lo] split copy feeding the GEMMs.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L

ALIGN = 1024
BN_EPS = 1e-3
ZERO_ARENA = 1 << 20          # floats (4 MB) of step-start-zero scratch behind the gradient buffer


def _pad_to(n, m):
    return (n + m - 1) // m * m


class VarSpec(object):
    def __init__(self, name, tf_shape, ishape, row_map=None, l2=0.0, shadow="none", init="zeros", trainable=True,
                 fans=None, pad_value=0.0, full_shape=None, col_range=None):
        self.name, self.tf_shape, self.ishape = name, tuple(tf_shape), tuple(ishape)
        # class-sharded head: this rank holds columns [col_range) of a variable whose unsharded TF shape is full_shape
        self.full_shape = None if full_shape is None else tuple(full_shape)
        self.col_range = col_range
        self.row_map, self.l2, self.shadow, self.init = row_map, float(l2), shadow, init
        self.trainable, self.fans, self.pad_value = trainable, fans, pad_value
        self.numel = int(np.prod(self.ishape))
        self.offset = None
        self.shadow_offset = None

    # TF-shaped numpy <-> internal (padded) numpy
    def to_internal(self, arr):
        arr = np.asarray(arr, dtype=np.float32)
        assert tuple(arr.shape) == self.tf_shape, "%s: expected %s, got %s" % (self.name, self.tf_shape, arr.shape)
        out = np.full(self.ishape, self.pad_value, dtype=np.float32)
        if len(self.ishape) == 1:
            if self.row_map is not None:      # 1-D index map (e.g. [mean | std] halves of a channel-padded vector)
                if self.pad_value != 0.0:
                    out[:] = self.pad_value
                out[self.row_map] = arr.reshape(-1)
            else:
                out[:arr.size] = arr.reshape(-1)
            return out
        a2 = arr.reshape(-1, arr.shape[-1])
        rows = np.arange(a2.shape[0]) if self.row_map is None else self.row_map
        out[rows, :a2.shape[1]] = a2
        if self.pad_value != 0.0:       # padded rows of a matrix are always zero
            mask = np.ones(self.ishape[0], dtype=bool)
            mask[rows] = False
            out[mask] = 0.0
            out[:, a2.shape[1]:] = 0.0
        return out

    def to_tf(self, arr):
        arr = np.asarray(arr, dtype=np.float32).reshape(self.ishape)
        if len(self.ishape) == 1:
            if self.row_map is not None:
                return arr[self.row_map].reshape(self.tf_shape).copy()
            return arr[:int(np.prod(self.tf_shape))].reshape(self.tf_shape).copy()
        ncol = self.tf_shape[-1]
        nrow = int(np.prod(self.tf_shape[:-1]))
        rows = np.arange(nrow) if self.row_map is None else self.row_map
        return arr[rows, :ncol].reshape(self.tf_shape).copy()


class ParamStore(object):
    """Name-keyed variables (TF variable names = checkpoint keys) in flat device buffers."""

    def __init__(self, device):
        self.device = device
        self.specs = OrderedDict()
        self.finalized = False

    def declare(self, spec):
        assert not self.finalized
        if spec.name in self.specs:
            return self.specs[spec.name]
        self.specs[spec.name] = spec
        return spec

    def __contains__(self, name):
        return name in self.specs

    def finalize(self):
        off = soff = boff = 0
        for s in self.specs.values():
            if s.trainable:
                s.offset = off
                off += _pad_to(s.numel, ALIGN)
                if s.shadow == "plain":
                    s.shadow_offset = soff
                    soff += _pad_to(s.numel, ALIGN)
                elif s.shadow == "split":
                    assert s.numel % ALIGN == 0, "%s: split shadows need numel %% 1024 == 0" % s.name
                    s.shadow_offset = soff
                    soff += 3 * s.numel
            else:
                s.offset = boff
                boff += _pad_to(s.numel, 32)
        self.n = max(off, ALIGN)
        dev = self.device
        self.params = torch.zeros(self.n, dtype=torch.float32, device=dev)
        # gradients + a "zero arena" in ONE allocation: every accumulator that must start a step at zero (BN statistics,
        # split-K outputs, loss scalars) is carved from the arena, so a step begins with a single fill kernel
        self.grads_ext = torch.zeros(self.n + ZERO_ARENA, dtype=torch.float32, device=dev)
        self.grads = self.grads_ext[:self.n]
        self.arena = self.grads_ext[self.n:]
        self.arena_used = 0
        self.state1 = None
        self.state2 = None
        self.shadow = torch.zeros(max(soff, ALIGN), dtype=torch.bfloat16, device=dev)
        self.buffers = torch.zeros(max(boff, 32), dtype=torch.float32, device=dev)
        nblk = self.n // ALIGN
        l2 = np.zeros(nblk, dtype=np.float32)
        sh = np.full(nblk, -1, dtype=np.int64)
        st = np.zeros(nblk, dtype=np.int64)
        for s in self.specs.values():
            if not s.trainable:
                continue
            b0, nb = s.offset // ALIGN, _pad_to(s.numel, ALIGN) // ALIGN
            l2[b0:b0 + nb] = s.l2
            if s.shadow != "none":
                sh[b0:b0 + nb] = s.shadow_offset + np.arange(nb) * ALIGN
                if s.shadow == "split":
                    st[b0:b0 + nb] = s.numel
        self.blk_l2 = torch.from_numpy(l2).to(dev)
        self.blk_shadow = torch.from_numpy(sh).to(dev)
        self.blk_stride = torch.from_numpy(st).to(dev)
        self.finalized = True

    # ---- views
    def _buf(self, s):
        return self.params if s.trainable else self.buffers

    def view(self, name):
        s = self.specs[name]
        return self._buf(s)[s.offset:s.offset + s.numel].view(*s.ishape)

    def grad(self, name):
        s = self.specs[name]
        return self.grads[s.offset:s.offset + s.numel].view(*s.ishape)

    def shadow_view(self, name):
        """bf16 [rows (x3 for split), cols] tensor-core copy of a kernel."""
        s = self.specs[name]
        assert s.shadow != "none"
        rows, cols = s.ishape
        if s.shadow == "split":
            return self.shadow[s.shadow_offset:s.shadow_offset + 3 * s.numel].view(3 * rows, cols)
        return self.shadow[s.shadow_offset:s.shadow_offset + s.numel].view(rows, cols)

    # ---- values
    def load_tf(self, values):
        """values: name -> array in TF shape (missing names keep their current value)."""
        for name, arr in values.items():
            if name not in self.specs:
                continue
            if isinstance(arr, torch.Tensor):
                arr = arr.detach().cpu().double().numpy()
            s = self.specs[name]
            if s.col_range is not None and tuple(np.shape(arr)) == s.full_shape and s.full_shape != s.tf_shape:
                arr = np.asarray(arr)[..., s.col_range[0]:s.col_range[1]]      # unsharded checkpoint -> this rank's columns
            self.view(name).copy_(torch.from_numpy(s.to_internal(arr)).to(self.device))
        self.refresh_shadows()

    def export_tf(self, grads=False, which=None):
        """name -> array in TF shape.  ``which``: "params" (default), "grads", "state1" / "state2" (optimizer slots:
        Momentum accumulator / Adam m, Adam v; trainable variables only)."""
        which = which or ("grads" if grads else "params")
        out = OrderedDict()
        for name, s in self.specs.items():
            if which != "params" and not s.trainable:
                continue
            if which == "params":
                t = self.view(name)
            elif which == "grads":
                t = self.grad(name)
            else:
                flat = getattr(self, which)
                if flat is None:
                    return out
                t = flat[s.offset:s.offset + s.numel].view(*s.ishape)
            out[name] = s.to_tf(t.detach().cpu().numpy())
        return out

    def init(self, seed=0):
        gen = torch.Generator().manual_seed(seed)
        vals = {}
        for name, s in self.specs.items():
            shp = s.full_shape if s.full_shape is not None else s.tf_shape   # shards draw the full tensor: same stream
            if s.init == "glorot":
                fan_in, fan_out = s.fans
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                v = (torch.rand(shp, generator=gen, dtype=torch.float64) * 2 - 1) * lim
            elif s.init == "ones":
                v = torch.ones(shp, dtype=torch.float64)
            elif s.init == "trunc_normal":
                v = torch.empty(shp, dtype=torch.float64)
                torch.nn.init.trunc_normal_(v, 0.0, 0.1, -0.2, 0.2, generator=gen)
            elif isinstance(s.init, float):
                v = torch.full(shp, s.init, dtype=torch.float64)
            else:
                v = torch.zeros(shp, dtype=torch.float64)
            vals[name] = v.numpy()
        self.load_tf(vals)          # column-sharded variables are sliced there

    def refresh_shadows(self):
        L.check(L.load().xv_shadow_refresh(L.ptr(self.params), L.ptr(self.blk_shadow), L.ptr(self.blk_stride),
                                           L.ptr(self.shadow), C.c_int64(self.n), L.stream_ptr()))

    def ensure_opt_state(self, opt):
        if opt != L.OPT_SGD and self.state1 is None:
            self.state1 = torch.zeros_like(self.params)
        if opt == L.OPT_ADAM and self.state2 is None:
            self.state2 = torch.zeros_like(self.params)


class FrameAct(object):
    """Handle of a frame-level tensor: bf16 [B*T, ld] flat-time, valid length per segment."""

    def __init__(self, data, B, T, valid, C_real, lengths=None, name="", ld=None):
        self.data, self.B, self.T, self.valid, self.C, self.lengths, self.name = data, B, T, valid, C_real, lengths, name
        self.grad = None
        self.needs_grad = False
        self.ld = data.shape[1] if data is not None else ld
        self.lazy = None          # (y, scale, shift, alpha, act): BN+activation not applied yet (fused into pooling)
        self.pool_grad = None     # (pooled, dpooled): upstream gradient given implicitly by the statistics pooling
        self.consumers = 0        # layers reading this tensor (a single consumer lets its dgrad fuse our BN-backward reductions)
        self.bn_bwd = None        # (y, scale, shift, mean, rstd, neg_slope, dgamma, dbeta) of the layer that produced us
        self.bn_reduced = False   # dgamma / dbeta already accumulated by the consumer's dgrad epilogue
        self.pool_sums = None     # [B, 4, C] sums of the fused pooling forward (BN backward reductions without a pass over y)
        self.bias = None          # layer bias of a pre-BN tensor (stored bias-free)
        self._materialize = None

    def materialize(self):
        if self.data is None:
            self._materialize()
        return self.data

    def dense(self):
        """fp32 [B, valid, C] copy (inspection / endpoints only; uniform valid length)."""
        d = self.materialize().view(self.B, self.T, self.ld)[:, :self.valid, :self.C].float()
        if self.bias is not None:        # pre-BN tensors are stored bias-free
            d = d + self.bias[:self.C]
        return d

    def lengths_ptr(self):
        return L.ptr(self.lengths)


class UttAct(object):
    """Handle of an utterance-level tensor: fp32 [B, C] (+ optional bf16 split copy [B, 3C])."""

    def __init__(self, data, split=None, name="", col_map=None):
        self.data, self.split, self.name, self.col_map = data, split, name, col_map
        self.grad = None
        self.needs_grad = False

    def dense(self):
        if self.col_map is None:
            return self.data
        c, cpad = self.col_map            # [mean | std] halves of a channel-padded pooling output
        return torch.cat([self.data[:, :c], self.data[:, cpad:cpad + c]], 1)


class ScaledUtt(UttAct):
    """View of an utterance tensor with a pending l2_scaling (model/trainer.py:183-186, model/common.py:45-58):
    the scaling itself is fused into the head's feature-preparation kernel; gradients land on the base."""

    def __init__(self, base, scaling):
        self.base, self.scaling = base, float(scaling)
        self.data, self.split, self.name, self.col_map = base.data, None, base.name, base.col_map
        self.needs_grad = base.needs_grad

    @property
    def grad(self):
        return self.base.grad

    @grad.setter
    def grad(self, g):
        self.base.grad = g

    def dense(self):
        x = self.base.dense()
        sq = (x * x).sum(-1, keepdim=True)
        return x * (torch.rsqrt(torch.clamp(sq, min=1e-12)) * self.scaling)


class Engine(object):
    """Sequences the sm_100a kernels of one replica on the current stream."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise L.XvError("xvector_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.lib = L.load()
        info = (C.c_int32 * 3)()
        L.check(self.lib.xv_device_info(info))
        self.num_sms = int(info[0])
        if int(info[1]) != 10:
            raise L.XvError("xvector_b200 kernels are sm_100a only (found cc %d.%d)" % (info[1], info[2]))
        self.store = ParamStore(self.device)
        self.ws = {}                     # (name, shape, dtype) -> view
        self._flat = {}                  # (name, dtype) -> backing allocation of the scratch buffers (largest shape seen)
        self._step_shapes = {}
        self.ws_generation = 0           # bumped whenever a backing allocation is replaced by a larger one
        self._arena_bufs = set()
        self.tape = []
        self.penalties = []
        self.launches = 0
        self._hs = torch.zeros(16, dtype=torch.float32, device=self.device)      # host-fed scalars, one launch per step
        self._hs[8] = 1.0
        self.hyper = self._hs[:8]
        self.sched = self._hs[8:10]                                              # (fa, fs) of the margin schedule
        self._hs_host = [0.0] * 8 + [1.0, 0.0]
        self._scalars = None             # [0] loss, [1] l2 loss, [2] grad sumsq, [3] penalty (lives in the zero arena)
        self.inv_global_batch = None     # set by the data-parallel wrapper (1 / (N * B))
        self.capturing = False           # True while a CUDA graph of the step is being captured
        self.epilogue_stats = True       # False: BN batch statistics from the separate xv_col_stats pass (tests)
        self.fuse_bn_bwd = True          # dgrad epilogues accumulate the producer layer's BN dgamma / dbeta
        import os as _os
        # shortest dgrad K loop whose epilogue also forms the producer's BN-backward sums (a separate reduce pass over y and
        # dX costs ~15 us at config 2; the packed-f32x2 column pass +4..8 us per launch).  XV_FUSE_BNBWD_MINK=512 also fuses the
        # K = 512 dgrad of tdnn4 (+7.7 us on an 18 us launch against a 14.7 us reduce kernel): 0.9326 -> 0.9276 ms per step in
        # one A/B pair, inside the run-to-run spread -- left off so that the GEMM launch does GEMM work only
        self.fuse_bn_bwd_min_k = int(_os.environ.get("XV_FUSE_BNBWD_MINK", "1024"))
        self.fold_inference_bn = True    # inference: BN (moving statistics) + relu in the GEMM epilogue, no separate apply pass
        self.side_wgrad = False          # frame-level wgrad GEMMs on a second stream: measured 1.067 vs 1.056 ms (no gain)
        # utterance-level weight-gradient work (tdnn6 / tdnn7 / head dW GEMMs, head_finish_dw) and the head's weight
        # normalisation run on a second stream: the utterance-level chain is ~20 dependent, latency-bound launches that
        # leave most SMs idle, and nothing but the optimizer consumes these results
        self.side_utt = True
        self._side = None
        self._side_used = False
        self._head_prefetch = None       # (kernel name, normalize) whose bf16 operand the side stream is preparing
        self.head_shard = None           # parallel.HeadShard: the speaker matrix is split by columns over the ranks
        self.sync_bn = None              # parallel.SyncBN: batch-norm statistics over the GLOBAL batch (all ranks)
        self.segmenter = None            # SegmentedGraph while a step containing collectives is being captured

    def collective(self, fn):
        """Run a host-side exchange (torch.distributed call on named workspace buffers) between kernels.  While a step is
        being captured the exchange cuts the CUDA graph: kernels before / after it land in consecutive graph segments
        and the exchange itself is replayed eagerly between them (SegmentedGraph)."""
        self.join_side_stream()          # a graph segment may not end with forked work in flight; the exchange follows both
        if self.capturing:
            if self.segmenter is None:
                raise L.XvError("a collective inside a captured step needs a SegmentedGraph")
            self.segmenter.cut(fn)
        else:
            fn()

    # ---- memory
    @property
    def scalars(self):
        if self._scalars is None:
            self._scalars = self._arena_alloc(8)
            if self._scalars is None:
                self._scalars = torch.zeros(8, dtype=torch.float32, device=self.device)
        return self._scalars

    def _arena_alloc(self, numel):
        st = self.store
        if not st.finalized:
            return None
        n = _pad_to(numel, 32)
        if st.arena_used + n > st.arena.numel():
            return None
        t = st.arena[st.arena_used:st.arena_used + numel]
        st.arena_used += n
        return t

    def buf(self, name, shape, dtype, zero=False):
        """Named workspace tensor (allocated once per name/shape).  ``zero=True``: the buffer starts every step at zero;
        fp32 buffers come from the zero arena cleared by begin_step's single fill, others are cleared here."""
        if not getattr(self, "training", True) and not self.capturing:
            # Inference (extraction over ragged batches): every distinct padded length would otherwise allocate -- and
            # zero-fill -- its own multi-GB set of activations.  One growing flat buffer per name instead; each kernel
            # writes its full output (invalid rows / padded channels included), so stale contents are never read.
            numel = int(np.prod(shape))
            ikey = ("infer", name, dtype)
            flat = self.ws.get(ikey)
            if flat is None or flat.numel() < numel:
                flat = torch.empty(max(int(numel * 1.25), 1024), dtype=dtype, device=self.device)
                self.ws[ikey] = flat
            t = flat[:numel].view(*shape)
            if zero:
                t.zero_()
            return t
        key = (name, tuple(shape), dtype)
        t = self.ws.get(key)
        if t is None:
            if zero and dtype == torch.float32:
                flat = self._arena_alloc(int(np.prod(shape)))
                if flat is not None:
                    t = flat.view(*shape)
                    self.ws[key] = t
                    self._arena_bufs.add(key)
                    t.zero_()
                    return t
            if zero:
                t = torch.zeros(shape, dtype=dtype, device=self.device)
                self.ws[key] = t
                return t
            # Scratch buffers are backed by ONE flat allocation per name, sized for the largest shape seen: the training
            # loop draws a new segment length per batch (data_loader.py:273, 200..400 frames), and a full set of activations
            # per distinct length would not fit in HBM.  Every length sees the same base address; a captured step of one
            # length stays valid as long as no buffer has to grow (ws_generation; Trainer.reserve sizes them up front).
            numel = int(np.prod(shape))
            fkey = (name, dtype)
            prev = self._step_shapes.get(name)
            if prev is not None and prev != tuple(shape):      # same name, two shapes inside one step: never alias those
                fkey = (name, dtype, tuple(shape))
            self._step_shapes[name] = tuple(shape)
            flat = self._flat.get(fkey)
            if flat is None or flat.numel() < numel:
                if flat is not None:
                    self.ws_generation += 1
                    for k in [k for k in self.ws if len(k) == 3 and k[0] == name and k[2] == dtype and k not in self._arena_bufs]:
                        del self.ws[k]
                flat = torch.zeros(numel, dtype=dtype, device=self.device)
                self._flat[fkey] = flat
            t = flat[:numel].view(*shape)
            self.ws[key] = t
        elif zero and key not in self._arena_bufs:
            t.zero_()
        else:
            self._step_shapes.setdefault(name, tuple(shape))
        return t

    def call(self, fn, *args):
        self.launches += 1
        L.check(fn(*args))

    def gemm(self, *a, **kw):
        self.launches += 1
        L.gemm(*a, **kw)

    def splits_for(self, M, N, K, max_splits=64, min_kb=4):
        """Split-K factor of an f32-output GEMM: the persistent grid runs ceil(tiles*s / units) rounds of equal-length
        work units, so occupancy is tiles*s / (rounds * units).  Pick the smallest s within 6 % (relative) of the best
        occupancy (fewer splits = fewer TMA reduce-add epilogues); every split keeps >= ``min_kb`` 64-deep k-blocks.
        Tile shape as xv_gemm_bf16 chooses it: 256 x 256 on the 74 SM pairs once there are >= 64 such tiles, else
        128 x 128 on the 148 SMs."""
        pair = ((M + 255) // 256) * ((N + 255) // 256)
        single = ((M + 127) // 128) * ((N + 127) // 128)
        kb = (K + 63) // 64
        smax = max(1, min(max_splits, kb // min_kb))
        # a problem that can reach the 64 pair tiles is kept on the CTA-pair kernel (256 x 256 tiles need a third less
        # L2 -> SM traffic per FLOP than 128 x 128 ones: tdnn5's wgrad ran 51 us on single tiles, 39 us on pairs)
        smin = 1
        pairs_ok = M > 128           # a single 128-row block would leave half of every pair tile empty
        if pairs_ok and pair * smax >= 64:
            smin = -(-64 // pair)
        effs = {}
        for s_ in range(smin, smax + 1):
            if pairs_ok and pair * s_ >= 64:
                units, cap = pair * s_, self.num_sms // 2
            else:
                units, cap = single * s_, self.num_sms
            rounds = (units + cap - 1) // cap
            effs[s_] = units / float(rounds * cap)
        best = max(effs.values())
        for s_ in sorted(effs):
            if effs[s_] >= 0.94 * best:
                return s_
        return 1

    # ---- step bookkeeping
    def begin_step(self, training):
        self.tape = []
        self.tape_mark = None
        self._step_shapes = {}
        self.penalties = []
        self.training = training
        sc = self.scalars                   # make sure the scalars exist before the fill
        st = self.store
        if sc.data_ptr() < st.arena.data_ptr() or sc.data_ptr() >= st.arena.data_ptr() + 4 * st.arena.numel():
            sc.zero_()
        if training:
            st.grads_ext.zero_()            # gradients + zero arena (BN statistics, split-K outputs, loss scalars)
        else:
            st.arena[:max(st.arena_used, 32)].zero_()

    def mark_utterance_level(self):
        """Called by the network builder right after the pooling layer: backward closures recorded from here on belong to
        the utterance-level layers and the head, whose gradients are complete early in the backward pass (the
        data-parallel wrapper all-reduces them while the frame-level backward is still running)."""
        self.tape_mark = len(self.tape)

    def on_side_stream(self, enabled=None):
        """Context manager: work enqueued inside runs on the side stream, ordered after everything already enqueued on
        the current stream (fork); join_side_stream() makes the current stream wait for it."""
        import contextlib
        if not (self.side_wgrad if enabled is None else enabled):
            return contextlib.nullcontext()
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(torch.cuda.current_stream())
        self._side_used = True
        return torch.cuda.stream(self._side)

    def join_side_stream(self):
        if self._side_used:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_used = False

    def backward(self):
        fns, self.tape = self.tape, []
        for fn in reversed(fns):
            fn()
        self.join_side_stream()

    # ---- frame-level ops ------------------------------------------------------------------------
    def pack_input(self, features, lengths=None, k=5, dpad=32):
        """features f32 [B, T, D] on device -> im2col FrameAct [B*T, 192] for tdnn1 (K = 5*32 padded to 192)."""
        assert features.is_cuda and features.dtype == torch.float32 and features.dim() == 3
        features = features.contiguous()
        B, T, D = features.shape
        assert D <= dpad
        ldo = _pad_to(k * dpad, 64)
        out = self.buf("input/col", (B * T, ldo), torch.bfloat16)
        self.call(self.lib.xv_pack_input, L.ptr(features), L.ptr(out), B, T, D, k, dpad, C.c_int64(ldo), L.stream_ptr())
        ln = None
        if lengths is not None:
            ln = lengths.to(device=self.device, dtype=torch.int32) - (k - 1)
        return FrameAct(out, B, T, T - (k - 1), k * dpad, ln, "input")

    def frame_affine(self, x, kernel, bias, k, cout, name, training, bn=None, act=L.ACT_RELU, alpha=None,
                     unbiased_moving_var=False, momentum=0.99, defer_apply=False):
        """affine (temporal conv of width k as implicit GEMM, or dense) -> [BN] -> activation.
        Returns (pre-BN FrameAct y, post-activation FrameAct a).  bn = (gamma, beta, moving_mean, moving_var) names."""
        st = self.store
        W = st.shadow_view(kernel)                     # [k*cin_pad, cout_pad] bf16
        K, cout_pad = W.shape
        R = x.B * x.T
        shrink = k - 1
        valid = x.valid - shrink
        lengths = None if x.lengths is None else (x.lengths - shrink)
        assert K == k * x.ld, "%s: kernel rows %d != k*ld %d" % (name, K, k * x.ld)
        # Inference (extraction): scale / shift are known before the GEMM (moving statistics), so BN + relu / leaky_relu run
        # in the GEMM epilogue and the activation is written directly -- the pre-BN tensor and the separate apply pass
        # (2 x 2 bytes per element written + read) disappear.  The pre-BN endpoint stays available lazily.
        fold_infer = (not training and not defer_apply and alpha is None and act in (L.ACT_NONE, L.ACT_RELU, L.ACT_LRELU)
                      and self.fold_inference_bn)
        y = None if fold_infer else self.buf(name + "/y", (R, cout_pad), torch.bfloat16)
        use_stats = training and bn is not None
        if use_stats:
            stats = self.buf(name + "/stats", (2, cout_pad), torch.float32, zero=True)
        xd = x.materialize()
        x.consumers += 1
        a_op = L.operand(xd, False, div=(x.ld if k > 1 else 0), tap_rows=(1 if k > 1 else 0))
        # BN statistics come out of the GEMM epilogue (warp-shuffle column sums + shared-memory atomics, one global
        # atomic per tile column): +4 us on the K=512 layers against 17-45 us for a separate pass over y.
        epi_stats = use_stats and self.epilogue_stats
        # y is stored WITHOUT the layer bias: in front of a batch-norm the bias cancels (its gradient is exactly zero), so it
        # is folded into the BN shift (inference) / the moving mean (training) instead of costing an epilogue pass; a layer
        # without BN gets it through shift = bias.
        def run_plain_gemm(y=y):
            self.gemm(a_op, L.operand(W, True), R, cout_pad, K, y, epilogue=L.EPI_BF16,
                      col_sum=stats[0] if epi_stats else None, col_sumsq=stats[1] if epi_stats else None,
                      seg_len=x.T, seg_valid=valid)
        if not fold_infer:
            run_plain_gemm()
        if use_stats and not epi_stats:
            self.call(self.lib.xv_col_stats, L.ptr(y), L.ptr(None), C.c_int64(R), cout_pad, C.c_int64(cout_pad),
                      x.T, valid, L.ptr(lengths), L.ptr(stats[0]), L.ptr(stats[1]), L.stream_ptr())
        if use_stats and lengths is not None:
            raise NotImplementedError("training-mode BN needs one valid length per batch (data_loader.py:273)")
        sync = self.sync_bn if (use_stats and self.sync_bn is not None and self.sync_bn.world > 1) else None
        if sync is not None:      # SyncBN: per-channel (sum, sum of squares) over every rank's rows
            self.collective(lambda: sync.all_reduce_sum_(stats))
        scale = self.buf(name + "/scale", (cout_pad,), torch.float32)
        shift = self.buf(name + "/shift", (cout_pad,), torch.float32)
        smean = self.buf(name + "/save_mean", (cout_pad,), torch.float32)
        srstd = self.buf(name + "/save_rstd", (cout_pad,), torch.float32)
        count = float(x.B * valid) * (sync.world if sync is not None else 1)
        if bn is None:
            scale.fill_(1.0)
            shift.copy_(st.view(bias))
            smean.zero_()
            srstd.fill_(1.0)
        elif training:
            fused_finalize = epi_stats and not defer_apply      # finalisation folded into the apply kernel below
            if not fused_finalize:
                self.call(self.lib.xv_bn_finalize_train, L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(st.view(bias)),
                          C.c_float(count), L.ptr(st.view(bn[0])), L.ptr(st.view(bn[1])), L.ptr(st.view(bn[2])),
                          L.ptr(st.view(bn[3])), C.c_float(momentum), C.c_float(BN_EPS), int(unbiased_moving_var),
                          L.ptr(scale), L.ptr(shift), L.ptr(smean), L.ptr(srstd), cout_pad, L.stream_ptr())
        else:
            self.call(self.lib.xv_bn_finalize_infer, L.ptr(st.view(bn[0])), L.ptr(st.view(bn[1])), L.ptr(st.view(bn[2])),
                      L.ptr(st.view(bn[3])), L.ptr(st.view(bias)), C.c_float(BN_EPS), L.ptr(scale), L.ptr(shift), cout_pad,
                      L.stream_ptr())
        alpha_t = None if alpha is None else st.view(alpha)
        lp = L.ptr(lengths)
        ya = FrameAct(y, x.B, x.T, valid, cout, lengths, name + "/y", ld=cout_pad)
        ya.affine = (scale, shift)
        ya.bias = st.view(bias)          # endpoints["tdnnN_conv"].dense() adds it back (the reference's tensor carries it)
        aa = FrameAct(None, x.B, x.T, valid, cout, lengths, name + "/a", ld=cout_pad)

        def apply_now():
            a = self.buf(name + "/a", (R, cout_pad), torch.bfloat16)
            if bn is not None and training and epi_stats and not defer_apply:
                self.call(self.lib.xv_bn_train_apply, L.ptr(y), L.ptr(a), L.ptr(stats[0]), L.ptr(stats[1]),
                          L.ptr(st.view(bias)), C.c_float(count), L.ptr(st.view(bn[0])), L.ptr(st.view(bn[1])),
                          L.ptr(st.view(bn[2])), L.ptr(st.view(bn[3])), C.c_float(momentum), C.c_float(BN_EPS),
                          int(unbiased_moving_var), L.ptr(scale), L.ptr(shift), L.ptr(smean), L.ptr(srstd), L.ptr(alpha_t),
                          act, C.c_int64(R), cout_pad, C.c_int64(cout_pad), x.T, valid, lp, L.stream_ptr())
            else:
                self.call(self.lib.xv_bn_act_apply, L.ptr(y), L.ptr(a), L.ptr(scale), L.ptr(shift), L.ptr(alpha_t), act,
                          C.c_int64(R), cout_pad, C.c_int64(cout_pad), x.T, valid, lp, L.stream_ptr())
            aa.data = a
        aa._materialize = apply_now
        if defer_apply:     # the consumer (statistics pooling) applies BN + activation on the fly
            aa.lazy = (y, scale, shift, alpha_t, act, smean, srstd)
        elif fold_infer:
            a = self.buf(name + "/a", (R, cout_pad), torch.bfloat16)
            neg = {L.ACT_NONE: 1.0, L.ACT_RELU: 0.0, L.ACT_LRELU: 0.2}[act]
            self.gemm(a_op, L.operand(W, True), R, cout_pad, K, a, epilogue=L.EPI_BF16, affine=(scale, shift, neg),
                      seg_len=(x.T if lengths is None else 0), seg_valid=valid)
            aa.data = a
            ya.data = None

            def materialize_y():            # pre-BN endpoint on demand (e.g. embedding_node = "tdnn5_dense")
                yb = self.buf(name + "/y", (R, cout_pad), torch.bfloat16)
                run_plain_gemm(yb)
                ya.data = yb
            ya._materialize = materialize_y
        else:
            apply_now()

        if training and bn is not None and alpha is None and act in (L.ACT_NONE, L.ACT_RELU, L.ACT_LRELU) and lengths is None:
            neg = {L.ACT_NONE: 1.0, L.ACT_RELU: 0.0, L.ACT_LRELU: 0.2}[act]
            aa.bn_bwd = (y, scale, shift, smean, srstd, neg, st.grad(bn[0]), st.grad(bn[1]))
        if training:
            def bwd():
                if aa.grad is None and aa.pool_grad is None:
                    return
                fused = aa.grad is None
                pooled, dpooled = aa.pool_grad if fused else (None, None)
                pool_args = (L.ptr(pooled), L.ptr(dpooled), cout_pad, cout)
                if bn is not None:
                    dgamma, dbeta = st.grad(bn[0]), st.grad(bn[1])
                else:
                    dgamma = self.buf(name + "/dgamma0", (cout_pad,), torch.float32, zero=True)
                    dbeta = st.grad(bias)
                dalpha = None if alpha is None else st.grad(alpha)
                if aa.bn_reduced:
                    pass        # the consumer's dgrad epilogue already accumulated dgamma / dbeta
                elif fused and aa.pool_sums is not None and bn is not None:
                    self.call(self.lib.xv_pool_bn_bwd_reduce, L.ptr(pooled), L.ptr(dpooled), L.ptr(aa.pool_sums), x.B, valid,
                              lp, cout, cout_pad, L.ptr(dgamma), L.ptr(dbeta), L.stream_ptr())
                else:
                    self.call(self.lib.xv_bn_act_bwd_reduce, L.ptr(y), L.ptr(aa.grad), L.ptr(scale), L.ptr(shift),
                              L.ptr(smean), L.ptr(srstd), L.ptr(alpha_t), act, C.c_int64(R), cout_pad, C.c_int64(cout_pad),
                              x.T, valid, lp, L.ptr(dgamma), L.ptr(dbeta), L.ptr(dalpha), *pool_args, L.stream_ptr())
                dy = self.buf(name + "/dy", (R, cout_pad), torch.bfloat16)
                if sync is not None:
                    # SyncBN backward: dgamma / dbeta over the global batch enter dy; they live in the flat gradient buffer,
                    # whose final all-reduce would sum them a second time -> pre-scaled by 1/N after use (below)
                    self.collective(lambda: (sync.all_reduce_sum_(dgamma), sync.all_reduce_sum_(dbeta)))
                if bn is not None:
                    dg_used, db_used = dgamma, dbeta
                else:       # no BN: dy = g; the reductions above were the bias gradient
                    dg_used = db_used = self.buf(name + "/zeros", (cout_pad,), torch.float32, zero=True)
                self.call(self.lib.xv_bn_act_bwd_apply, L.ptr(y), L.ptr(aa.grad), L.ptr(dy), L.ptr(scale), L.ptr(shift),
                          L.ptr(smean), L.ptr(srstd), L.ptr(dg_used), L.ptr(db_used), C.c_float(count), L.ptr(alpha_t),
                          act, C.c_int64(R), cout_pad, C.c_int64(cout_pad), x.T, valid, lp, *pool_args, L.stream_ptr())
                if sync is not None:
                    dgamma.mul_(1.0 / sync.world)
                    dbeta.mul_(1.0 / sync.world)
                # wgrad: dW[(j,c), n] = sum_r X[r+j, c] dY[r, n].  Nothing downstream of it until the optimizer, so it runs
                # on a side stream: its CTAs fill the SMs that the dgrad's last (partial) wave leaves idle and overlap the
                # memory-bound BN kernels of the next layer.
                gw = st.grad(kernel)
                with self.on_side_stream():
                    self.gemm(L.operand(xd, True, div=(x.ld if k > 1 else 0), tap_rows=(1 if k > 1 else 0)),
                              L.operand(dy, True), K, cout_pad, R, gw, epilogue=L.EPI_F32,
                              splits=self.splits_for(K, cout_pad, R))
                if x.needs_grad:
                    # dgrad: dX[q, c] = sum_j dY[q-j, :] W_j[c, :]^T
                    # a shared activation (e.g. tdnn4_relu feeding tdnn5 and the attention key net) gets the sum
                    fan_in = x.grad is not None
                    dx = x.grad if fan_in else self.buf(x.name + "/grad", (R, x.ld), torch.bfloat16)
                    # Sole consumer of x and a K loop long enough to hide it: this dgrad's epilogue also forms the
                    # BN-backward reductions (dgamma, dbeta) of the layer that produced x from the dX tile it holds.
                    fuse_bn = (self.fuse_bn_bwd and x.bn_bwd is not None and x.consumers == 1 and not fan_in
                               and k * cout_pad >= self.fuse_bn_bwd_min_k and x.ld % 32 == 0)
                    bnb = x.bn_bwd[:6] if fuse_bn else None
                    self.gemm(L.operand(dy, False, div=(cout_pad if k > 1 else 0), tap_rows=(-1 if k > 1 else 0)),
                              L.operand(W, False, div=(cout_pad if k > 1 else 0), tap_rows=(x.ld if k > 1 else 0)),
                              R, x.ld, k * cout_pad, dx, epilogue=L.EPI_BF16, accumulate=fan_in, bn_bwd=bnb,
                              col_sum=x.bn_bwd[7] if fuse_bn else None, col_sumsq=x.bn_bwd[6] if fuse_bn else None)
                    x.bn_reduced = fuse_bn
                    x.grad = dx
            self.tape.append(bwd)
            aa.needs_grad = True
        return ya, aa

    def stats_pool(self, x, training, ragged=None):
        cpad = x.ld
        x.consumers += 1
        if ragged is not None:
            # extraction: ``x`` is one flat row space holding the concatenated utterances; (starts, lengths) int32 device
            # tensors give each utterance's rows in the POOLED domain (see xv_stats_pool_ragged)
            assert not training, "the ragged layout is an inference-only path (training draws one length per batch)"
            starts, plen = ragged
            n = int(starts.shape[0])
            out = self.buf("pool/out", (n, 2 * cpad), torch.float32)
            out3 = self.buf("pool/out3", (n, 6 * cpad), torch.bfloat16)
            if x.lazy is not None:
                y_, scale_, shift_, alpha_, act_, _, _ = x.lazy
                src, sc, sh, al, ac = y_, scale_, shift_, alpha_, act_
            else:
                src, sc, sh, al, ac = x.materialize(), None, None, None, 0
            self.call(self.lib.xv_stats_pool_ragged, L.ptr(src), L.ptr(out), L.ptr(out3), n, L.ptr(starts), L.ptr(plen), x.C,
                      cpad, C.c_int64(cpad), L.ptr(sc), L.ptr(sh), L.ptr(al), ac, L.stream_ptr())
            return UttAct(out, out3, "pool", (x.C, cpad))
        out = self.buf("pool/out", (x.B, 2 * cpad), torch.float32)
        out3 = self.buf("pool/out3", (x.B, 6 * cpad), torch.bfloat16)
        if x.lazy is not None:      # fused tdnn5 BN + activation: pool act(y*scale + shift) straight from y
            y_, scale_, shift_, alpha_, act_, smean_, srstd_ = x.lazy
            sums = None
            if training and act_ != L.ACT_PRELU:
                # per-(segment, channel) sums that give the BN dgamma / dbeta without another pass over y (78 MB)
                sums = self.buf("pool/bwd_sums", (x.B, 4, cpad), torch.float32)
            x.pool_sums = sums
            self.call(self.lib.xv_stats_pool_fwd, L.ptr(y_), L.ptr(out), L.ptr(out3), x.B, x.T, x.valid, x.lengths_ptr(),
                      x.C, cpad, C.c_int64(cpad), L.ptr(scale_), L.ptr(shift_), L.ptr(alpha_), act_,
                      L.ptr(smean_ if sums is not None else None), L.ptr(srstd_ if sums is not None else None),
                      L.ptr(sums), L.stream_ptr())
        else:
            self.call(self.lib.xv_stats_pool_fwd, L.ptr(x.data), L.ptr(out), L.ptr(out3), x.B, x.T, x.valid,
                      x.lengths_ptr(), x.C, cpad, C.c_int64(cpad), L.ptr(None), L.ptr(None), L.ptr(None), 0,
                      L.ptr(None), L.ptr(None), L.ptr(None), L.stream_ptr())
        u = UttAct(out, out3, "pool", (x.C, cpad))
        if training:
            def bwd():
                if u.grad is None:
                    return
                if x.lazy is not None:      # the tdnn5 BN backward evaluates the pooling gradient on the fly
                    x.pool_grad = (out, u.grad)
                    return
                dx = self.buf(x.name + "/grad", (x.B * x.T, cpad), torch.bfloat16)
                self.call(self.lib.xv_stats_pool_bwd, L.ptr(x.data), L.ptr(out), L.ptr(u.grad), L.ptr(dx), x.B, x.T,
                          x.valid, x.lengths_ptr(), x.C, cpad, C.c_int64(cpad), L.stream_ptr())
                x.grad = dx
            self.tape.append(bwd)
            u.needs_grad = True
        return u

    def att_pool(self, key, value, query, H, split_key, use_scale, penalty_coef, training):
        """Attentive statistics pooling (model/pooling.py:120-189): scores = <key, query_h> (* rsqrt(dk)), softmax over the
        valid frames, weighted mean / stddev of the value per head, optional multi-head penalty.
        key / value: FrameAct (materialised bf16); returns (UttAct [B, 2*dv], weights f32 [B, H, T] view)."""
        st = self.store
        assert key.B == value.B and key.T == value.T and key.valid == value.valid, "key and value must have the same length"
        B, T, valid = value.B, value.T, value.valid
        lengths = value.lengths
        lp = L.ptr(lengths)
        kd, vd = key.materialize(), value.materialize()
        key.consumers += 1
        value.consumers += 1
        ldk, cpad = key.ld, value.ld
        q = st.view(query)                               # f32 [H, dq]
        dq = q.shape[1]
        scale = (1.0 / math.sqrt(dq)) if use_scale else 1.0
        s = L.stream_ptr
        qpad = self.buf("att/qpad", (H, ldk), torch.float32)
        self.call(self.lib.xv_att_expand_query, L.ptr(q), L.ptr(qpad), H, dq, ldk, int(split_key), s())
        w = self.buf("att/weights", (B, H, T), torch.float32)
        self.call(self.lib.xv_att_scores_fwd, L.ptr(kd), L.ptr(qpad), L.ptr(w), B, T, valid, lp, H, ldk,
                  C.c_float(scale), s())
        self.call(self.lib.xv_att_softmax_fwd, L.ptr(w), L.ptr(w), B, H, T, valid, lp, s())
        out = self.buf("pool/out", (B, 2 * cpad), torch.float32)
        out3 = self.buf("pool/out3", (B, 6 * cpad), torch.bfloat16)
        self.call(self.lib.xv_att_pool_fwd, L.ptr(vd), L.ptr(w), L.ptr(out), L.ptr(out3), B, H, T, valid, lp, value.C, cpad,
                  C.c_int64(cpad), s())
        gram = None
        if penalty_coef != 0.0 and self.inv_global_batch is not None:
            # data parallel: the kernels divide by the rank-local B, the penalty (pooling.py:185-188) by the GLOBAL batch
            # -- gradients and the scalar are summed over the ranks afterwards, like the head's 1/(N*B)
            penalty_coef = penalty_coef * (B * self.inv_global_batch)
        if penalty_coef != 0.0:
            gram = self.buf("att/gram", (B, H, H), torch.float32)
            self.call(self.lib.xv_att_penalty_fwd, L.ptr(w), L.ptr(gram), L.ptr(self.scalars[3:4]), B, H, T, valid, lp,
                      C.c_float(penalty_coef), s())
        u = UttAct(out, out3, "pool", (value.C, cpad))
        if training:
            def bwd():
                if u.grad is None:
                    return
                dw = self.buf("att/dweights", (B, H, T), torch.float32)
                acc_v = value.grad is not None
                dv = value.grad if acc_v else self.buf(value.name + "/grad", (B * T, cpad), torch.bfloat16)
                self.call(self.lib.xv_att_pool_bwd, L.ptr(vd), L.ptr(w), L.ptr(out), L.ptr(u.grad), L.ptr(dv), L.ptr(dw), B,
                          H, T, valid, lp, value.C, cpad, C.c_int64(cpad), int(acc_v), s())
                value.grad = dv
                self.call(self.lib.xv_att_softmax_bwd, L.ptr(w), L.ptr(dw), L.ptr(gram), B, H, T, valid, lp,
                          C.c_float(penalty_coef), C.c_float(scale), s())
                dqpad = self.buf("att/dqpad", (H, ldk), torch.float32, zero=True)
                acc_k = key.grad is not None
                dk = key.grad if acc_k else self.buf(key.name + "/grad", (B * T, ldk), torch.bfloat16)
                self.call(self.lib.xv_att_scores_bwd, L.ptr(kd), L.ptr(qpad), L.ptr(dw), L.ptr(dk), L.ptr(dqpad), B, T,
                          valid, lp, H, ldk, int(acc_k), s())
                key.grad = dk
                self.call(self.lib.xv_att_fold_query_grad, L.ptr(dqpad), L.ptr(st.grad(query)), H, dq, ldk,
                          int(split_key), s())
            self.tape.append(bwd)
            u.needs_grad = True
        return u, w[:, :, :valid]

    def vlad_pool(self, logits, value, centers, K, G, final_norm, training):
        """NetVLAD / GhostVLAD aggregation (model/pooling.py:249-276): posteriors = softmax of the cluster logits over the
        K real + G ghost clusters, residuals to the centres summed over the valid frames, intra-cluster (and optionally
        final) L2 normalisation.  logits / value: FrameAct; returns (UttAct [B, K*dv], posteriors f32 [B, T, K+G] view)."""
        st = self.store
        assert logits.B == value.B and logits.T == value.T and logits.valid == value.valid, \
            "key and value must have the same length"
        B, T, valid = value.B, value.T, value.valid
        lp = L.ptr(value.lengths)
        ld_, vd = logits.materialize(), value.materialize()
        logits.consumers += 1
        value.consumers += 1
        KG = K + G
        ldl, cpad, c_real = logits.ld, value.ld, value.C
        cen = st.view(centers)                          # f32 [KG, cpad]
        ldc = cen.shape[1]
        s = L.stream_ptr
        post = self.buf("vlad/post", (B, T, KG), torch.float32)
        self.call(self.lib.xv_vlad_post_fwd, L.ptr(ld_), L.ptr(post), B, T, valid, lp, KG, ldl, s())
        res = self.buf("vlad/res", (B, K, cpad), torch.float32)
        mass = self.buf("vlad/mass", (B, K), torch.float32)
        sumsq = self.buf("vlad/sumsq", (B, K), torch.float32)
        out = self.buf("pool/out", (B, K * cpad), torch.float32)
        out3 = self.buf("pool/out3", (B, 3 * K * cpad), torch.bfloat16)
        self.call(self.lib.xv_vlad_pool_fwd, L.ptr(vd), L.ptr(post), L.ptr(cen), L.ptr(res), L.ptr(mass), L.ptr(sumsq),
                  L.ptr(out), L.ptr(out3), B, T, valid, lp, K, KG, c_real, cpad, C.c_int64(cpad), ldc, int(final_norm), s())
        u = UttAct(out, out3, "pool", None)
        u.dense = lambda: out.view(B, K, cpad)[:, :, :c_real].reshape(B, K * c_real)
        if training:
            def bwd():
                if u.grad is None:
                    return
                gres = self.buf("vlad/gres", (B, K, cpad), torch.float32)
                gc = self.buf("vlad/gc", (B, K), torch.float32)
                dl = self.buf(logits.name + "/grad", (B * T, ldl), torch.bfloat16)
                acc_v = value.grad is not None
                dv = value.grad if acc_v else self.buf(value.name + "/grad", (B * T, cpad), torch.bfloat16)
                self.call(self.lib.xv_vlad_pool_bwd, L.ptr(vd), L.ptr(post), L.ptr(cen), L.ptr(mass), L.ptr(sumsq), L.ptr(out),
                          L.ptr(u.grad), L.ptr(gres), L.ptr(gc), L.ptr(dl), L.ptr(dv), L.ptr(st.grad(centers)), B, T, valid, lp,
                          K, KG, c_real, cpad, C.c_int64(cpad), ldl, ldc, int(final_norm), int(acc_v), s())
                assert logits.grad is None, "the cluster logits have one consumer"
                logits.grad = dl
                value.grad = dv
            self.tape.append(bwd)
            u.needs_grad = True
        return u, post[:, :valid, :]

    def utt_bn_act(self, u, bn, name, training, act=L.ACT_RELU, alpha=None, momentum=0.99):
        """BN over the batch + activation on an utterance-level tensor without a preceding dense layer
        (att_post_bn / att_post_relu, model/pooling.py:175-182).  Returns (BN output UttAct, activation UttAct)."""
        st = self.store
        B, Cn = u.data.shape
        mode = 1 if training else 2
        if training and self.sync_bn is not None and self.sync_bn.world > 1:
            raise NotImplementedError("sync_bn does not cover att_apply_nonlinear's post-pooling batch-norm")
        a = self.buf(name + "/a", (B, Cn), torch.float32)
        a3 = self.buf(name + "/a3", (B, 3 * Cn), torch.bfloat16)
        bn_out = self.buf(name + "/bn", (B, Cn), torch.float32)
        smean = self.buf(name + "/save_mean", (Cn,), torch.float32)
        srstd = self.buf(name + "/save_rstd", (Cn,), torch.float32)
        g = lambda i: L.ptr(st.view(bn[i]))
        alpha_t = None if alpha is None else st.view(alpha)
        self.call(self.lib.xv_bn_rows_fwd, L.ptr(u.data), B, Cn, mode, g(0), g(1), g(2), g(3), C.c_float(momentum),
                  C.c_float(BN_EPS), L.ptr(alpha_t), act, L.ptr(bn_out), L.ptr(a), L.ptr(a3), 3, L.ptr(smean),
                  L.ptr(srstd), L.stream_ptr())
        au = UttAct(a, a3, name + "/a", u.col_map)
        bu = UttAct(bn_out, None, name + "/bn", u.col_map)
        if training:
            def bwd():
                if au.grad is None:
                    return
                du = self.buf(u.name + "/grad_bn", (B, Cn), torch.float32)
                self.call(self.lib.xv_bn_rows_bwd, L.ptr(u.data), L.ptr(au.grad), B, Cn, mode, g(0), g(1), L.ptr(smean),
                          L.ptr(srstd), L.ptr(alpha_t), act, L.ptr(du), L.ptr(None), L.ptr(st.grad(bn[0])),
                          L.ptr(st.grad(bn[1])), L.ptr(None if alpha is None else st.grad(alpha)), L.ptr(None),
                          L.stream_ptr())
                u.grad = du
            self.tape.append(bwd)
            au.needs_grad = True
        return bu, au

    # ---- utterance-level ops ----------------------------------------------------------------------
    def utt_affine(self, u, kernel, bias, name, training, bn=None, act=L.ACT_RELU, alpha=None, momentum=0.99):
        """dense -> [BN over the batch] -> activation on fp32 [B, C] (tdnn6 / tdnn7, model/tdnn.py:147-189).
        Returns (pre-BN UttAct, BN-output tensor or None, post-activation UttAct)."""
        st = self.store
        W3 = st.shadow_view(kernel)                    # [3K, cout] = [hi; lo; hi]
        K = W3.shape[0] // 3
        cout = W3.shape[1]
        B = u.data.shape[0]
        assert u.split is not None and u.split.shape[1] == 3 * K, "%s: input split %s vs K %d" % (name, tuple(u.split.shape), K)
        splits = self.splits_for(B, cout, 3 * K)
        y = self.buf(name + "/y", (B, cout), torch.float32, zero=(splits > 1))
        self.gemm(L.operand(u.split, False), L.operand(W3, True), B, cout, 3 * K, y, epilogue=L.EPI_F32, splits=splits,
                  bias=st.view(bias))
        mode = 0 if bn is None else (1 if training else 2)
        sync = self.sync_bn if (mode == 1 and self.sync_bn is not None and self.sync_bn.world > 1) else None
        # SyncBN on [B, C] tensors: the rows of every rank are all-gathered (a few hundred KB) and every rank normalises
        # all N*B rows, so the batch statistics -- and later dgamma / dbeta -- are the global ones without a reduction
        Bn = B * sync.world if sync is not None else B
        r0 = sync.rank * B if sync is not None else 0
        y_bn = y
        if sync is not None:
            y_bn = self.buf(name + "/y_all", (Bn, cout), torch.float32)
            self.collective(lambda: sync.all_gather(y_bn, y))
        a_f = self.buf(name + "/a", (Bn, cout), torch.float32)
        a3_f = self.buf(name + "/a3", (Bn, 3 * cout), torch.bfloat16)
        bn_f = self.buf(name + "/bn", (Bn, cout), torch.float32) if bn is not None else None
        smean = self.buf(name + "/save_mean", (cout,), torch.float32)
        srstd = self.buf(name + "/save_rstd", (cout,), torch.float32)
        g = lambda i: (L.ptr(st.view(bn[i])) if bn is not None else L.ptr(None))
        alpha_t = None if alpha is None else st.view(alpha)
        self.call(self.lib.xv_bn_rows_fwd, L.ptr(y_bn), Bn, cout, mode, g(0), g(1), g(2), g(3), C.c_float(momentum),
                  C.c_float(BN_EPS), L.ptr(alpha_t), act, L.ptr(bn_f), L.ptr(a_f), L.ptr(a3_f), 3, L.ptr(smean),
                  L.ptr(srstd), L.stream_ptr())
        a, a3 = a_f[r0:r0 + B], a3_f[r0:r0 + B]
        bn_out = None if bn_f is None else bn_f[r0:r0 + B]
        yu = UttAct(y, None, name + "/y")
        au = UttAct(a, a3, name + "/a")
        if bn_out is not None:
            bn_out = UttAct(bn_out, None, name + "/bn")
        if training:
            def bwd():
                if au.grad is None and yu.grad is None:
                    return
                da = au.grad if au.grad is not None else self.buf(name + "/da0", (B, cout), torch.float32, zero=True)
                da_bn = da
                if sync is not None:
                    da_bn = self.buf(name + "/da_all", (Bn, cout), torch.float32)
                    da_c = da.contiguous()
                    self.collective(lambda: sync.all_gather(da_bn, da_c))
                dy_f = self.buf(name + "/dy", (Bn, cout), torch.float32)
                dyb_f = self.buf(name + "/dyb", (Bn, cout), torch.bfloat16)
                gg = lambda i: (L.ptr(st.grad(bn[i])) if bn is not None else L.ptr(None))
                self.call(self.lib.xv_bn_rows_bwd, L.ptr(y_bn), L.ptr(da_bn), Bn, cout, mode, g(0), g(1), L.ptr(smean),
                          L.ptr(srstd), L.ptr(alpha_t), act, L.ptr(dy_f), L.ptr(dyb_f), gg(0), gg(1),
                          L.ptr(None if alpha is None else st.grad(alpha)), L.ptr(st.grad(bias)), L.stream_ptr())
                dy, dyb = dy_f[r0:r0 + B], dyb_f[r0:r0 + B]
                if sync is not None:     # every rank holds the GLOBAL dgamma / dbeta / dbias: undo the final all-reduce's sum
                    for t_ in ([st.grad(bn[0]), st.grad(bn[1])] if bn is not None else []) + [st.grad(bias)] + \
                              ([st.grad(alpha)] if alpha is not None else []):
                        t_.mul_(1.0 / sync.world)
                # wgrad: dW[kk, n] = sum_i u[i, kk] dy[i, n]   (A = bf16(u) MN-major window of the split copy)
                with self.on_side_stream(self.side_utt):
                    self.gemm(L.operand(u.split, True, cols=K), L.operand(dyb, True), K, cout, B, st.grad(kernel),
                              epilogue=L.EPI_F32)
                if u.needs_grad:
                    du = self.buf(u.name + "/grad", (B, K), torch.float32)
                    self.gemm(L.operand(dyb, False), L.operand(W3, False, rows=K), B, K, cout, du, epilogue=L.EPI_F32)
                    u.grad = du
            self.tape.append(bwd)
            au.needs_grad = True
            yu.needs_grad = True
        return yu, bn_out, au

    # ---- head -------------------------------------------------------------------------------------
    def prefetch_head_weights(self, kernel, normalize):
        """Start the head's weight preparation (column normalisation + bf16 [hi; lo; hi] split, 37 MB of traffic at 7200
        speakers) on the side stream at the beginning of the step: it depends on the parameters only, so it overlaps the
        trunk forward instead of sitting in the latency-bound utterance-level chain.  margin_head() joins it."""
        if not self.side_utt or self.head_shard is not None or kernel not in self.store:
            return
        st = self.store
        Wm = st.view(kernel)
        E, cpad = Wm.shape
        wn3 = self.buf("head/wn3", (3 * E, cpad), torch.bfloat16)
        inv_norm = self.buf("head/inv_norm", (cpad,), torch.float32)
        with self.on_side_stream(True):
            self.call(self.lib.xv_head_prep_weights, L.ptr(Wm), L.ptr(wn3), L.ptr(inv_norm), E, cpad, C.c_int64(cpad),
                      int(normalize), L.stream_ptr())
        self._head_prefetch = (kernel, int(normalize))

    def margin_head(self, u, labels, kernel, bias, head_type, num_outputs, training, margin=0.0, asoftmax_m=1,
                    scaling=0.0, want_logits=False, aux=None):
        """Fused normalise -> cosine GEMM -> margin -> online log-sum-exp (model/loss.py heads + l2_scaling).
        ``u`` is the network output before feature_norm; returns (loss scalar tensor view, logits or None, x)."""
        st = self.store
        Wm = st.view(kernel)                           # fp32 [E, Cpad]
        E, cpad = Wm.shape
        Cn = num_outputs
        B = u.data.shape[0]
        normalize = 0 if head_type == L.HEAD_SOFTMAX else 1
        wn3 = self.buf("head/wn3", (3 * E, cpad), torch.bfloat16)
        inv_norm = self.buf("head/inv_norm", (cpad,), torch.float32)
        if self._head_prefetch == (kernel, normalize):
            self.join_side_stream()          # prepared on the side stream while the trunk forward was running
            self._head_prefetch = None
        else:
            self.call(self.lib.xv_head_prep_weights, L.ptr(Wm), L.ptr(wn3), L.ptr(inv_norm), E, cpad, C.c_int64(cpad),
                      normalize, L.stream_ptr())
        x = self.buf("head/x", (B, E), torch.float32)
        x3 = self.buf("head/x3", (B, 3 * E), torch.bfloat16)
        xnorm = self.buf("head/xnorm", (B,), torch.float32)
        urinv = self.buf("head/urinv", (B,), torch.float32)
        self.call(self.lib.xv_head_prep_features, L.ptr(u.data), C.c_float(scaling), L.ptr(x), L.ptr(x3), L.ptr(xnorm),
                  L.ptr(urinv), B, E, L.stream_ptr())
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        nblk = 2 * ((Cn + 127) // 128)      # one (max, sum) partial per 64-column half tile (two epilogue warps per row)
        pmax = self.buf("head/pmax", (nblk, B), torch.float32)
        psum = self.buf("head/psum", (nblk, B), torch.float32)
        tgt = self.buf("head/target", (B,), torch.float32)
        lse = self.buf("head/lse", (B,), torch.float32)
        gnorm = self.buf("head/gnorm", (B,), torch.float32, zero=True)
        logits = self.buf("head/logits", (B, cpad), torch.float32) if want_logits else None
        inv_batch = self.inv_global_batch if self.inv_global_batch is not None else 1.0 / B
        h = L.HeadArgs()
        h.type, h.asoftmax_m, h.margin = head_type, asoftmax_m, margin
        h.cos_m, h.sin_m, h.threshold = math.cos(margin), math.sin(margin), math.cos(math.pi - margin)
        h.sched = self.sched.data_ptr()
        h.labels, h.xnorm = labels.data_ptr(), xnorm.data_ptr()
        h.part_max, h.part_sum, h.target_logit = pmax.data_ptr(), psum.data_ptr(), tgt.data_ptr()
        h.logits_out = 0 if logits is None else logits.data_ptr()
        h.lse, h.inv_batch, h.gnorm = lse.data_ptr(), inv_batch, gnorm.data_ptr()
        bias_t = None if bias is None else st.view(bias)
        dummy = self.buf("head/dummy", (8,), torch.float32)
        self.gemm(L.operand(x3, False), L.operand(wn3, True, cols=Cn), B, Cn, 3 * E, dummy, epilogue=L.EPI_HEAD_FWD,
                  bias=bias_t, head=h, ldc=cpad)
        self.call(self.lib.xv_head_combine, L.ptr(pmax), L.ptr(psum), L.ptr(tgt), nblk, B, C.c_float(inv_batch),
                  L.ptr(lse), L.ptr(None), L.ptr(self.scalars[0:1]), L.stream_ptr())
        # auxiliary losses (model/loss.py:985-1037); they add to the head loss like the reference's ``loss += loss_aux``
        aux = aux or {}
        ring = aux.get("ring")            # (variable name of r, lambda)
        mhe = aux.get("mhe")              # lambda
        if ring is not None:
            self.call(self.lib.xv_ring_loss, L.ptr(xnorm), L.ptr(st.view(ring[0])), B, C.c_float(ring[1] * inv_batch),
                      L.ptr(self.scalars[0:1]), L.ptr(None), L.ptr(None), L.stream_ptr())
        if mhe is not None:
            if not normalize:
                raise NotImplementedError("mhe_loss needs a head with normalised weights (asoftmax / AM / AAM)")
            # data parallel: every replica evaluates the energy on its own labels; the summed scalars / gradients are the
            # mean over replicas (the term is not linear in the batch, like per-replica batch-norm)
            rep = 1.0 / max(1, int(round(1.0 / (inv_batch * B))))
            mt = self.buf("head/mhe_t", (E,), torch.float32)
            mS = self.buf("head/mhe_S", (E,), torch.float32)
            mk = self.buf("head/mhe_kappa", (8,), torch.float32)
            mh = self.buf("head/mhe_hist", (cpad,), torch.float32, zero=True)
            self.call(self.lib.xv_mhe_forward, L.ptr(Wm), L.ptr(inv_norm), L.ptr(labels), B, E, Cn, C.c_int64(cpad),
                      C.c_float(mhe), C.c_float(rep), L.ptr(mt), L.ptr(mS), L.ptr(mh), L.ptr(mk), L.ptr(self.scalars[0:1]),
                      L.stream_ptr())
            self.launches += 1
        if training:
            def bwd():
                d = self.buf("head/d", (B, cpad), torch.bfloat16)
                self.gemm(L.operand(x3, False), L.operand(wn3, True, cols=Cn), B, Cn, 3 * E, d, epilogue=L.EPI_HEAD_BWD,
                          bias=bias_t, head=h, col_sum=(st.grad(bias) if bias is not None else None))
                if ring is not None:      # d ring / d||x_i|| joins the margin term's gnorm; d ring / dr is a scalar
                    self.call(self.lib.xv_ring_loss, L.ptr(xnorm), L.ptr(st.view(ring[0])), B, C.c_float(ring[1] * inv_batch),
                              L.ptr(None), L.ptr(gnorm), L.ptr(st.grad(ring[0])), L.stream_ptr())
                # dWn[e, c] = sum_i x[i, e] d[i, c]
                gw = st.grad(kernel)
                with self.on_side_stream(self.side_utt):
                    self.gemm(L.operand(x3, True, cols=E), L.operand(d, True, cols=Cn), E, Cn, B, gw, epilogue=L.EPI_F32)
                    if mhe is not None:
                        self.call(self.lib.xv_mhe_backward, L.ptr(gw), L.ptr(mt), L.ptr(mS), L.ptr(mh), L.ptr(mk), E, Cn,
                                  C.c_int64(cpad), L.stream_ptr())
                    if normalize:
                        self.call(self.lib.xv_head_finish_dw, L.ptr(gw), L.ptr(Wm), L.ptr(inv_norm), E, cpad, L.stream_ptr())
                # dx[i, e] = sum_c d[i, c] wn[e, c]
                dxg = self.buf("head/dxg", (B, E), torch.float32, zero=True)
                sp = self.splits_for(B, E, Cn)
                self.gemm(L.operand(d, False, cols=Cn), L.operand(wn3, False, rows=E, cols=Cn), B, E, Cn, dxg,
                          epilogue=L.EPI_F32, splits=sp)
                du = self.buf(u.name + "/grad", (B, E), torch.float32)
                use_margin = head_type != L.HEAD_SOFTMAX or ring is not None
                self.call(self.lib.xv_head_finish_dx, L.ptr(dxg), L.ptr(gnorm if use_margin else None), L.ptr(x),
                          L.ptr(xnorm), L.ptr(u.data), L.ptr(urinv), C.c_float(scaling), L.ptr(du), B, E, L.stream_ptr())
                u.grad = du
            self.tape.append(bwd)
        return self.scalars[0], logits, x

    def metric_loss(self, u, labels, kind, training, scaling=0.0, margin=0.0, squared=False, angular_kind=0, hard=False,
                    speakers=0, segments=0):
        """Pairwise metric-learning losses on the embeddings (model/loss.py:358-705): kind = "semihard" | "angular" | "e2e_valid".
        One fp32 Gram matrix of the (optionally l2-scaled) embeddings, a mining kernel per anchor row, and -- in training --
        dLoss/dx = coef x + diag o x through a second fp32 GEMM (csrc/xv_metric.cu).  Data parallel: the mining is per
        replica (triplets never cross ranks, like per-replica batch-norm); the summed scalars / gradients are replica means."""
        B, E = u.data.shape
        s = L.stream_ptr
        x = self.buf("metric/x", (B, E), torch.float32)
        x3 = self.buf("metric/x3", (B, 3 * E), torch.bfloat16)
        xnorm = self.buf("metric/xnorm", (B,), torch.float32)
        urinv = self.buf("metric/urinv", (B,), torch.float32)
        self.call(self.lib.xv_head_prep_features, L.ptr(u.data), C.c_float(scaling), L.ptr(x), L.ptr(x3), L.ptr(xnorm),
                  L.ptr(urinv), B, E, s())
        rep = 1.0 if self.inv_global_batch is None else 1.0 / max(1, int(round(1.0 / (self.inv_global_batch * B))))
        loss = self.scalars[0:1]
        if kind == "e2e_valid":
            assert speakers * segments == B, "e2e_valid_loss: the batch must hold num_valid_speakers_per_batch x " \
                                             "num_valid_segments_per_speaker speaker-ordered rows"
            work = self.buf("metric/e2e_work", ((B + speakers) * E + speakers,), torch.float32)
            self.call(self.lib.xv_e2e_valid_loss, L.ptr(x), speakers, segments, E, C.c_int64(E), C.c_float(rep), L.ptr(loss),
                      L.ptr(work), s())
            return self.scalars[0], x
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        gram = self.buf("metric/gram", (B, B), torch.float32)
        coef = self.buf("metric/coef", (B, B), torch.float32)
        diag = self.buf("metric/diag", (B,), torch.float32)
        work = self.buf("metric/work", (2 * B * B + 8,), torch.float32)
        self.call(self.lib.xv_gram_f32, L.ptr(x), L.ptr(gram), B, E, C.c_int64(E), s())
        if kind == "semihard":
            self.call(self.lib.xv_semihard_triplet, L.ptr(gram), L.ptr(labels), B, C.c_float(margin), int(squared),
                      C.c_float(rep), L.ptr(loss), L.ptr(coef), L.ptr(diag), L.ptr(work), s())
            self.launches += 3
        elif kind == "angular":
            self.call(self.lib.xv_angular_triplet, L.ptr(gram), L.ptr(labels), B, int(angular_kind), C.c_float(margin),
                      int(hard), C.c_float(rep), L.ptr(loss), L.ptr(coef), L.ptr(diag), L.ptr(work), s())
            self.launches += 1 if hard else 2
        else:
            raise NotImplementedError("metric loss %s" % kind)
        if training:
            def bwd():
                dxg = self.buf("metric/dx", (B, E), torch.float32)
                self.call(self.lib.xv_pairwise_bwd, L.ptr(coef), L.ptr(diag), L.ptr(x), L.ptr(dxg), B, E, C.c_int64(E), s())
                du = self.buf(u.name + "/grad", (B, E), torch.float32)
                self.call(self.lib.xv_head_finish_dx, L.ptr(dxg), L.ptr(None), L.ptr(x), L.ptr(xnorm), L.ptr(u.data),
                          L.ptr(urinv), C.c_float(scaling), L.ptr(du), B, E, s())
                u.grad = du
            self.tape.append(bwd)
        return self.scalars[0], x

    def centre_triplet_head(self, u, labels, kernel, num_outputs, training, scaling=0.0, average=False, momentum=0.0,
                            margin=0.0, target_margin=0.0, topn=1, w_triplet=1.0, w_center=0.0, w_between=0.0):
        """Generalized angular triplet loss against class centres (model/loss.py:708-901, loss_compute "raw").  The cosine of
        every sample to every (normalised) centre is the head's tcgen05 GEMM on the [hi|hi|lo] split operands; a row kernel
        mines the hardest top-n centres and writes dLoss/dcos, which feeds the head's own dW / dx GEMMs and normalisation
        Jacobians.  ``average``: the centres are a moving average of the class members, updated before the cosines are
        taken (no gradient), instead of trainable parameters."""
        st = self.store
        Wm = st.view(kernel)                           # fp32 [E, cpad]
        E, cpad = Wm.shape
        Cn = num_outputs
        B = u.data.shape[0]
        s = L.stream_ptr
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        rep = 1.0 if self.inv_global_batch is None else 1.0 / max(1, int(round(1.0 / (self.inv_global_batch * B))))
        if average and rep != 1.0:
            raise NotImplementedError("triplet_center = average under data parallelism (every replica would keep its own centres)")
        x3 = self.buf("head/x3", (B, 3 * E), torch.bfloat16)
        xnorm = self.buf("head/xnorm", (B,), torch.float32)
        urinv = self.buf("head/urinv", (B,), torch.float32)
        if average and training:
            # the centres move towards the features the loss function was GIVEN (l2-scaled when feature_norm is on)
            feats = u.data
            if scaling > 0.0:
                feats = self.buf("head/xs", (B, E), torch.float32)
                self.call(self.lib.xv_head_prep_features, L.ptr(u.data), C.c_float(scaling), L.ptr(feats), L.ptr(x3), L.ptr(xnorm),
                          L.ptr(urinv), B, E, s())
            delta = self.buf("head/centre_delta", (B, E), torch.float32)
            if self._head_prefetch is not None:        # normalised from the OLD centres: discard
                self.join_side_stream()
                self._head_prefetch = None
            self.call(self.lib.xv_center_update, L.ptr(Wm), L.ptr(feats), L.ptr(labels), L.ptr(delta), B, E, C.c_int64(cpad),
                      C.c_float(1.0 - momentum), s())
            self.launches += 1
        wn3 = self.buf("head/wn3", (3 * E, cpad), torch.bfloat16)
        inv_norm = self.buf("head/inv_norm", (cpad,), torch.float32)
        if self._head_prefetch == (kernel, 1):
            self.join_side_stream()
            self._head_prefetch = None
        else:
            self.call(self.lib.xv_head_prep_weights, L.ptr(Wm), L.ptr(wn3), L.ptr(inv_norm), E, cpad, C.c_int64(cpad), 1, s())
        x = self.buf("head/x", (B, E), torch.float32)          # l2_normalize(features): scaling 1 on the raw output
        self.call(self.lib.xv_head_prep_features, L.ptr(u.data), C.c_float(1.0), L.ptr(x), L.ptr(x3), L.ptr(xnorm), L.ptr(urinv),
                  B, E, s())
        cosm = self.buf("head/cos", (B, cpad), torch.float32)
        self.gemm(L.operand(x3, False), L.operand(wn3, True, cols=Cn), B, Cn, 3 * E, cosm, epilogue=L.EPI_F32)
        trainable_w = training and not average
        d = self.buf("head/d", (B, cpad), torch.bfloat16) if training else None
        counters = self.buf("head/gt_counters", (8,), torch.float32)
        loss = self.scalars[0:1]
        self.call(self.lib.xv_center_triplet, L.ptr(cosm), L.ptr(labels), B, Cn, C.c_int64(cpad), C.c_float(margin),
                  C.c_float(target_margin), int(topn), C.c_float(w_triplet), C.c_float(w_center), C.c_float(rep), L.ptr(loss),
                  L.ptr(d), L.ptr(counters), s())
        tsum = self.buf("head/centre_sum", (E,), torch.float32)
        self.call(self.lib.xv_center_between, L.ptr(Wm), L.ptr(inv_norm), E, Cn, C.c_int64(cpad), C.c_float(w_between * rep),
                  L.ptr(tsum), L.ptr(loss), s())
        self.launches += 2
        if training:
            def bwd():
                if trainable_w:
                    gw = st.grad(kernel)
                    with self.on_side_stream(self.side_utt):
                        self.gemm(L.operand(x3, True, cols=E), L.operand(d, True, cols=Cn), E, Cn, B, gw, epilogue=L.EPI_F32)
                        if w_between != 0.0:
                            self.call(self.lib.xv_center_between_bwd, L.ptr(gw), L.ptr(Wm), L.ptr(inv_norm), L.ptr(tsum), E, Cn,
                                      C.c_int64(cpad), C.c_float(w_between * rep), s())
                        self.call(self.lib.xv_head_finish_dw, L.ptr(gw), L.ptr(Wm), L.ptr(inv_norm), E, cpad, s())
                dxg = self.buf("head/dxg", (B, E), torch.float32, zero=True)
                sp = self.splits_for(B, E, Cn)
                self.gemm(L.operand(d, False, cols=Cn), L.operand(wn3, False, rows=E, cols=Cn), B, E, Cn, dxg,
                          epilogue=L.EPI_F32, splits=sp)
                du = self.buf(u.name + "/grad", (B, E), torch.float32)
                self.call(self.lib.xv_head_finish_dx, L.ptr(dxg), L.ptr(None), L.ptr(x), L.ptr(xnorm), L.ptr(u.data), L.ptr(urinv),
                          C.c_float(1.0), L.ptr(du), B, E, s())
                u.grad = du
            self.tape.append(bwd)
        return self.scalars[0], cosm, x

    def margin_head_sharded(self, u, labels, kernel, bias, head_type, num_outputs, training, margin=0.0, asoftmax_m=1,
                            scaling=0.0):
        """Class-sharded variant of margin_head (north_star "Data parallelism"; SURVEY 8e 2'): this rank holds columns
        [lo, hi) of the speaker matrix.  Embeddings and labels of ALL ranks are all-gathered (R = N*B rows), every rank runs
        the fused cosine-GEMM / margin / online-LSE epilogue on its own columns, the per-row (max, sum, target) triples are
        exchanged once (the all-reduce(max) + all-reduce(sum) of a sharded softmax, evaluated on the gathered pairs), and the
        backward reduce-scatters dLoss/dx; dLoss/dW of a shard is complete locally, so the 512 x C gradient never crosses
        NVLink.  The loss scalar is the global-batch mean on every rank."""
        sh = self.head_shard
        st = self.store
        Wm = st.view(kernel)                           # fp32 [E, cpad_local]
        E, cpad = Wm.shape
        Cn = sh.n_local
        N, rk = sh.world, sh.rank
        B = u.data.shape[0]
        R = N * B
        normalize = 0 if head_type == L.HEAD_SOFTMAX else 1
        s = L.stream_ptr
        wn3 = self.buf("head/wn3", (3 * E, cpad), torch.bfloat16)
        inv_norm = self.buf("head/inv_norm", (cpad,), torch.float32)
        self.call(self.lib.xv_head_prep_weights, L.ptr(Wm), L.ptr(wn3), L.ptr(inv_norm), E, cpad, C.c_int64(cpad),
                  normalize, s())
        labels = labels.to(device=self.device, dtype=torch.int32).contiguous()
        u_loc, l_loc = u.data, labels                  # persistent buffers (named workspace / the step's static batch)
        assert u_loc.is_contiguous() and tuple(u_loc.shape) == (B, E)
        u_all = self.buf("head/u_all", (R, E), torch.float32)
        l_all = self.buf("head/labels_all", (R,), torch.int32)

        def gather_rows():
            sh.all_gather(u_all, u_loc)
            sh.all_gather(l_all, l_loc)
        self.collective(gather_rows)
        x = self.buf("head/x", (R, E), torch.float32)
        x3 = self.buf("head/x3", (R, 3 * E), torch.bfloat16)
        xnorm = self.buf("head/xnorm", (R,), torch.float32)
        urinv = self.buf("head/urinv", (R,), torch.float32)
        self.call(self.lib.xv_head_prep_features, L.ptr(u_all), C.c_float(scaling), L.ptr(x), L.ptr(x3), L.ptr(xnorm),
                  L.ptr(urinv), R, E, s())
        lab = self.buf("head/labels_shard", (R,), torch.int32)
        self.call(self.lib.xv_head_local_labels, L.ptr(l_all), sh.lo, Cn, L.ptr(lab), R, s())
        nblk = 2 * ((Cn + 127) // 128)
        pmax = self.buf("head/pmax", (nblk, R), torch.float32)
        psum = self.buf("head/psum", (nblk, R), torch.float32)
        tgt = self.buf("head/target", (R,), torch.float32, zero=True)       # written by the owning shard only
        lse = self.buf("head/lse", (R,), torch.float32)
        gnorm = self.buf("head/gnorm", (R,), torch.float32, zero=True)
        inv_batch = self.inv_global_batch if self.inv_global_batch is not None else 1.0 / R
        h = L.HeadArgs()
        h.type, h.asoftmax_m, h.margin = head_type, asoftmax_m, margin
        h.cos_m, h.sin_m, h.threshold = math.cos(margin), math.sin(margin), math.cos(math.pi - margin)
        h.sched = self.sched.data_ptr()
        h.labels, h.xnorm = lab.data_ptr(), xnorm.data_ptr()
        h.part_max, h.part_sum, h.target_logit = pmax.data_ptr(), psum.data_ptr(), tgt.data_ptr()
        h.logits_out = 0
        h.lse, h.inv_batch, h.gnorm = lse.data_ptr(), inv_batch, gnorm.data_ptr()
        bias_t = None if bias is None else st.view(bias)
        dummy = self.buf("head/dummy", (8,), torch.float32)
        self.gemm(L.operand(x3, False), L.operand(wn3, True, cols=Cn), R, Cn, 3 * E, dummy, epilogue=L.EPI_HEAD_FWD,
                  bias=bias_t, head=h, ldc=cpad)
        part = self.buf("head/shard_part", (3, R), torch.float32)
        parts = self.buf("head/shard_parts", (N, 3, R), torch.float32)
        self.call(self.lib.xv_head_shard_partials, L.ptr(pmax), L.ptr(psum), L.ptr(tgt), nblk, R, L.ptr(part), s())
        self.collective(lambda: sh.all_gather(parts, part))
        self.call(self.lib.xv_head_combine_shards, L.ptr(parts), N, R, C.c_float(inv_batch), L.ptr(lse), L.ptr(None),
                  L.ptr(self.scalars[0:1]), s())
        if training:
            def bwd():
                d = self.buf("head/d", (R, cpad), torch.bfloat16)
                self.gemm(L.operand(x3, False), L.operand(wn3, True, cols=Cn), R, Cn, 3 * E, d, epilogue=L.EPI_HEAD_BWD,
                          bias=bias_t, head=h, col_sum=(st.grad(bias) if bias is not None else None))
                gw = st.grad(kernel)                     # complete for this shard: all R rows are here
                self.gemm(L.operand(x3, True, cols=E), L.operand(d, True, cols=Cn), E, Cn, R, gw, epilogue=L.EPI_F32)
                if normalize:
                    self.call(self.lib.xv_head_finish_dw, L.ptr(gw), L.ptr(Wm), L.ptr(inv_norm), E, cpad, s())
                dx_all = self.buf("head/dxg_all", (R, E), torch.float32, zero=True)      # partial: this shard's columns
                sp = self.splits_for(R, E, Cn)
                self.gemm(L.operand(d, False, cols=Cn), L.operand(wn3, False, rows=E, cols=Cn), R, E, Cn, dx_all,
                          epilogue=L.EPI_F32, splits=sp)
                dxg = self.buf("head/dxg", (B, E), torch.float32)
                gn = self.buf("head/gnorm_local", (B,), torch.float32)

                def scatter_dx():
                    sh.reduce_scatter_sum(dxg, dx_all)
                    sh.reduce_scatter_sum(gn, gnorm)
                self.collective(scatter_dx)
                du = self.buf(u.name + "/grad", (B, E), torch.float32)
                use_margin = head_type != L.HEAD_SOFTMAX
                r0, r1 = rk * B, (rk + 1) * B
                self.call(self.lib.xv_head_finish_dx, L.ptr(dxg), L.ptr(gn if use_margin else None), L.ptr(x[r0:r1]),
                          L.ptr(xnorm[r0:r1]), L.ptr(u_loc), L.ptr(urinv[r0:r1]), C.c_float(scaling), L.ptr(du), B, E, s())
                u.grad = du
            self.tape.append(bwd)
        return self.scalars[0], None, x[rk * B:(rk + 1) * B]

    # ---- regulariser + optimizer ---------------------------------------------------------------------
    def l2_loss(self):
        st = self.store
        self.call(self.lib.xv_l2_loss, L.ptr(st.params), L.ptr(st.blk_l2), C.c_int64(st.n), L.ptr(self.scalars[1:2]),
                  L.stream_ptr())
        return self.scalars[1]

    def _set_scalars(self, dst, vals):
        arr = (C.c_float * len(vals))(*[float(v) for v in vals])
        L.check(self.lib.xv_set_scalars(L.ptr(dst), arr, len(vals), L.stream_ptr()))

    def set_hyper(self, lr, momentum=0.0, adam_t=1.0, clip_norm=0.0, flush=True):
        """Host -> device scalars (the learning rate placeholder is fed every step, trainer.py:326,493-494)."""
        if self.capturing:
            return
        self._hs_host[:8] = [lr, momentum, 0.9, 0.999, 1e-8, adam_t, clip_norm, 0.0]
        if flush:
            self._set_scalars(self._hs, self._hs_host)

    def set_sched(self, fa, fs):
        """Margin annealing factors of loss.py:144-147 for the current global_step (one launch also carries the
        hyper-parameters staged by set_hyper(flush=False))."""
        if self.capturing:
            return
        self._hs_host[8:10] = [fa, fs]
        self._set_scalars(self._hs, self._hs_host)

    def optimizer_step(self, opt, clip=False, with_l2_loss=False):
        """Fused L2 + clip + optimizer + bf16 shadow refresh; ``with_l2_loss`` also accumulates the regularisation
        loss of the (pre-update) parameters into scalars[1], replacing the separate l2_loss() pass."""
        st = self.store
        st.ensure_opt_state(opt)
        gs = None
        if clip:
            gs = self.scalars[2:3]
            self.call(self.lib.xv_grad_sumsq, L.ptr(st.params), L.ptr(st.grads), L.ptr(st.blk_l2), C.c_int64(st.n),
                      L.ptr(gs), L.stream_ptr())
        self.call(self.lib.xv_opt_step, L.ptr(st.params), L.ptr(st.grads), L.ptr(st.state1), L.ptr(st.state2),
                  L.ptr(st.blk_l2), L.ptr(st.blk_shadow), L.ptr(st.blk_stride), L.ptr(st.shadow), C.c_int64(st.n), opt,
                  L.ptr(self.hyper), L.ptr(gs), L.ptr(self.scalars[1:2] if with_l2_loss else None), L.stream_ptr())


_default_engine = None


def get_engine():
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine()
    return _default_engine


def set_engine(e):
    global _default_engine
    _default_engine = e
    return e
