"""Batch-sharded data parallelism: one process per GPU, ``torch.distributed`` for the plumbing.

The reference has no multi-GPU path at all (README.md:1,82,115), so correctness is defined as
"N-GPU step == 1-GPU step on the concatenated batch" up to BN statistics (per-replica here, stated wherever a
number is reported).  Every gradient lives in ONE flat fp32 buffer (runtime.ParamStore.grads), so the exchange is
a single sum all-reduce (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests); the head kernels already
scale dLoss/dlogit by 1/(N*B), so the sum is the global-batch gradient and the optimizer step is replicated.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun contract)."""
    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    rank = int(os.environ["RANK"])
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def shard_batch(features, labels, rank, world):
    """Replica r of N takes rows [r*B/N, (r+1)*B/N) of the global batch."""
    n = features.shape[0]
    assert n % world == 0, "global batch must divide by the number of replicas"
    per = n // world
    return features[rank * per:(rank + 1) * per], labels[rank * per:(rank + 1) * per]


class FlatAllReduce(object):
    """Sum all-reduce of a flat gradient buffer and broadcast of a flat parameter buffer."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def allreduce_(self, flat):
        if self.world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        return flat

    def broadcast_(self, flat, src=0):
        if self.world > 1:
            dist.broadcast(flat, src=src, group=self.group)
        return flat

    def mean_scalar(self, value, device=None):
        if self.world == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())


class HeadShard(object):
    """Column partition of the speaker matrix [E, C] over the ranks, and the exchanges of a class-sharded head step.

    Shard r owns classes [r*per, min((r+1)*per, C)) with per = ceil(C / N) rounded up to a multiple of 8 (aligned column
    offsets for the 16-byte tensor-map rows).  Exchanges (NCCL over NVLink on GPUs, gloo in the CPU tests):
      forward   all-gather of the embeddings [B, E] and labels [B]; all-gather of the per-row (max, sum, target) triples
      backward  sum reduce-scatter of dLoss/dx [N*B, E] and of the target-column ||x|| gradients [N*B]."""

    def __init__(self, num_outputs, rank=None, world=None, group=None):
        self.group = group
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rank, self.world, self.num_outputs = int(rank), int(world), int(num_outputs)
        per = -(-self.num_outputs // self.world)
        self.per = (per + 7) // 8 * 8
        if (self.world - 1) * self.per >= self.num_outputs:
            raise ValueError("%d classes cannot be split over %d ranks in multiples of 8 columns"
                             % (self.num_outputs, self.world))
        self.lo, self.hi = self.range_of(self.rank)
        self.n_local = self.hi - self.lo

    def range_of(self, r):
        lo = min(r * self.per, self.num_outputs)
        return lo, min(lo + self.per, self.num_outputs)

    def local_labels(self, labels):
        """Host-side statement of xv_head_local_labels (tests)."""
        l = labels - self.lo
        return torch.where((l >= 0) & (l < self.n_local), l, torch.full_like(l, -1))

    def all_gather(self, out, local):
        if self.world == 1:
            out.view(-1).copy_(local.reshape(-1))
        else:
            dist.all_gather_into_tensor(out.view(-1), local.reshape(-1), group=self.group)
        return out

    def reduce_scatter_sum(self, out, full):
        """out [n] <- rows [rank*n, (rank+1)*n) of the sum over ranks of full [world*n]."""
        n = out.numel()
        if self.world == 1:
            out.view(-1).copy_(full.reshape(-1))
        elif dist.get_backend(self.group) == "gloo":       # gloo has no reduce-scatter: all-reduce a copy, keep our rows
            tmp = full.reshape(-1).clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=self.group)
            out.view(-1).copy_(tmp[self.rank * n:(self.rank + 1) * n])
        else:
            dist.reduce_scatter_tensor(out.view(-1), full.reshape(-1), op=dist.ReduceOp.SUM, group=self.group)
        return out

    def gather_columns(self, local, pad_value=0.0):
        """[E, n_local] shards -> the full [E, C] matrix on every rank (checkpoint export)."""
        E = local.shape[0]
        buf = torch.full((E, self.per), pad_value, dtype=local.dtype, device=local.device)
        buf[:, :self.n_local] = local[:, :self.n_local]
        allb = torch.empty((self.world, E, self.per), dtype=local.dtype, device=local.device)
        self.all_gather(allb, buf)
        return torch.cat([allb[r, :, :(self.range_of(r)[1] - self.range_of(r)[0])] for r in range(self.world)], 1)


class SyncBN(object):
    """Exchanges of synchronised batch normalisation: with ``params.sync_bn`` every BN layer normalises
    with the statistics of the GLOBAL batch, so an N-GPU step equals the 1-GPU step on the concatenated batch.
    Frame-level layers all-reduce their per-channel (sum, sum of squares) forward and (dgamma, dbeta) backward;
    utterance-level layers all-gather their [B, C] rows (runtime.Engine.frame_affine / utt_affine)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_reduce_sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather(self, out, local):
        if self.world == 1:
            out.view(-1).copy_(local.reshape(-1))
        else:
            dist.all_gather_into_tensor(out.view(-1), local.reshape(-1), group=self.group)
        return out


class SegmentedGraph(object):
    """A step that contains host-side collectives, captured as consecutive CUDA graphs with the collectives replayed
    eagerly between them (all on the current stream, so stream order carries every dependency).  Engine.collective()
    calls cut() while the step is being captured."""

    def __init__(self):
        self.items = []
        self.pool = None
        self.stream = torch.cuda.Stream()
        self._g = None
        self._ctx = None

    def _open(self):
        self._g = torch.cuda.CUDAGraph()
        self._ctx = torch.cuda.graph(self._g, pool=self.pool, stream=self.stream, capture_error_mode="thread_local")
        self._ctx.__enter__()

    def _close(self):
        self._ctx.__exit__(None, None, None)
        if self.pool is None:
            self.pool = self._g.pool()
        self.items.append(self._g)
        self._g = self._ctx = None

    def begin(self):
        self._open()

    def cut(self, fn):
        self._close()
        self.items.append(fn)       # not executed during capture: no kernel has run, and every rank skips it alike
        self._open()

    def end(self):
        self._close()

    def abort(self):
        if self._ctx is not None:
            try:
                self._ctx.__exit__(None, None, None)
            except Exception:
                pass
            self._g = self._ctx = None

    def replay(self):
        for it in self.items:
            if isinstance(it, torch.cuda.CUDAGraph):
                it.replay()
            else:
                it()

    @property
    def num_graphs(self):
        return sum(isinstance(it, torch.cuda.CUDAGraph) for it in self.items)


class DataParallel(object):
    """Attach to a Trainer: per-rank batches of size B, loss scaled by 1/(N*B), one all-reduce per step."""

    def __init__(self, trainer, local_batch):
        self.trainer = trainer
        self.comm = FlatAllReduce()
        self.world, self.rank = self.comm.world, self.comm.rank
        eng = trainer.engine
        eng.inv_global_batch = 1.0 / float(self.world * local_batch)
        st = eng.store
        # class-sharded head: its kernel (declared last) differs per rank and its gradient is complete locally, so
        # broadcast / all-reduce cover only the replicated prefix of the flat buffers
        self.grad_dtype = str(getattr(getattr(trainer, "params", None), "dict", {}).get("dp_grad_dtype", "fp32"))
        self._g16 = None
        self.head_shard = getattr(eng, "head_shard", None)
        self.dp_numel = st.params.numel()
        if self.head_shard is not None:
            self.dp_numel = min(s.offset for s in st.specs.values() if s.col_range is not None)
            assert all(s.offset < self.dp_numel for s in st.specs.values() if s.trainable and s.col_range is None), \
                "sharded variables must be declared last"
        self.comm.broadcast_(st.params[:self.dp_numel])
        self.comm.broadcast_(st.buffers)
        # Gradient exchange ("dp_allreduce"): "auto" (default) / "multimem": the flat gradient buffer is re-homed in CUDA
        # symmetric memory and reduced in place by our own kernel INSIDE the captured step (xv_dp_allreduce_multimem: NVLS
        # in-switch reduction, N > 2; xv_dp_allreduce_p2p: peer loads / stores, N = 2); "auto" falls back to NCCL when
        # symmetric memory or the multicast mapping is unavailable.  "nccl": one torch.distributed all-reduce between two
        # graphs.  "symm": PyTorch's symm_mem multimem / two-shot library kernels (measured slower at 39 MB).
        self._symm = None
        self.allreduce_impl = "nccl"
        self._mm = None
        self.graph_safe = False          # True: the exchange is a plain kernel launch and may sit inside the captured step
        want = str(getattr(getattr(trainer, "params", None), "dict", {}).get("dp_allreduce", "auto"))
        if want in ("symm", "multimem", "auto") and self.world > 1 and st.params.is_cuda:
            self._try_symmetric_memory(eng, st, strict=(want != "auto"), own_kernel=(want != "symm"))
        params = getattr(trainer, "params", None)
        if params is not None and bool(params.dict.get("sync_bn", False)) and self.world > 1:
            eng.sync_bn = SyncBN()
        st.refresh_shadows()
        trainer.dp = self
        # The loss / penalty scalars are the first thing carved from the zero arena, i.e. they sit directly behind the
        # gradients in the same (symmetric) buffer: the gradient exchange covers them too, so the logged loss is the
        # global-batch mean on every rank without a second collective (Trainer._finish_step).
        self.scalars_reduced = False
        self.reduce_numel = self.dp_numel
        if self.head_shard is None and st.arena_used == 0:
            eng._scalars = None
            sc = eng.scalars
            if sc.data_ptr() == st.arena.data_ptr() and self.dp_numel == st.n:
                self.reduce_numel = st.n + 32
                self.scalars_reduced = True

    def _try_symmetric_memory(self, eng, st, strict, own_kernel=True):
        try:
            import torch.distributed._symmetric_memory as symm_mem
            if st.arena_used != 0 or eng.ws:
                raise RuntimeError("workspaces were already carved from the gradient buffer: attach DataParallel before the first step")
            group_name = dist.group.WORLD.group_name
            try:
                symm_mem.enable_symm_mem_for_group(group_name)
            except Exception:
                pass
            buf = symm_mem.empty(st.grads_ext.numel(), dtype=torch.float32, device=st.grads_ext.device)
            hdl = symm_mem.rendezvous(buf, group_name)
            buf.zero_()
            st.grads_ext = buf
            st.grads = buf[:st.n]
            st.arena = buf[st.n:]
            self._symm = hdl
            self._symm_group = group_name
            multicast = int(getattr(hdl, "multicast_ptr", 0) or 0) != 0
            self._symm_op = (torch.ops.symm_mem.multimem_all_reduce_ if multicast else torch.ops.symm_mem.two_shot_all_reduce_)
            self.allreduce_impl = "symm_mem multimem (NVLS)" if multicast else "symm_mem two-shot (peer memory)"
            if own_kernel:
                if not multicast:
                    raise RuntimeError("no multicast (NVLS) mapping for the symmetric gradient buffer")
                # xv_dp_allreduce_multimem: rank flags in a second symmetric buffer, per-block launch counters locally
                flags = symm_mem.empty(8192, dtype=torch.int32, device=buf.device)
                hdl_f = symm_mem.rendezvous(flags, group_name)
                flags.zero_()
                torch.cuda.synchronize()
                dist.barrier()               # nobody signals before every rank has zeroed its flags
                mode = os.environ.get("XV_AR_MODE") or ("p2p" if self.world <= 2 else "multimem")
                self._mm = {"mc": int(hdl.multicast_ptr), "bufs": int(hdl.buffer_ptrs_dev), "mode": mode,
                            "flags": int(hdl_f.buffer_ptrs_dev), "flags_t": flags, "hdl_f": hdl_f,
                            "epoch": torch.zeros(1024, dtype=torch.int32, device=buf.device),
                            "grid": int(min(eng.num_sms, 8192 // max(self.world, 1)))}
                self.graph_safe = True
                self.allreduce_impl = ("xv_dp_allreduce_multimem (multimem.ld_reduce / multimem.st over NVSwitch, in the captured step)"
                                       if mode == "multimem" else
                                       "xv_dp_allreduce_p2p (peer-memory loads / stores over NVLink, in the captured step)")
        except Exception as ex:
            self._symm = None
            self._mm = None
            self.graph_safe = False
            self.allreduce_impl = "nccl (symmetric memory unavailable: %s)" % (str(ex).splitlines()[0][:120] if str(ex) else type(ex).__name__)
            if strict:
                raise

    def allreduce_gradients(self):
        st = self.trainer.engine.store
        if self.scalars_reduced and self.grad_dtype != "bf16":
            g = st.grads_ext[:self.reduce_numel]         # gradients + the step's loss scalars
        else:
            g = st.grads[:self.dp_numel]
        if self._mm is not None and self.grad_dtype != "bf16":
            self.allreduce_range(0, g.numel())
            return
        if self._symm is not None and self.grad_dtype != "bf16":
            self._symm_op(g, "sum", self._symm_group)
            return
        if self.grad_dtype == "bf16" and self.world > 1:
            # opt-in: exchange the gradients in bf16 (39 -> 19.5 MB at config 2); every activation gradient of the
            # frame-level path is already stored in bf16, so this adds one more rounding of the same size per element
            import ctypes as C
            from . import _lib as L
            if self._g16 is None or self._g16.numel() != g.numel():
                self._g16 = torch.empty(g.numel(), dtype=torch.bfloat16, device=g.device)
            L.check(L.load().xv_grad_pack_bf16(L.ptr(g), L.ptr(self._g16), C.c_int64(g.numel()), L.stream_ptr()))
            self.comm.allreduce_(self._g16)
            L.check(L.load().xv_grad_unpack_bf16(L.ptr(self._g16), L.ptr(g), C.c_int64(g.numel()), L.stream_ptr()))
            return
        self.comm.allreduce_(g)

    def allreduce_range(self, lo, hi, grid=None):
        """In-graph exchange of gradient floats [lo, hi) (multiples of 4) with our own kernel on the CURRENT stream."""
        import ctypes as C
        from . import _lib as L
        m = self._mm
        grid = m["grid"] if grid is None else max(1, min(int(grid), m["grid"]))
        if m["mode"] == "multimem":
            L.check(L.load().xv_dp_allreduce_multimem(C.c_void_p(m["mc"]), C.c_void_p(m["flags"]), L.ptr(m["epoch"]), self.rank,
                                                      self.world, C.c_int64(lo), C.c_int64(hi - lo), grid, L.stream_ptr()))
        else:
            L.check(L.load().xv_dp_allreduce_p2p(C.c_void_p(m["bufs"]), C.c_void_p(m["flags"]), L.ptr(m["epoch"]), self.rank,
                                                 self.world, C.c_int64(lo), C.c_int64(hi - lo), grid, L.stream_ptr()))
        self.trainer.engine.launches += 1

    def mean_scalar(self, local_mean_scaled):
        # each rank's loss scalar is sum_i CE_i / (N*B): the global mean is their sum
        if self.head_shard is not None:
            return float(local_mean_scaled)      # the sharded head already sums the rows of every rank
        return self.comm.mean_scalar(local_mean_scaled, device=self.trainer.engine.device)

    def sum_scalar(self, value):
        return self.comm.mean_scalar(value, device=self.trainer.engine.device)
