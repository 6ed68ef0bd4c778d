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


class DataParallel(object):
    """Attach to a Trainer: per-rank batches of size B, loss scaled by 1/(N*B), one all-reduce per step."""

    def __init__(self, trainer, local_batch):
        self.trainer = trainer
        self.comm = FlatAllReduce()
        self.world, self.rank = self.comm.world, self.comm.rank
        eng = trainer.engine
        eng.inv_global_batch = 1.0 / float(self.world * local_batch)
        st = eng.store
        self.comm.broadcast_(st.params)
        self.comm.broadcast_(st.buffers)
        st.refresh_shadows()
        trainer.dp = self

        # Two gradient buckets of the flat buffer: [tdnn6 .. head] is complete once the utterance-level backward has
        # run (a third of the way into the backward pass), [tdnn1 .. pooling] only at its end.
        self.split = st.specs["tdnn/tdnn6_dense/kernel"].offset if "tdnn/tdnn6_dense/kernel" in st.specs else 0
        self._pending = []

    def allreduce_gradients(self):
        self.comm.allreduce_(self.trainer.engine.store.grads)

    def allreduce_bucket_async(self, which):
        """Enqueue the sum all-reduce of one gradient bucket on NCCL's stream (ordered after the work already on the
        current stream) and return immediately, so later kernels of the current stream overlap it."""
        g = self.trainer.engine.store.grads
        t = g[self.split:] if which == "head" else g[:self.split]
        if self.world > 1 and t.numel() > 0:
            self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.comm.group, async_op=True))

    def wait_all(self):
        for w in self._pending:
            w.wait()          # the current stream waits for the collective; the host does not block
        self._pending = []

    def mean_scalar(self, local_mean_scaled):
        # each rank's loss scalar is sum_i CE_i / (N*B): the global mean is their sum
        return self.comm.mean_scalar(local_mean_scaled, device=self.trainer.engine.device)
