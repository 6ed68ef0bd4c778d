"""Trainer with the reference's surface (model/trainer.py:17-726): build / train / valid / predict / save / load.

What changes: ``sess.run(train_op)`` (trainer.py:505-508) becomes ``train_step`` -- forward, backward, (data-parallel
gradient all-reduce,) fused optimizer and BN moving-statistics update, all hand-written sm_100a kernels enqueued
on one CUDA stream.  Host control flow (epochs, logging, checkpoints, LR bookkeeping) stays Python.  The Kaldi
feature pipeline (dataset/data_loader.py) stays on the host: ``train``/``valid`` take any object with the
``start()/fetch()/stop()`` queue protocol of KaldiDataRandomQueue (data_loader.py:310-414).
"""
import ctypes as C
import glob
import gc
import os
import re
import sys
import time

import numpy as np
import torch

from .. import _lib as L
from ..runtime import Engine, ScaledUtt, set_engine
from . import tdnn as tdnn_mod
from .tdnn import tdnn
from .loss import (softmax, asoftmax, additive_margin_softmax, additive_angular_margin_softmax,
                   semihard_triplet_loss, angular_triplet_loss, e2e_valid_loss, generalized_angular_triplet_loss,
                   METRIC_LOSSES,
                   declare_head_variables, margin_schedule)


class LossHandle(object):
    """Result of ``train_step(..., fetch_loss="async")``: the step's {"raw_loss", "loss"} once its 20-byte device-to-host
    copy has landed.  The pinned slot is reused after four further asynchronous fetches; an unread handle is latched first."""

    def __init__(self, trainer, host, event, seq):
        self._trainer, self._host, self._event, self._value, self._seq = trainer, host, event, None, seq

    def result(self):
        if self._value is None:
            self._event.synchronize()
            vals = self._host[:5].tolist()
            self._value = {"raw_loss": vals[0], "loss": vals[0] + vals[1] + vals[3]}
            if self._seq > self._trainer._loss_seq_read:       # train_ops mirrors the NEWEST step that has been read
                self._trainer._loss_seq_read = self._seq
                self._trainer.train_ops = dict(self._value)
            self._host = self._event = None
        return dict(self._value)


class Trainer(object):
    def __init__(self, params, model_dir, single_cpu=False, engine=None):
        """Args mirror model/trainer.py:24.  ``single_cpu`` is accepted and ignored (the CUDA path has no CPU mode)."""
        self.params = params
        if params.network_type == "tdnn":
            self.network = tdnn
        else:
            raise NotImplementedError("Not implement %s network" % params.network_type)
        self.loss_type = None
        self.loss_network = None
        self.model = os.path.join(model_dir, "nnet")
        self.model_log = os.path.join(model_dir, "log")
        self.engine = set_engine(engine if engine is not None else Engine())
        self.is_built = False
        self.is_loaded = False
        self.global_step = None
        self.num_speakers = None
        self.dim = None
        self.modes = set()
        self.opt = L.OPT_SGD
        self.adam_t = 0
        self.dp = None               # optional data-parallel wrapper (parallel.DataParallel)
        self.endpoints = None
        self.embeddings = None
        self.train_ops = {}
        self.valid_ops = {}
        self._static = {}
        self._copy_stream = None
        self._h2d_event = None
        self._step_done = None
        self._up = None
        self._loss_ring = None
        self._loss_seq = 0
        self._loss_seq_read = 0
        self.use_cuda_graph = bool(params.dict.get("cuda_graph", True))
        self._captured = 0           # batch shapes whose step has been captured
        self._in_memory = False      # True once this object holds parameters newer than (or loaded from) the checkpoint

    # ------------------------------------------------------------------ network (trainer.py:168-188)
    def entire_network(self, features, params, is_training, reuse_variables, lengths=None, ragged=None):
        features, endpoints = self.network(features, params, is_training, reuse_variables, lengths=lengths, ragged=ragged)
        endpoints["output"] = features
        if "feature_norm" in params.dict and params.feature_norm:
            assert "feature_scaling_factor" in params.dict, "If feature normalization is applied, scaling factor is necessary."
            features = ScaledUtt(features, params.feature_scaling_factor)
            endpoints["output"] = features
        return features, endpoints

    # ------------------------------------------------------------------ build (trainer.py:190-449)
    def build(self, mode, dim, loss_type=None, num_speakers=None, noupdate_var_list=None):
        assert (mode == "train" or mode == "valid" or mode == "predict")
        if noupdate_var_list is not None:
            raise NotImplementedError("noupdate_var_list (fine-tuning) is outside the accelerated path")
        eng = self.engine
        self.dim = dim
        if mode != "predict":
            self.loss_type = loss_type
            if loss_type == "softmax":
                self.loss_network = softmax
            elif loss_type == "asoftmax":
                self.loss_network = asoftmax
            elif loss_type == "additive_margin_softmax":
                self.loss_network = additive_margin_softmax
            elif loss_type == "additive_angular_margin_softmax":
                self.loss_network = additive_angular_margin_softmax
            elif loss_type == "semihard_triplet_loss":
                self.loss_network = semihard_triplet_loss
            elif loss_type == "angular_triplet_loss":
                self.loss_network = angular_triplet_loss
            elif loss_type == "generalized_angular_triplet_loss":
                self.loss_network = generalized_angular_triplet_loss
            else:
                raise NotImplementedError("Not implement %s loss" % self.loss_type)
            self.num_speakers = num_speakers
            if self.global_step is None:
                self.global_step = 0
                self.params.dict["global_step"] = 0
        if mode == "train":
            if "optimizer" not in self.params.dict:
                self.params.dict["optimizer"] = "sgd"
            if self.params.optimizer == "sgd":
                if "momentum" in self.params.dict:
                    sys.exit("Using sgd as the optimizer and you should not specify the momentum.")
                self.opt = L.OPT_SGD
            elif self.params.optimizer == "momentum":
                self.opt = L.OPT_NESTEROV if self.params.use_nesterov else L.OPT_MOMENTUM
            elif self.params.optimizer == "adam":
                self.opt = L.OPT_ADAM
            else:
                sys.exit("Optimizer %s is not supported." % self.params.optimizer)
        if (mode != "predict" and not eng.store.finalized and bool(self.params.dict.get("head_class_shard", False))):
            # north_star "Data parallelism": optionally split the speaker matrix by columns over the ranks
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size() > 1:
                from ..parallel import HeadShard
                eng.head_shard = HeadShard(num_speakers)
                if bool(self.params.dict.get("clip_gradient", False)):
                    raise NotImplementedError("clip_gradient needs the global gradient norm: not available with "
                                              "head_class_shard")
        if not eng.store.finalized:
            tdnn_mod.declare_variables(eng, dim, self.params)
            if mode != "predict":
                e = int(self.params.dict.get("num_nodes_last_layer", 512))
                if loss_type not in METRIC_LOSSES:         # the metric-learning losses have no speaker matrix
                    declare_head_variables(eng, e, num_speakers, self.params, loss_type)
            eng.store.finalize()
            eng.store.init(int(self.params.dict.get("seed", 0)))
        elif mode != "predict" and loss_type not in METRIC_LOSSES and "softmax/output/kernel" not in eng.store:
            raise L.XvError("build('predict') was called first: the head variables cannot be added afterwards; "
                            "build train/valid before predict")
        self.modes.add(mode)
        self.is_built = True

    # ------------------------------------------------------------------ one step = sess.run(train_op)
    def _to_device(self, features, labels=None):
        dev = self.engine.device
        if hasattr(features, "decode_into"):       # dataset.feeder.CompressedSegmentBatch: raw uint8 crops, decoded on the device
            features = features.to_device(dev)
        if not torch.is_tensor(features):
            features = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
        features = features.to(dev, dtype=torch.float32, non_blocking=True)
        if labels is not None:
            if not torch.is_tensor(labels):
                labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32))
            labels = labels.to(dev, dtype=torch.int32, non_blocking=True)
        return features, labels

    def _fwd_bwd(self, features, labels, l2_loss=True):
        """Forward + backward of one device-resident batch; gradients are left in the flat gradient buffer.
        ``l2_loss=False``: the regularisation loss is left to the optimizer kernel (same pass over the parameters)."""
        eng = set_engine(self.engine)      # the operator functions (model/tdnn.py, loss.py) act on the current engine
        eng.begin_step(True)
        if self.loss_type not in METRIC_LOSSES:
            eng.prefetch_head_weights("softmax/output/kernel", normalize=(self.loss_type != "softmax"))
        if self._h2d_event is not None:
            self._h2d_event.wait(torch.cuda.current_stream())     # this step's batch has arrived (see _static_batch)
        out, endpoints = self.entire_network(features, self.params, True, True)
        loss, endpoints_loss = self.loss_network(out, labels, self.num_speakers, self.params, True, True)
        endpoints.update(endpoints_loss)
        self.endpoints = endpoints
        if l2_loss:
            eng.l2_loss()
        eng.backward()

    def forward_backward(self, features, labels, global_step):
        """Eager forward + backward (no optimizer step); used by tests and gradient inspection."""
        eng = self.engine
        features, labels = self._to_device(features, labels)
        self.params.dict["global_step"] = int(global_step)
        eng.set_sched(*margin_schedule(self.loss_type, self.params, global_step))
        self._fwd_bwd(features, labels)
        return eng.scalars[0], eng.scalars[1]

    def _static_batch(self, features, labels):
        """Static device buffers per batch shape (a captured CUDA graph replays on fixed addresses)."""
        key = tuple(features.shape)
        st = self._static.get(key)
        if st is None:
            dev = self.engine.device
            st = {"x": torch.empty(key, dtype=torch.float32, device=dev),
                  "y": torch.empty((key[0],), dtype=torch.int32, device=dev), "calls": 0, "graphs": None, "launches": 0}
            self._static[key] = st
        if not torch.is_tensor(labels):
            labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32))
        if not hasattr(features, "decode_into") and not torch.is_tensor(features):
            features = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
        from_host = hasattr(features, "decode_into") or not features.is_cuda
        main = torch.cuda.current_stream()
        if self._h2d_event is None:
            self._copy_stream = torch.cuda.Stream(device=self.engine.device)
            self._h2d_event = torch.cuda.Event(enable_timing=False, external=True)
        if not from_host:
            # a batch written straight into the static buffers (Trainer.input_buffers) needs no copy at all
            if features.data_ptr() != st["x"].data_ptr():
                st["x"].copy_(features, non_blocking=True)
            if not (labels.is_cuda and labels.data_ptr() == st["y"].data_ptr()):
                st["y"].copy_(labels.to(torch.int32), non_blocking=True)
            self._h2d_event.record(main)
            return st
        # Host batches go through two device staging buffers on a copy stream: the upload of step n+1 runs while step n is
        # still computing (it only has to wait until step n-1's staging buffer has been handed over); a 3 MB device-to-device
        # copy then moves it into the static buffers the captured graph reads.  Without the staging pair the upload could not
        # start before the previous step had finished with the static buffers (60 us of PCIe time exposed per step).
        if "stage" not in st:
            dev = self.engine.device
            st["stage"] = [(torch.empty(key, dtype=torch.float32, device=dev), torch.empty((key[0],), dtype=torch.int32, device=dev))
                           for _ in range(2)]
            st["k"] = 0
        k = st["k"] = 1 - st["k"]
        sx, sy = st["stage"][k]
        cs = self._copy_stream          # in-order: the upload into stage k follows the hand-over that last read stage k
        with torch.cuda.stream(cs):
            if hasattr(features, "decode_into"):
                # dataset.feeder.CompressedSegmentBatch: H2D of the raw uint8 crops + on-device dequantise / transpose
                features.decode_into(sx)
                self.engine.launches += 1
            else:
                sx.copy_(features, non_blocking=True)            # H2D of this step's batch
            sy.copy_(labels.to(torch.int32), non_blocking=True)
            # ... and, once the previous step has finished with the static buffers, the device-to-device hand-over -- still on
            # the copy stream, so that the step's first kernels (gradient fill, scalar feed, head weight preparation) run
            # beside it; the captured graph waits for _h2d_event only in front of the first kernel that reads the features
            if self._step_done is not None:
                cs.wait_event(self._step_done)
            st["x"].copy_(sx, non_blocking=True)
            st["y"].copy_(sy, non_blocking=True)
            self._h2d_event.record(cs)
        st["host_ref"] = (features, labels)        # keep the (pinned) source alive until the next call
        return st

    def input_buffers(self, batch, frames, dim=None):
        """The static device tensors (features float32 [batch, frames, dim], labels int32 [batch]) the captured step of this
        batch shape reads.  A producer that already works on the device (an on-device augmentation or decode kernel) can
        write the next batch straight into them and pass them to train_step, which then skips its device-to-device copy."""
        key = (int(batch), int(frames), int(dim if dim is not None else self.dim))
        st = self._static.get(key)
        if st is None:
            dev = self.engine.device
            st = {"x": torch.empty(key, dtype=torch.float32, device=dev),
                  "y": torch.empty((key[0],), dtype=torch.int32, device=dev), "calls": 0, "graphs": None, "launches": 0}
            self._static[key] = st
        return st["x"], st["y"]

    def train_step(self, features, labels, learning_rate, global_step=None, fetch_loss=False):
        """The hot-loop body = sess.run(train_op) (trainer.py:491-508): forward, backward, [all-reduce], optimizer,
        BN moving statistics.  After two eager warm-up calls per batch shape the whole step is captured into a CUDA
        graph and replayed; learning rate / margin schedule live in device scalars set before each replay.
        Returns {"loss": total, "raw_loss": loss} when fetch_loss (one 20-byte D2H read, synchronous), a LossHandle when
        fetch_loss == "async" (the same read queued behind the step; ``.result()`` waits for it), else None."""
        eng = self.engine
        if global_step is None:
            global_step = self.global_step or 0
        st = self._static_batch(features, labels)
        self.params.dict["global_step"] = int(global_step)
        if self.opt == L.OPT_ADAM:
            self.adam_t += 1
        clip = bool(self.params.dict.get("clip_gradient", False))
        eng.set_hyper(float(learning_rate), float(self.params.dict.get("momentum", 0.0)), float(max(self.adam_t, 1)),
                      float(self.params.dict.get("clip_gradient_norm", 0.0)) if clip else 0.0, flush=False)
        eng.set_sched(*margin_schedule(self.loss_type, self.params, global_step))

        def part_a():
            self._fwd_bwd(st["x"], st["y"], l2_loss=False)

        def part_b():
            if eng.head_shard is not None:       # regularisation loss of this rank's head columns, for the logged total
                hs = eng.store.specs["softmax/output/kernel"]
                nb = (hs.numel + 1023) // 1024 * 1024
                eng.call(eng.lib.xv_l2_loss, L.ptr(eng.store.params[hs.offset:]), L.ptr(eng.store.blk_l2[hs.offset // 1024:]),
                         C.c_int64(nb), L.ptr(eng.scalars[4:5]), L.stream_ptr())
            eng.optimizer_step(self.opt, clip=clip, with_l2_loss=True)

        if self.dp is not None and (eng.head_shard is not None or eng.sync_bn is not None):
            # class-sharded head / SyncBN: the step contains collectives (row all-gather, partial exchange, dx
            # reduce-scatter, BN statistics, trunk-gradient all-reduce); it is captured as consecutive CUDA graphs with
            # the exchanges between them
            def body():
                part_a()
                eng.collective(self.dp.allreduce_gradients)
                part_b()

            if st["graphs"] is not None:
                st["graphs"].replay()
                eng.launches += st["launches"]
            else:
                st["calls"] += 1
                if self.use_cuda_graph and st["calls"] > 2:
                    from ..parallel import SegmentedGraph
                    seg = SegmentedGraph()
                    l0 = eng.launches
                    eng.capturing, eng.segmenter = True, seg
                    try:
                        seg.begin()
                        body()
                        seg.end()
                    except Exception:
                        seg.abort()
                        raise
                    finally:
                        eng.capturing, eng.segmenter = False, None
                    st["launches"] = eng.launches - l0
                    st["graphs"] = seg
                    seg.replay()
                else:
                    body()
            return self._finish_step(global_step, fetch_loss)

        # Default: the gradient exchange is our own kernel with in-kernel rank barriers (xv_dp_allreduce_*), so the whole
        # data-parallel step (forward, backward, exchange, optimizer) is ONE captured graph, like the single-GPU step.
        # Fallback (no symmetric memory / dp_allreduce = "nccl"): graph(forward + backward) | NCCL all-reduce | graph(optimizer).
        in_graph_exchange = self.dp is not None and bool(getattr(self.dp, "graph_safe", False))

        def run(ga, gb):
            ga.replay() if ga is not None else part_a()
            if self.dp is None or in_graph_exchange:
                if ga is None:
                    if in_graph_exchange:
                        self.dp.allreduce_gradients()
                    part_b()
                return
            self.dp.allreduce_gradients()
            gb.replay() if gb is not None else part_b()

        if st["graphs"] is not None and st.get("gen") != eng.ws_generation:
            st["graphs"], st["calls"] = None, 1      # a scratch buffer grew since the capture: its addresses are stale
        if st["graphs"] is not None:
            run(*st["graphs"])
            eng.launches += st["launches"]
        else:
            st["calls"] += 1
            # two eager calls for the first batch shape (lazy allocations, optimizer slots), one for every further
            # segment length (its views of the shared scratch buffers), then capture
            warm = 2 if self._captured == 0 else 1
            if self.use_cuda_graph and st["calls"] > warm:
                l0 = eng.launches
                eng.capturing = True
                # Capture in "thread_local" error mode with the garbage collector paused: in the default "global" mode ANY
                # thread's unsafe CUDA call invalidates the capture -- the loader's feeder thread (pinned allocations, event
                # synchronisation) runs beside the step, and a collection pass that finalises a stale event or pinned buffer
                # in the middle of the capture does the same (seen once in ~3 full test runs).
                gc_was_on = gc.isenabled()
                gc.disable()
                try:
                    ga = torch.cuda.CUDAGraph()
                    gb = None
                    with torch.cuda.graph(ga, capture_error_mode="thread_local"):
                        part_a()
                        if self.dp is None:
                            part_b()
                        elif in_graph_exchange:
                            self.dp.allreduce_gradients()
                            part_b()
                    if self.dp is not None and not in_graph_exchange:
                        gb = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gb, capture_error_mode="thread_local"):
                            part_b()
                finally:
                    eng.capturing = False
                    if gc_was_on:
                        gc.enable()
                st["launches"] = eng.launches - l0
                st["graphs"] = (ga, gb)
                st["gen"] = eng.ws_generation
                self._captured += 1
                run(ga, gb)
            else:
                run(None, None)
        if self._step_done is None:
            self._step_done = torch.cuda.Event()
        self._step_done.record()          # the static input buffers may be overwritten from here on (see _static_batch)
        return self._finish_step(global_step, fetch_loss)

    def _finish_step(self, global_step, fetch_loss):
        eng = self.engine
        self.global_step = int(global_step) + 1
        if fetch_loss == "async" and (self.dp is None or (self.dp.scalars_reduced and eng.head_shard is None)):
            # the step's loss record is copied to pinned host memory behind the step; the caller reads it when it needs it
            # (LossHandle.result()), typically one step later, so the host never waits for the step it has just queued
            if self._loss_ring is None:
                self._loss_ring = [[torch.empty(8, dtype=torch.float32).pin_memory(), torch.cuda.Event(), None] for _ in range(4)]
                self._loss_slot = 0
            self._loss_slot = (self._loss_slot + 1) % len(self._loss_ring)
            slot = self._loss_ring[self._loss_slot]
            if slot[2] is not None:
                slot[2].result()          # a handle nobody has read yet owns this slot: latch its value (copied 4 steps ago)
            host, ev = slot[0], slot[1]
            host[:5].copy_(eng.scalars[:5], non_blocking=True)
            ev.record()
            self._loss_seq += 1
            slot[2] = LossHandle(self, host, ev, self._loss_seq)
            return slot[2]
        if fetch_loss:
            vals = eng.scalars[:5].tolist()          # one D2H read
            raw, l2, pen = vals[0], vals[1], vals[3]
            if self.dp is not None and not self.dp.scalars_reduced:
                # NCCL fallback / segmented step: the loss and penalty scalars are rank-local sums over the local rows
                raw = self.dp.mean_scalar(raw)
                if float(self.params.dict.get("att_penalty_term", 0.0) or 0.0) != 0.0:      # same decision on every rank
                    pen = self.dp.sum_scalar(pen)
            if self.dp is not None and eng.head_shard is not None:       # regularisation loss of the other ranks' head columns
                l2 += self.dp.sum_scalar(vals[4]) - vals[4]
            self.train_ops = {"raw_loss": raw, "loss": raw + l2 + pen}
            return dict(self.train_ops)
        return None

    def reserve(self, batch, max_frames, dim=None):
        """Size every scratch buffer for the longest batch the loop will see ([batch, max_frames, dim]) with one eager
        forward/backward on zeros, so that no buffer grows (and no captured step is invalidated) when a longer segment
        length is drawn later.  Moving statistics are restored; gradients are cleared by the next step anyway."""
        eng = self.engine
        dim = self.dim if dim is None else dim
        saved = eng.store.buffers.clone()
        x = torch.zeros((int(batch), int(max_frames), int(dim)), dtype=torch.float32, device=eng.device)
        y = torch.zeros((int(batch),), dtype=torch.int32, device=eng.device)
        eng.set_sched(*margin_schedule(self.loss_type, self.params, self.global_step or 0))
        self._fwd_bwd(x, y)
        eng.store.buffers.copy_(saved)
        torch.cuda.synchronize()

    def _open_loader(self, data, spklist, kind, **kw):
        """``data``: a Kaldi data directory (the reference's call, train.py:98-104) or an object that already speaks the
        start()/fetch()/stop() protocol of KaldiDataRandomQueue."""
        if hasattr(data, "fetch"):
            return data
        from ..dataset.data_loader import KaldiDataRandomQueue, KaldiDataSeqQueue
        p = self.params
        if kind == "random":
            return KaldiDataRandomQueue(data, spklist, num_parallel=p.num_parallel_datasets, max_qsize=p.max_queue_size,
                                        num_speakers=kw.get("num_speakers", p.num_speakers_per_batch),
                                        num_segments=kw.get("num_segments", p.num_segments_per_speaker),
                                        min_len=p.min_segment_len, max_len=p.max_segment_len, shuffle=True,
                                        base_seed=p.dict.get("data_seed"))
        return KaldiDataSeqQueue(data, spklist, num_parallel=2, max_qsize=10,
                                 batch_size=p.num_speakers_per_batch * p.num_segments_per_speaker,
                                 min_len=p.min_segment_len, max_len=p.max_segment_len, shuffle=kw.get("shuffle", True),
                                 base_seed=p.dict.get("data_seed"))

    def train(self, data, spklist, learning_rate, aux_data=None):
        """One epoch (model/trainer.py:451-520).  ``data`` / ``spklist``: the training data directory and the speaker
        list, as nnet/lib/train.py:98 passes them (a loader object with start()/fetch()/stop() is accepted too)."""
        from ..dataset.data_loader import DataOutOfRange
        assert "train" in self.modes
        curr_step = 0
        if self._in_memory:                      # this object has been training: its state is the newest there is
            curr_step = self.global_step or 0
        elif os.path.isfile(os.path.join(self.model, "checkpoint")):
            curr_step = self.load()              # trainer.py:467-469
        data_loader = self._open_loader(data, spklist, "random")
        if not hasattr(data, "fetch") and "max_segment_len" in self.params.dict:
            shape = (self.params.num_speakers_per_batch * self.params.num_segments_per_speaker, self.params.max_segment_len)
            if getattr(self, "_reserved", None) != shape:
                self.reserve(*shape)
                self._reserved = shape
        if hasattr(data_loader, "start"):
            data_loader.start()
        steps = int(self.params.num_steps_per_epoch)
        epoch = int(curr_step / steps)
        pending = None
        try:
            for step in range(curr_step % steps, steps):
                try:
                    show = (step % int(self.params.save_summary_steps) == 0 or
                            step % int(self.params.show_training_progress) == 0)
                    t0 = time.time()
                    features, labels = data_loader.fetch()
                    res = self.train_step(features, labels, learning_rate, curr_step, fetch_loss="async" if show else False)
                    if pending is not None:      # the progress line of the previous logged step: its loss has landed by now
                        self._print_progress(*pending)
                        pending = None
                    if show:
                        pending = (epoch, step, steps, t0, res)
                    self._in_memory = True
                    if step % int(self.params.save_checkpoints_steps) == 0 and curr_step != 0:
                        self.save(curr_step)
                    curr_step += 1
                except DataOutOfRange:
                    print("Finished reading features.")
                    break
        finally:
            if pending is not None:
                self._print_progress(*pending)
            if hasattr(data_loader, "stop"):
                data_loader.stop()
        self.global_step = curr_step
        self.save(curr_step)
        return

    @staticmethod
    def _print_progress(epoch, step, steps, t0, res):
        r = res.result() if hasattr(res, "result") else res
        print("Epoch: [%2d] step: [%2d/%2d] time: %.4f s/step, raw loss: %f, total loss: %f"
              % (epoch, step, steps, time.time() - t0, r["raw_loss"], r["loss"]), flush=True)

    # ------------------------------------------------------------------ validation (trainer.py:261-303, 592-706)
    def _valid_params(self):
        vp = type(self.params).__new__(type(self.params))
        vp.__dict__.update(self.params.dict)
        if self.loss_type == "asoftmax":
            vp.asoftmax_m = 1
        elif self.loss_type == "additive_margin_softmax":
            vp.amsoftmax_m = 0
        elif self.loss_type == "additive_angular_margin_softmax":
            vp.arcsoftmax_m = 0
        if "aux_loss_func" in vp.dict:
            vp.aux_loss_func = []
        return vp

    def valid_step(self, features, labels, with_loss=True):
        """Loss of the validation graph (is_training=False, margins neutralised) and the output embeddings.
        ``with_loss=False``: embeddings only (the ordered pass of model/trainer.py:624-655 fetches no loss)."""
        eng = set_engine(self.engine)
        features, labels = self._to_device(features, labels)
        vp = self._valid_params()
        vp.dict["global_step"] = self.global_step or 0
        eng.begin_step(False)
        out, endpoints = self.entire_network(features, vp, False, True)
        if not with_loss:
            self.endpoints = endpoints
            return None, endpoints["output"].dense()
        # angular triplet training validates with the softmax GE2E loss (model/trainer.py:272-275, 300-301)
        loss_network = e2e_valid_loss if self.loss_type == "angular_triplet_loss" else self.loss_network
        loss, _ = loss_network(out, labels, self.num_speakers, vp, False, True)
        self.endpoints = endpoints
        return float(loss.item()), endpoints["output"].dense()

    def valid(self, data, spklist, batch_type="softmax", output_embeddings=False, aux_data=None):
        """model/trainer.py:592-706: mean validation loss over at most ``valid_max_iterations`` batches (margins
        neutralised, BN in inference mode) and, with ``output_embeddings``, the embeddings / labels of every segment read in
        order.  ``data``: the validation data directory (or a loader object)."""
        from ..dataset.data_loader import DataOutOfRange
        assert "valid" in self.modes or "train" in self.modes
        assert batch_type == "softmax" or batch_type == "end2end", "The batch_type can only be softmax or end2end"
        if not self._in_memory:
            if os.path.isfile(os.path.join(self.model, "checkpoint")):
                self.load()
            else:
                print("[Warning] Cannot find model in %s. Random initialization is used in validation." % self.model)
        embeddings_val, labels_val = None, None
        if output_embeddings:
            loader = self._open_loader(data, spklist, "seq", shuffle=False)
            if hasattr(loader, "start"):
                loader.start()
            embs, labs = [], []
            while True:
                try:
                    features, labels = loader.fetch()
                except (DataOutOfRange, StopIteration):
                    break
                _, e = self.valid_step(features, labels, with_loss=False)
                embs.append(e.cpu().numpy())
                labs.append(np.asarray(labels))
                if hasattr(data, "fetch") and len(embs) >= int(self.params.valid_max_iterations):
                    break
            if hasattr(loader, "stop"):
                loader.stop()
            if embs:
                embeddings_val, labels_val = np.concatenate(embs, 0), np.concatenate(labs, 0)
            if hasattr(data, "fetch"):          # a caller-supplied loader is consumed once: both results from this pass
                return 0.0, embeddings_val, labels_val
        if batch_type == "softmax":
            loader = self._open_loader(data, spklist, "seq", shuffle=True)
        else:
            assert "num_valid_speakers_per_batch" in self.params.dict and "num_valid_segments_per_speaker" in self.params.dict, \
                "Valid parameters should be set if E2E loss is selected"
            loader = self._open_loader(data, spklist, "random", num_speakers=self.params.num_valid_speakers_per_batch,
                                       num_segments=self.params.num_valid_segments_per_speaker)
        if hasattr(loader, "start"):
            loader.start()
        losses = []
        for _ in range(int(self.params.valid_max_iterations)):
            try:
                features, labels = loader.fetch()
            except (DataOutOfRange, StopIteration):
                break
            l, _ = self.valid_step(features, labels)
            losses.append(l)
        if hasattr(loader, "stop"):
            loader.stop()
        loss = float(np.mean(losses)) if losses else 0.0
        print("[Validation %d batches] valid loss: %f" % (len(losses), loss), flush=True)
        return loss, embeddings_val, labels_val

    # ------------------------------------------------------------------ prediction (trainer.py:708-726)
    def predict(self, features):
        """features: np [T, D] or [N, T, D] -> embeddings np [E] or [N, E] from endpoints[embedding_node]."""
        features = np.asarray(features, dtype=np.float32)
        rank = features.ndim
        if rank == 2:
            features = features[None]
        emb = self.predict_batch_padded(features, None)
        return emb[0] if rank == 2 else emb

    def _upload(self, features):
        """Pinned host batch -> one of two persistent device buffers, on the copy stream: the upload of batch n+1 runs beside
        the kernels of batch n (a buffer is overwritten only after the batch that read it has finished).  Anything else
        (numpy, pageable, device tensors) takes the plain path.  Returns (device tensor, release callback)."""
        if not (torch.is_tensor(features) and not features.is_cuda and features.is_pinned() and features.dtype == torch.float32):
            return self._to_device(features)[0], (lambda: None)
        dev = self.engine.device
        if self._up is None:
            self._up = [{"buf": None, "free": None, "ready": torch.cuda.Event()} for _ in range(2)]
            self._up_k = 0
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
        self._up_k = 1 - self._up_k
        slot = self._up[self._up_k]
        n = features.numel()
        cs, main = self._copy_stream, torch.cuda.current_stream()
        if slot["buf"] is None or slot["buf"].numel() < n:
            if slot["free"] is not None:
                slot["free"].synchronize()
            slot["buf"] = torch.empty(int(n * 1.25), dtype=torch.float32, device=dev)
        if slot["free"] is not None:
            cs.wait_event(slot["free"])
        with torch.cuda.stream(cs):
            d = slot["buf"][:n].view(features.shape)
            d.copy_(features, non_blocking=True)
            slot["ready"].record(cs)
        main.wait_event(slot["ready"])

        def release():
            if slot["free"] is None:
                slot["free"] = torch.cuda.Event()
            slot["free"].record(main)
        return d, release

    def predict_batch_padded(self, features, lengths, as_device=False):
        """[N, Tmax, D] (+ optional lengths [N]) -> np [N, E]; rows are independent (BN in inference mode, masked
        pooling), so a ragged batch gives the same result as one call per utterance (extract.py:90).
        ``as_device``: return a device tensor (a copy: the workspace is reused by the next call) without synchronising,
        so the caller can overlap the next batch's upload with this one's compute."""
        eng = set_engine(self.engine)
        feats, release = self._upload(features)
        ln = None if lengths is None else torch.as_tensor(np.asarray(lengths), dtype=torch.int32).to(eng.device, non_blocking=True)
        eng.begin_step(False)
        _, endpoints = self.entire_network(feats, self.params, False, True, lengths=ln)
        release()
        self.endpoints = endpoints
        node = endpoints[self.params.embedding_node]
        if as_device:
            return node.dense().float().clone()
        return node.dense().float().cpu().numpy()

    def predict_ragged(self, flat_features, starts, lengths, as_device=False):
        """Extraction without padding: ``flat_features`` [R, D] holds the utterances back to back, utterance i occupying
        rows [starts[i], starts[i] + lengths[i]) -> [N, E] embeddings, identical to one predict() call per utterance
        (a valid frame never reads rows of a neighbouring utterance; rows straddling two utterances are never pooled).
        Statistics pooling only."""
        if self.params.pooling_type != "statistics_pooling":
            raise NotImplementedError("predict_ragged supports statistics_pooling; use predict_batch_padded")
        eng = set_engine(self.engine)
        feats, release = self._upload(flat_features)
        feats = feats.view(1, feats.shape[0], feats.shape[1])
        st = torch.as_tensor(np.asarray(starts), dtype=torch.int32).to(eng.device, non_blocking=True)
        ln = torch.as_tensor(np.asarray(lengths), dtype=torch.int32).to(eng.device, non_blocking=True)
        eng.begin_step(False)
        _, endpoints = self.entire_network(feats, self.params, False, True, ragged=(st, ln))
        release()
        self.endpoints = endpoints
        node = endpoints[self.params.embedding_node]
        if not hasattr(node, "col_map"):
            raise NotImplementedError("predict_ragged: embedding_node must be an utterance-level endpoint")
        if as_device:
            return node.dense().float().clone()
        return node.dense().float().cpu().numpy()

    # ------------------------------------------------------------------ checkpoints (trainer.py:142-166)
    def _export_full(self, which):
        """Variables (or optimizer slots) in their full TF shapes; the column shards of a class-sharded head are
        gathered (a collective: every rank calls this)."""
        st = self.engine.store
        vals = st.export_tf(which=which)
        sh = self.engine.head_shard
        if sh is not None:
            for name, spec in st.specs.items():
                if spec.col_range is not None and name in vals:
                    loc = torch.from_numpy(vals[name]).to(self.engine.device)
                    full = sh.gather_columns(loc.reshape(-1, loc.shape[-1]))
                    vals[name] = full.reshape(spec.full_shape).cpu().numpy()
        return vals

    def _slot_names(self):
        """TF slot-variable suffixes of the optimizer in use (tf.train.MomentumOptimizer / AdamOptimizer)."""
        if self.opt in (L.OPT_MOMENTUM, L.OPT_NESTEROV):
            return {"state1": "/Momentum"}
        if self.opt == L.OPT_ADAM:
            return {"state1": "/Adam", "state2": "/Adam_1"}
        return {}

    def save(self, step):
        """Every rank calls save() (the sharded variables are gathered collectively); rank 0 alone writes, to a temporary
        name followed by os.replace, and the ranks meet at a barrier before anyone goes on (a later load() or pruning
        never sees a torn file)."""
        import torch.distributed as dist
        multi = self.dp is not None and dist.is_initialized() and dist.get_world_size() > 1
        rank = dist.get_rank() if multi else 0
        vals = self._export_full("params")
        slots = {}
        for which, suffix in self._slot_names().items():
            if getattr(self.engine.store, which) is not None:
                for k, v in self._export_full(which).items():
                    slots[k + suffix] = v
        if rank == 0:
            os.makedirs(self.model, exist_ok=True)
            keep = int(self.params.dict.get("keep_checkpoint_max", 5))
            if str(self.params.dict.get("checkpoint_format", "npz")) == "tf":
                # TF tensor-bundle files under the reference's names (model-<step>.index / .data-00000-of-00001), slot
                # variables under tf.train.Saver's names so that a reference training graph restores them
                from ..misc.tf_checkpoint import write_tf_checkpoint
                allv = dict(vals)
                allv.update(slots)
                allv["global_step"] = np.array(step, dtype=np.int64)
                if self.opt == L.OPT_ADAM:
                    allv["beta1_power"] = np.array(0.9 ** max(self.adam_t, 0), dtype=np.float32)
                    allv["beta2_power"] = np.array(0.999 ** max(self.adam_t, 0), dtype=np.float32)
                tmp = os.path.join(self.model, ".tmp-model-%d" % step)
                write_tf_checkpoint(tmp, allv)
                for ext in (".data-00000-of-00001", ".index"):
                    os.replace(tmp + ext, os.path.join(self.model, "model-%d%s" % (step, ext)))
                pattern, rx = "model-*.index", r"model-(\d+)\.index"
            else:
                path = os.path.join(self.model, "model-%d.npz" % step)
                tmp = os.path.join(self.model, ".tmp-model-%d.npz" % step)
                np.savez(tmp, __step=np.int64(step), __adam_t=np.int64(self.adam_t), **vals,
                         **{"__slot:" + k: v for k, v in slots.items()})
                os.replace(tmp, path)
                pattern, rx = "model-*.npz", r"model-(\d+)\.npz"
            tmpck = os.path.join(self.model, ".tmp-checkpoint")
            with open(tmpck, "w") as f:
                f.write('model_checkpoint_path: "model-%d"\n' % step)
            os.replace(tmpck, os.path.join(self.model, "checkpoint"))
            ckpts = sorted(glob.glob(os.path.join(self.model, pattern)), key=lambda p: int(re.search(rx, p).group(1)))
            for old in ckpts[:-keep] if keep > 0 else []:
                os.remove(old)
                if old.endswith(".index"):
                    data = old[:-len(".index")] + ".data-00000-of-00001"
                    if os.path.exists(data):
                        os.remove(data)
        if multi:
            dist.barrier()

    def _load_slots(self, values):
        """values: '<var><slot suffix>' -> array in the variable's (full) TF shape; each rank keeps its own columns."""
        st = self.engine.store
        for which, suffix in self._slot_names().items():
            found = {k[:-len(suffix)]: v for k, v in values.items() if k.endswith(suffix) and k[:-len(suffix)] in st.specs}
            if not found:
                continue
            st.ensure_opt_state(self.opt)
            flat = getattr(st, which)
            for name, arr in found.items():
                sp = st.specs[name]
                if not sp.trainable:
                    continue
                arr = np.asarray(arr)
                if sp.col_range is not None and tuple(arr.shape) == sp.full_shape and sp.full_shape != sp.tf_shape:
                    arr = arr[..., sp.col_range[0]:sp.col_range[1]]
                flat[sp.offset:sp.offset + sp.numel].copy_(torch.from_numpy(sp.to_internal(arr)).reshape(-1).to(flat.device))

    def load(self):
        ck = os.path.join(self.model, "checkpoint")
        if not os.path.isfile(ck):
            sys.exit("Cannot find model in %s" % self.model)
        name = re.search(r'"(.*)"', open(ck).readline()).group(1)
        step = int(next(re.finditer(r"(\d+)(?!.*\d)", name)).group(0))      # trainer.py:149-153
        st = self.engine.store
        base = os.path.join(self.model, os.path.basename(name))
        if not os.path.exists(base + ".npz") and os.path.exists(base + ".index"):
            # a checkpoint in tf.train.Saver's format (the reference's own, trainer.py:160-166): variables and optimizer
            # slots are matched by their TF names
            from ..misc.tf_checkpoint import read_tf_checkpoint
            vals = read_tf_checkpoint(base)
            st.load_tf(vals)
            self._load_slots(vals)
            if "beta1_power" in vals and self.opt == L.OPT_ADAM:
                b1p = float(np.asarray(vals["beta1_power"]))
                self.adam_t = int(round(np.log(max(b1p, 1e-300)) / np.log(0.9))) if 0 < b1p < 1 else 0
            self.global_step = step
            self.is_loaded = True
            return step
        z = np.load(base + ".npz")
        st.load_tf({k: z[k] for k in z.files if not k.startswith("__")})
        self._load_slots({k[len("__slot:"):]: z[k] for k in z.files if k.startswith("__slot:")})
        if "__opt_state1" in z.files:       # round-1 checkpoints: flat single-rank slot buffers
            st.ensure_opt_state(L.OPT_MOMENTUM)
            st.state1.copy_(torch.from_numpy(z["__opt_state1"]))
        if "__opt_state2" in z.files:
            st.ensure_opt_state(L.OPT_ADAM)
            st.state2.copy_(torch.from_numpy(z["__opt_state2"]))
        self.adam_t = int(z["__adam_t"]) if "__adam_t" in z.files else 0
        self.global_step = step
        self.is_loaded = True
        self._in_memory = True
        return step

    def reset(self):
        self.engine = set_engine(Engine())
        self.is_built = False
        self.is_loaded = False
        self.global_step = None
        self.modes = set()

    def close(self):
        torch.cuda.synchronize()
