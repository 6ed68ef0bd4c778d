"""Trainer with the reference's surface (model/trainer.py:17-726): build / train / valid / predict / save / load.

What changes: ``sess.run(train_op)`` (trainer.py:505-508) becomes ``train_step`` -- forward, backward, (data-parallel
gradient all-reduce,) fused optimizer and BN moving-statistics update, all hand-written sm_100a kernels enqueued
on one CUDA stream.  Host control flow (epochs, logging, checkpoints, LR bookkeeping) stays Python.  The Kaldi
feature pipeline (dataset/data_loader.py) stays on the host: ``train``/``valid`` take any object with the
``start()/fetch()/stop()`` queue protocol of KaldiDataRandomQueue (data_loader.py:310-414).
"""
import ctypes as C
import glob
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
                   declare_head_variables, margin_schedule)


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
        self.use_cuda_graph = bool(params.dict.get("cuda_graph", True))

    # ------------------------------------------------------------------ network (trainer.py:168-188)
    def entire_network(self, features, params, is_training, reuse_variables, lengths=None):
        features, endpoints = self.network(features, params, is_training, reuse_variables, lengths=lengths)
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
                declare_head_variables(eng, e, num_speakers, self.params, loss_type)
            eng.store.finalize()
            eng.store.init(int(self.params.dict.get("seed", 0)))
        elif mode != "predict" and "softmax/output/kernel" not in eng.store:
            raise L.XvError("build('predict') was called first: the head variables cannot be added afterwards; "
                            "build train/valid before predict")
        self.modes.add(mode)
        self.is_built = True

    # ------------------------------------------------------------------ one step = sess.run(train_op)
    def _to_device(self, features, labels=None):
        dev = self.engine.device
        if not torch.is_tensor(features):
            features = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
        features = features.to(dev, dtype=torch.float32, non_blocking=True)
        if labels is not None:
            if not torch.is_tensor(labels):
                labels = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32))
            labels = labels.to(dev, dtype=torch.int32, non_blocking=True)
        return features, labels

    def _fwd_bwd(self, features, labels, l2_loss=True, backward_part=None):
        """Forward + backward of one device-resident batch; gradients are left in the flat gradient buffer.
        ``l2_loss=False``: the regularisation loss is left to the optimizer kernel (same pass over the parameters)."""
        eng = self.engine
        eng.begin_step(True)
        eng.prefetch_head_weights("softmax/output/kernel", normalize=(self.loss_type != "softmax"))
        if self._h2d_event is not None:
            self._h2d_event.wait(torch.cuda.current_stream())     # this step's batch has arrived (see _static_batch)
        out, endpoints = self.entire_network(features, self.params, True, True)
        loss, endpoints_loss = self.loss_network(out, labels, self.num_speakers, self.params, True, True)
        endpoints.update(endpoints_loss)
        self.endpoints = endpoints
        if l2_loss:
            eng.l2_loss()
        eng.backward(backward_part)

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
        # Host batches are uploaded on a copy stream: the step waits for them only in front of the first kernel that reads
        # the features (an external event-wait node inside the captured graph, see _fwd_bwd), so the gradient-buffer fill,
        # the scalar feed and the head weight preparation at the start of the step overlap the PCIe transfer.
        cs = self._copy_stream if from_host else main
        if from_host:
            cs.wait_stream(main)            # the previous step has finished reading the static buffers
        with torch.cuda.stream(cs):
            if hasattr(features, "decode_into"):
                # dataset.feeder.CompressedSegmentBatch: H2D of the raw uint8 crops + on-device dequantise / transpose
                features.decode_into(st["x"])
                self.engine.launches += 1
            else:
                st["x"].copy_(features, non_blocking=True)       # H2D (or D2D) of this step's batch
            st["y"].copy_(labels.to(torch.int32), non_blocking=True)
            self._h2d_event.record(cs)
        return st

    def train_step(self, features, labels, learning_rate, global_step=None, fetch_loss=False):
        """The hot-loop body = sess.run(train_op) (trainer.py:491-508): forward, backward, [all-reduce], optimizer,
        BN moving statistics.  After two eager warm-up calls per batch shape the whole step is captured into a CUDA
        graph and replayed; learning rate / margin schedule live in device scalars set before each replay.
        Returns {"loss": total, "raw_loss": loss} when fetch_loss (one 16-byte D2H read), else None."""
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

        # dp_overlap: all-reduce the [tdnn6 .. head] gradient bucket while the frame-level backward runs.  Measured on
        # 2 x B200: 1.169 ms/step with, 1.160 ms without (the 39 MB exchange costs ~75 us either way and the NCCL CTAs
        # take SMs from the persistent GEMMs), so it stays opt-in.
        overlap = self.dp is not None and bool(self.params.dict.get("dp_overlap", False))

        def part_a():
            self._fwd_bwd(st["x"], st["y"], l2_loss=False, backward_part=("head" if overlap else None))

        def part_a2():
            # the head-bucket all-reduce is in flight: cap the persistent GEMM grids so that NCCL's CTAs do not strand
            # GEMM CTAs behind them (xv_gemm_set_cta_limit); the cap is baked into the captured launches
            reserve = int(self.params.dict.get("dp_overlap_reserve_sms", 16))
            L.check(eng.lib.xv_gemm_set_cta_limit(max(eng.num_sms - reserve, 2)))
            try:
                eng.backward("trunk")
            finally:
                L.check(eng.lib.xv_gemm_set_cta_limit(0))

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

        # the multimem all-reduce is a plain kernel with in-kernel rank barriers: the whole data-parallel step (forward,
        # backward, gradient exchange, optimizer) is then ONE captured graph, like the single-GPU step
        in_graph_exchange = self.dp is not None and bool(getattr(self.dp, "graph_safe", False)) and not overlap
        # dp_bucket_overlap: the [tdnn6 .. head] gradients are complete a third of the way into the backward pass; their
        # exchange runs as a small-grid kernel on a second stream beside the tdnn5 backward (whose two GEMMs leave it a
        # few SMs), only the [tdnn1 .. tdnn5] bucket is exchanged after the backward pass
        bucket_overlap = (in_graph_exchange and bool(self.params.dict.get("dp_bucket_overlap", False)) and not clip
                          and getattr(self.dp, "_mm", None) is not None and self.dp.split > 0
                          and self.dp.grad_dtype != "bf16")

        def part_ab_overlapped():
            ex_sms = int(self.params.dict.get("dp_bucket_overlap_sms", 16))
            self._fwd_bwd(st["x"], st["y"], l2_loss=False, backward_part="head")
            with eng.fork_exchange_stream():
                self.dp.allreduce_range(self.dp.split, self.dp.dp_numel, grid=ex_sms)
            eng.cap_next_gemms(int(self.params.dict.get("dp_bucket_overlap_gemms", 2)), eng.num_sms - ex_sms)
            try:
                eng.backward("trunk")
            finally:
                eng.cap_next_gemms(0, 0)
            eng.join_exchange_stream()
            self.dp.allreduce_range(0, self.dp.split)
            part_b()

        def run(ga, ga2, gb):
            """forward + head backward | all-reduce(head bucket) overlapping the frame-level backward | all-reduce(trunk
            bucket) | optimizer.  ga / ga2 / gb are captured graphs or None (eager)."""
            if bucket_overlap:
                ga.replay() if ga is not None else part_ab_overlapped()
                return
            ga.replay() if ga is not None else part_a()
            if self.dp is None or in_graph_exchange:
                if ga is None:
                    if in_graph_exchange:
                        self.dp.allreduce_gradients()
                    part_b()
                return
            if overlap:
                self.dp.allreduce_bucket_async("head")
                ga2.replay() if ga2 is not None else part_a2()
                self.dp.allreduce_bucket_async("trunk")
                self.dp.wait_all()
            else:
                self.dp.allreduce_gradients()
            gb.replay() if gb is not None else part_b()

        if st["graphs"] is not None:
            run(*st["graphs"])
            eng.launches += st["launches"]
        else:
            st["calls"] += 1
            if self.use_cuda_graph and st["calls"] > 2:
                l0 = eng.launches
                eng.capturing = True
                try:
                    ga = torch.cuda.CUDAGraph()
                    ga2 = gb = None
                    with torch.cuda.graph(ga):
                        if bucket_overlap:
                            part_ab_overlapped()
                        else:
                            part_a()
                            if self.dp is None:
                                part_b()
                            elif in_graph_exchange:
                                self.dp.allreduce_gradients()
                                part_b()
                    if self.dp is not None and not in_graph_exchange:
                        if overlap:
                            ga2 = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(ga2):
                                part_a2()
                        gb = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gb):
                            part_b()
                finally:
                    eng.capturing = False
                st["launches"] = eng.launches - l0
                st["graphs"] = (ga, ga2, gb)
                run(ga, ga2, gb)
            else:
                run(None, None, None)
        return self._finish_step(global_step, fetch_loss)

    def _finish_step(self, global_step, fetch_loss):
        eng = self.engine
        self.global_step = int(global_step) + 1
        if fetch_loss:
            vals = eng.scalars[:5].tolist()          # one D2H read
            raw = vals[0]
            l2 = vals[1]
            if self.dp is not None:
                raw = self.dp.mean_scalar(raw)
                if eng.head_shard is not None:       # add the regularisation loss of the other ranks' head columns
                    l2 += self.dp.sum_scalar(vals[4]) - vals[4]
            self.train_ops = {"raw_loss": raw, "loss": raw + l2 + vals[3]}
            return dict(self.train_ops)
        return None

    def train(self, data, spklist, learning_rate, aux_data=None):
        """One epoch (trainer.py:451-520).  ``data``: object with start()/fetch()/stop() yielding (features, labels)."""
        assert "train" in self.modes
        if not hasattr(data, "fetch"):
            raise NotImplementedError("The Kaldi feature pipeline stays on the host: pass a loader object with "
                                      "start()/fetch()/stop() (e.g. dataset.data_loader.KaldiDataRandomQueue).")
        curr_step = self.global_step or 0
        if hasattr(data, "start"):
            data.start()
        steps = int(self.params.num_steps_per_epoch)
        t0 = time.time()
        for step in range(curr_step % steps, steps):
            show = (step % int(self.params.show_training_progress) == 0)
            features, labels = data.fetch()
            res = self.train_step(features, labels, learning_rate, curr_step, fetch_loss=show)
            if show:
                dt = time.time() - t0
                t0 = time.time()
                print("Epoch: [%2d] step: [%2d/%2d] time: %.4f s/step, raw loss: %f, total loss: %f"
                      % (curr_step // steps, step, steps, dt / int(self.params.show_training_progress),
                         res["raw_loss"], res["loss"]))
            if curr_step % int(self.params.save_checkpoints_steps) == 0 and curr_step != 0:
                self.save(curr_step)
            curr_step += 1
        self.save(curr_step)
        if hasattr(data, "stop"):
            data.stop()
        return

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

    def valid_step(self, features, labels):
        """Loss of the validation graph (is_training=False, margins neutralised) and the output embeddings."""
        eng = self.engine
        features, labels = self._to_device(features, labels)
        vp = self._valid_params()
        vp.dict["global_step"] = self.global_step or 0
        eng.begin_step(False)
        out, endpoints = self.entire_network(features, vp, False, True)
        loss, _ = self.loss_network(out, labels, self.num_speakers, vp, False, True)
        self.endpoints = endpoints
        return float(loss.item()), endpoints["output"].dense()

    def valid(self, data, spklist, batch_type="softmax", output_embeddings=False, aux_data=None):
        assert "valid" in self.modes or "train" in self.modes
        if not hasattr(data, "fetch"):
            raise NotImplementedError("pass a loader object with start()/fetch()/stop()")
        if hasattr(data, "start"):
            data.start()
        losses, embs, labs = [], [], []
        for _ in range(int(self.params.valid_max_iterations)):
            try:
                features, labels = data.fetch()
            except Exception:
                break
            l, e = self.valid_step(features, labels)
            losses.append(l)
            if output_embeddings:
                embs.append(e.cpu().numpy())
                labs.append(np.asarray(labels))
        if hasattr(data, "stop"):
            data.stop()
        loss = float(np.mean(losses)) if losses else 0.0
        if output_embeddings:
            return loss, np.concatenate(embs, 0), np.concatenate(labs, 0)
        return loss, None, None

    # ------------------------------------------------------------------ prediction (trainer.py:708-726)
    def predict(self, features):
        """features: np [T, D] or [N, T, D] -> embeddings np [E] or [N, E] from endpoints[embedding_node]."""
        features = np.asarray(features, dtype=np.float32)
        rank = features.ndim
        if rank == 2:
            features = features[None]
        emb = self.predict_batch_padded(features, None)
        return emb[0] if rank == 2 else emb

    def predict_batch_padded(self, features, lengths):
        """[N, Tmax, D] (+ optional lengths [N]) -> np [N, E]; rows are independent (BN in inference mode, masked
        pooling), so a ragged batch gives the same result as one call per utterance (extract.py:90)."""
        eng = self.engine
        feats, _ = self._to_device(features)
        ln = None if lengths is None else torch.as_tensor(np.asarray(lengths), dtype=torch.int32, device=eng.device)
        eng.begin_step(False)
        _, endpoints = self.entire_network(feats, self.params, False, True, lengths=ln)
        self.endpoints = endpoints
        node = endpoints[self.params.embedding_node]
        return node.dense().float().cpu().numpy()

    # ------------------------------------------------------------------ checkpoints (trainer.py:142-166)
    def save(self, step):
        os.makedirs(self.model, exist_ok=True)
        st = self.engine.store
        vals = st.export_tf()
        sh = self.engine.head_shard
        if sh is not None:          # every rank calls save(); the sharded variables are written in their full TF shape
            for name, spec in st.specs.items():
                if spec.col_range is not None:
                    loc = torch.from_numpy(vals[name]).to(self.engine.device)
                    full = sh.gather_columns(loc.reshape(-1, loc.shape[-1]))
                    vals[name] = full.reshape(spec.full_shape).cpu().numpy()
            if sh.rank != 0:
                return
        if str(self.params.dict.get("checkpoint_format", "npz")) == "tf":
            # TF tensor-bundle files under the reference's names (model-<step>.index / .data-00000-of-00001)
            from ..misc.tf_checkpoint import write_tf_checkpoint
            vals["global_step"] = np.array(step, dtype=np.int64)
            write_tf_checkpoint(os.path.join(self.model, "model-%d" % step), vals)
            with open(os.path.join(self.model, "checkpoint"), "w") as f:
                f.write('model_checkpoint_path: "model-%d"\n' % step)
            return
        path = os.path.join(self.model, "model-%d.npz" % step)
        extra = {}
        if st.state1 is not None:
            extra["__opt_state1"] = st.state1.cpu().numpy()
        if st.state2 is not None:
            extra["__opt_state2"] = st.state2.cpu().numpy()
        np.savez(path, __step=np.int64(step), __adam_t=np.int64(self.adam_t), **vals, **extra)
        with open(os.path.join(self.model, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "model-%d"\n' % step)
        keep = int(self.params.dict.get("keep_checkpoint_max", 5))
        ckpts = sorted(glob.glob(os.path.join(self.model, "model-*.npz")),
                       key=lambda p: int(re.search(r"model-(\d+)\.npz", p).group(1)))
        for old in ckpts[:-keep] if keep > 0 else []:
            os.remove(old)

    def load(self):
        ck = os.path.join(self.model, "checkpoint")
        if not os.path.isfile(ck):
            sys.exit("Cannot find model in %s" % self.model)
        name = re.search(r'"(.*)"', open(ck).readline()).group(1)
        step = int(next(re.finditer(r"(\d+)(?!.*\d)", name)).group(0))      # trainer.py:149-153
        st = self.engine.store
        base = os.path.join(self.model, os.path.basename(name))
        if not os.path.exists(base + ".npz") and os.path.exists(base + ".index"):
            # a checkpoint written by tf.train.Saver (the reference's own format, trainer.py:160-166): variables are
            # matched by their TF names; optimizer slots and global_step are ignored
            from ..misc.tf_checkpoint import read_tf_checkpoint
            st.load_tf(read_tf_checkpoint(base))
            self.global_step = step
            self.is_loaded = True
            return step
        z = np.load(base + ".npz")
        st.load_tf({k: z[k] for k in z.files if not k.startswith("__")})
        if "__opt_state1" in z.files:
            st.ensure_opt_state(L.OPT_MOMENTUM)
            st.state1.copy_(torch.from_numpy(z["__opt_state1"]))
        if "__opt_state2" in z.files:
            st.ensure_opt_state(L.OPT_ADAM)
            st.state2.copy_(torch.from_numpy(z["__opt_state2"]))
        self.adam_t = int(z["__adam_t"]) if "__adam_t" in z.files else 0
        self.global_step = step
        self.is_loaded = True
        return step

    def reset(self):
        self.engine = set_engine(Engine())
        self.is_built = False
        self.is_loaded = False
        self.global_step = None
        self.modes = set()

    def close(self):
        torch.cuda.synchronize()
