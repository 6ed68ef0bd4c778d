"""Embedding extraction with the reference's semantics (egs/voxceleb/v1/nnet/lib/extract.py:11-96) on the CUDA path.

Per utterance the result equals the reference's: utterances shorter than ``min_chunk_size`` frames are skipped;
utterances longer than ``chunk_size`` are cut into chunks of ``chunk_size`` with hop ``chunk_size/2`` (the last one
shorter), each chunk embedded separately, optionally L2-normalised, and averaged with chunk-length weights; the
embedding is ``endpoints[params.embedding_node]`` with BN in inference mode; optional final L2 normalisation.

What changes is the batching: the reference runs one ``[1, T, D]`` ``sess.run`` per utterance on one CPU thread
(``extract.py:90``, ``trainer.py:46-50``).  Here the input is STREAMED in bounded windows (``window_utts`` utterances /
``window_frames`` frames: memory stays O(window), output is written window by window in input order).  With statistics
pooling the chunks of a batch are simply CONCATENATED in one flat row space (``Trainer.predict_ragged``): frame layers in
inference mode are row-local apart from the temporal taps, a valid frame never reads rows of a neighbouring utterance,
and the pooling kernel is given each utterance's (start, length) -- no padding, no length sorting, every batch as large
as ``max_batch_frames`` allows.  (Attention pooling keeps the padded, length-bucketed ``[N, Tmax, D]`` batches with
length masks.)  Packer threads fill pinned staging buffers for the next batches while the GPU runs the current one;
neither layout changes any result.

Several GPUs: the reference fans extraction out as ``nj`` independent jobs over a split data directory
(``run_extract_embeddings.sh:68-71``); the same works here (one process per GPU, ``--gpu JOB``).  Under ``torchrun`` the
ranks shard ONE input stream instead: rank r takes utterances r, r + N, ... and writes ``<wspecifier>.<r>`` (or
substitutes a literal ``JOB`` in the specifiers by r + 1).

    python -m tf_kaldi_speaker_b200.extract [-g GPU] [-m MIN] [-s CHUNK] [-n] [--node NODE] model_dir rspecifier wspecifier
"""
import argparse
import collections
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .dataset.kaldi_io import open_or_fd, read_mat_ark, write_vec_flt


def split_chunks(num_frames, chunk_size):
    """[(start, length)] exactly as extract.py:69-80 (py2 integer division for the hop)."""
    if num_frames <= chunk_size:
        return [(0, num_frames)]
    half = chunk_size // 2
    n = int(np.ceil(float(num_frames - chunk_size) / half)) + 1
    out = []
    for i in range(n):
        start = i * half
        out.append((start, chunk_size if num_frames - start > chunk_size else num_frames - start))
    return out


def _bucket(t, q=256):
    return max(q, (t + q - 1) // q * q)


_COPY_THREADS = int(os.environ.get("XV_EXTRACT_COPY_THREADS", "6"))      # packer sub-copies (NumPy releases the GIL)
_copy_executor = None


def _copy_pool():
    global _copy_executor
    if _copy_executor is None:
        _copy_executor = ThreadPoolExecutor(max_workers=_COPY_THREADS)
    return _copy_executor


class _Staging(object):
    """A ring of pinned host buffers; a slot is refilled only after the H2D copy that read it has completed."""

    def __init__(self, depth=3):
        self.slots = [{"buf": None, "event": None} for _ in range(depth)]
        self.next = 0
        self.lock = threading.Lock()

    def take(self):
        with self.lock:
            s = self.slots[self.next % len(self.slots)]
            self.next += 1
        return s


def _pack(slot, group, jobs, tmax, dim):
    """Packer thread: the chunks of one batch -> slot's pinned [N, tmax, dim] (zero padded) + lengths."""
    if slot["event"] is not None:
        slot["event"].synchronize()
    need = len(group) * tmax * dim
    if slot["buf"] is None or slot["buf"].numel() < need:
        slot["buf"] = torch.empty(int(need * 1.25), dtype=torch.float32, pin_memory=torch.cuda.is_available())
    batch_t = slot["buf"][:need].view(len(group), tmax, dim)
    batch = batch_t.numpy()
    lengths = np.zeros((len(group),), dtype=np.int32)
    for r, j in enumerate(group):
        f = jobs[j][2]
        n = f.shape[0]
        batch[r, :n] = f
        batch[r, n:] = 0.0
        lengths[r] = n
    return batch_t, lengths


def _pack_ragged(slot, group, jobs, dim):
    """Packer thread, padding-free layout: the chunks of one batch back to back in slot's pinned [R, dim]."""
    if slot["event"] is not None:
        slot["event"].synchronize()
    lengths = np.array([jobs[j][2].shape[0] for j in group], dtype=np.int32)
    starts = np.zeros_like(lengths)
    starts[1:] = np.cumsum(lengths)[:-1]
    rows = int(lengths.sum())
    need = rows * dim
    if slot["buf"] is None or slot["buf"].numel() < need:
        slot["buf"] = torch.empty(int(need * 1.25), dtype=torch.float32, pin_memory=torch.cuda.is_available())
    flat_t = slot["buf"][:need].view(rows, dim)
    flat = flat_t.numpy()

    def copy_range(lo, hi):          # NumPy releases the GIL inside large slice copies: the sub-copies run in parallel
        for k in range(lo, hi):
            flat[starts[k]:starts[k] + lengths[k]] = jobs[group[k]][2]
    n = len(group)
    if rows * dim * 4 < (8 << 20) or n < 2 * _COPY_THREADS:
        copy_range(0, n)
    else:
        # contiguous utterance ranges of roughly equal frame counts
        cum = np.concatenate([[0], np.cumsum(lengths)])
        cuts = [int(np.searchsorted(cum, rows * t / _COPY_THREADS)) for t in range(_COPY_THREADS + 1)]
        cuts[0], cuts[-1] = 0, n
        list(_copy_pool().map(lambda ab: copy_range(*ab), [(cuts[t], cuts[t + 1]) for t in range(_COPY_THREADS) if cuts[t + 1] > cuts[t]]))
    return flat_t, (starts, lengths)


def _run_window(trainer, utts, jobs, max_batch_frames, normalize, fd, out, staging, pool):
    """Embed the chunks of one window and emit its utterances in input order."""
    ragged = trainer.params.pooling_type == "statistics_pooling"
    groups = []
    if ragged:
        # statistics pooling: utterances are CONCATENATED (Trainer.predict_ragged) -- no padding, no length sorting, every
        # batch as large as max_batch_frames allows
        group, rows = [], 0
        for i in range(len(jobs)):
            n = jobs[i][2].shape[0]
            if group and rows + n > max_batch_frames:
                groups.append((0, group))
                group, rows = [], 0
            group.append(i)
            rows += n
        if group:
            groups.append((0, group))
    else:
        order = sorted(range(len(jobs)), key=lambda i: jobs[i][2].shape[0])
        i = 0
        while i < len(order):
            tmax = _bucket(jobs[order[i]][2].shape[0])
            group = []
            while i < len(order) and _bucket(jobs[order[i]][2].shape[0]) == tmax and \
                    (len(group) + 1) * tmax <= max(max_batch_frames, tmax):
                group.append(order[i])
                i += 1
            groups.append((tmax, group))
    results = {}
    pending = collections.deque()      # packed batches waiting for the GPU
    done = collections.deque()         # (group, device embeddings) whose D2H read is deferred by one batch

    def launch(item):
        (tmax, group), slot, fut = item
        batch_t, lengths = fut.result()
        if ragged:
            emb = trainer.predict_ragged(batch_t, lengths[0], lengths[1], as_device=True)
        else:
            emb = trainer.predict_batch_padded(batch_t, lengths, as_device=True)
        if torch.cuda.is_available():
            if slot["event"] is None:
                slot["event"] = torch.cuda.Event()
            slot["event"].record()
        done.append((group, emb))
        while len(done) > 1:
            collect(done.popleft())

    def collect(item):
        group, emb = item
        e = emb.cpu().numpy()
        for r, j in enumerate(group):
            results[(jobs[j][0], jobs[j][1])] = e[r]

    for g in groups:
        slot = staging.take()
        dim = jobs[g[1][0]][2].shape[1]
        if ragged:
            pending.append((g, slot, pool.submit(_pack_ragged, slot, g[1], jobs, dim)))
        else:
            pending.append((g, slot, pool.submit(_pack, slot, g[1], jobs, g[0], dim)))
        if len(pending) >= len(staging.slots) - 1:
            launch(pending.popleft())
    while pending:
        launch(pending.popleft())
    while done:
        collect(done.popleft())
    for ui, (key, chunks) in enumerate(utts):
        embs = np.stack([results[(ui, ci)] for ci in range(len(chunks))]).astype(np.float32)
        if len(chunks) > 1:
            if normalize:
                embs = embs / np.sqrt(np.sum(np.square(embs), axis=1, keepdims=True))
            ln = np.array([l for _, l in chunks], dtype=np.float32)[:, None]
            e = np.sum(embs * ln, axis=0) / np.sum(ln)
        else:
            e = embs[0]
        if normalize:
            e = e / np.sqrt(np.sum(np.square(e)))
        e = e.astype(np.float32)
        out.append((key, e))
        if fd is not None:
            write_vec_flt(fd, e, key=key)


def extract_embeddings(trainer, features, wspecifier=None, chunk_size=10000, min_chunk_size=25, normalize=False,
                       max_batch_frames=600000, log=None, window_utts=2048, window_frames=6000000, shard=(0, 1)):
    """features: iterable of (key, np[T, D]) (e.g. ``read_mat_ark(rspecifier)``) or an rspecifier string.
    Writes Kaldi binary float vectors to ``wspecifier`` (path / fd) if given; returns [(key, embedding)].
    ``shard = (rank, world)``: this process takes utterances rank, rank + world, ... of the stream."""
    if isinstance(features, str):
        if features.rsplit(".", 1)[-1] == "scp":
            sys.exit("The rspecifier must be ark or input pipe")             # extract.py:60-62
        features = read_mat_ark(features)
    rank, world = shard
    out = []
    fd = open_or_fd(wspecifier, "wb") if wspecifier is not None else None
    staging = _Staging()
    utts, jobs, frames = [], [], 0            # (key, [(start, len)]); (utt index, chunk index, np[len, D])
    with ThreadPoolExecutor(max_workers=2) as pool:
        for index, (key, feat) in enumerate(features):
            if index % world != rank:
                continue
            feat = np.asarray(feat, dtype=np.float32)
            if feat.shape[0] < min_chunk_size:
                if log:
                    log("[INFO] Key %s length too short, %d < %d, skip." % (key, feat.shape[0], min_chunk_size))
                continue
            chunks = split_chunks(feat.shape[0], chunk_size)
            if log and len(chunks) > 1:
                log("[INFO] Key %s length %d > %d, split to %d segments." % (key, feat.shape[0], chunk_size, len(chunks)))
            ui = len(utts)
            utts.append((key, chunks))
            for ci, (s, l) in enumerate(chunks):
                jobs.append((ui, ci, feat[s:s + l]))
            frames += feat.shape[0]
            if len(utts) >= window_utts or frames >= window_frames:
                _run_window(trainer, utts, jobs, max_batch_frames, normalize, fd, out, staging, pool)
                utts, jobs, frames = [], [], 0
        if utts:
            _run_window(trainer, utts, jobs, max_batch_frames, normalize, fd, out, staging, pool)
    if fd is not None and fd is not wspecifier:
        fd.close()
    return out


def main(argv=None):
    """Command line of egs/voxceleb/v1/nnet/lib/extract.py:11-22 (what wrap/extract_wrapper.sh invokes)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("-g", "--gpu", type=int, default=-1,
                        help="The GPU id (-1: the default CUDA device; this implementation has no CPU mode).")
    parser.add_argument("-m", "--min-chunk-size", type=int, default=25,
                        help="The minimum length of the segments. Any segment shorted than this value will be ignored.")
    parser.add_argument("-s", "--chunk-size", type=int, default=10000,
                        help="The length of the segments used to extract the embeddings. Segments longer than this value "
                             "will be splited before extraction. Then the splited embeddings will be averaged to get the "
                             "final embedding. L2 normalizaion will be applied before the averaging if specified.")
    parser.add_argument("-n", "--normalize", action="store_true", help="Normalize the embedding before averaging and output.")
    parser.add_argument("--node", type=str, default="", help="The node to output the embeddings.")
    parser.add_argument("model_dir", type=str, help="The model directory.")
    parser.add_argument("rspecifier", type=str, help="Kaldi feature rspecifier (or ark file).")
    parser.add_argument("wspecifier", type=str, help="Kaldi output wspecifier (or ark file).")
    args = parser.parse_args(argv)

    from .misc.utils import Params
    from .model.trainer import Trainer
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        sys.exit("xvector_b200 extraction needs a CUDA device (there is no CPU mode)")
    ndev = torch.cuda.device_count()
    torch.cuda.set_device((args.gpu if args.gpu >= 0 else int(os.environ.get("LOCAL_RANK", rank))) % ndev)

    nnet_dir = os.path.join(args.model_dir, "nnet")
    config_json = os.path.join(nnet_dir, "config.json")
    if not os.path.isfile(config_json):
        sys.exit("Cannot find params.json in %s" % config_json)
    params = Params(config_json)
    if len(args.node) != 0:
        params.embedding_node = args.node
    print("Extract embedding from %s" % params.embedding_node, file=sys.stderr)
    trainer = Trainer(params, args.model_dir, single_cpu=True)
    with open(os.path.join(nnet_dir, "feature_dim"), "r") as f:
        dim = int(f.readline().strip())
    trainer.build("predict", dim=dim)
    trainer.load()

    rspec, wspec, shard = args.rspecifier, args.wspecifier, (0, 1)
    if world > 1:
        if "JOB" in rspec:          # Kaldi's run.pl convention: per-job inputs / outputs
            rspec = rspec.replace("JOB", str(rank + 1))
        else:
            shard = (rank, world)
        wspec = wspec.replace("JOB", str(rank + 1)) if "JOB" in wspec else "%s.%d" % (wspec, rank)
    extract_embeddings(trainer, rspec, wspec, chunk_size=args.chunk_size, min_chunk_size=args.min_chunk_size,
                       normalize=args.normalize, log=lambda m: print(m, file=sys.stderr), shard=shard)
    trainer.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
