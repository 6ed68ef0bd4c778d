"""Embedding extraction with the reference's semantics (egs/voxceleb/v1/nnet/lib/extract.py:65-94) on the CUDA path.

Per utterance the result equals the reference's: utterances shorter than ``min_chunk_size`` frames are skipped;
utterances longer than ``chunk_size`` are cut into chunks of ``chunk_size`` with hop ``chunk_size/2`` (the last one
shorter), each chunk embedded separately, optionally L2-normalised, and averaged with chunk-length weights; the
embedding is ``endpoints[params.embedding_node]`` with BN in inference mode; optional final L2 normalisation.

What changes is the batching: the reference runs one ``[1, T, D]`` ``sess.run`` per utterance on one CPU thread
(``extract.py:90``, ``trainer.py:46-50``).  Here chunks of many utterances are packed into padded ``[N, Tmax, D]``
batches (Tmax bucketed to a multiple of 256 frames to bound the number of distinct shapes) and the length-masked
statistics pooling keeps every row independent of its padding, so batching does not change any result.
Output order = input order.
"""
import sys

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


def extract_embeddings(trainer, features, wspecifier=None, chunk_size=10000, min_chunk_size=25, normalize=False,
                       max_batch_frames=600000, log=None):
    """features: iterable of (key, np[T, D]) (e.g. ``read_mat_ark(rspecifier)``) or an rspecifier string.
    Writes Kaldi binary float vectors to ``wspecifier`` (path / fd) if given; returns [(key, embedding)]."""
    if isinstance(features, str):
        if features.rsplit(".", 1)[-1] == "scp":
            sys.exit("The rspecifier must be ark or input pipe")             # extract.py:60-62
        features = read_mat_ark(features)
    utts = []            # (key, [(start, len)])
    jobs = []            # (utt index, chunk index, np[len, D])
    for key, feat in features:
        feat = np.asarray(feat, dtype=np.float32)
        if feat.shape[0] < min_chunk_size:
            if log:
                log("[INFO] Key %s length too short, %d < %d, skip." % (key, feat.shape[0], min_chunk_size))
            continue
        chunks = split_chunks(feat.shape[0], chunk_size)
        ui = len(utts)
        utts.append((key, chunks))
        for ci, (s, l) in enumerate(chunks):
            jobs.append((ui, ci, feat[s:s + l]))
    # length-sorted packing into padded batches
    order = sorted(range(len(jobs)), key=lambda i: jobs[i][2].shape[0])
    results = {}
    i = 0
    eng = trainer.engine
    staging = None
    while i < len(order):
        tmax = _bucket(jobs[order[i]][2].shape[0])
        group = []
        while i < len(order) and _bucket(jobs[order[i]][2].shape[0]) == tmax and (len(group) + 1) * tmax <= max(max_batch_frames, tmax):
            group.append(order[i])
            i += 1
        dim = jobs[group[0]][2].shape[1]
        need = len(group) * tmax * dim
        if staging is None or staging.numel() < need:       # one pinned staging buffer, grown geometrically: async H2D
            staging = torch.empty(int(need * 1.25), dtype=torch.float32, pin_memory=torch.cuda.is_available())
        batch_t = staging[:need].view(len(group), tmax, dim)
        batch = batch_t.numpy()
        lengths = np.zeros((len(group),), dtype=np.int32)
        for r, j in enumerate(group):
            f = jobs[j][2]
            batch[r, :f.shape[0]] = f
            batch[r, f.shape[0]:] = 0.0
            lengths[r] = f.shape[0]
        emb = trainer.predict_batch_padded(batch_t, lengths)
        for r, j in enumerate(group):
            results[(jobs[j][0], jobs[j][1])] = emb[r]
        if len(eng.ws) > 400:       # bound the workspace cache when many distinct batch shapes were seen
            eng.ws.clear()
    out = []
    fd = open_or_fd(wspecifier, "wb") if wspecifier is not None else None
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
    if fd is not None and fd is not wspecifier:
        fd.close()
    return out
