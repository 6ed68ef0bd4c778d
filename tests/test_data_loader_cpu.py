"""Host data loader against the reference's sampling rules (dataset/data_loader.py:19-56, 229-307, 417-462) on a
synthetic Kaldi directory: batch layout, speaker / label bookkeeping, segment crops that decode bit-exactly to the rows
of the source matrices (CPU oracle of the 'CM ' format), the sequential queue visiting every segment once."""
import io
import os
import struct

import numpy as np

from oracle import kaldi_cm_oracle as KO
from tests.xv_testlib import make_kaldi_dir
from tf_kaldi_speaker_b200.dataset import data_loader as DL


def _decode_full(mats_dir_entry):
    """Whole matrix of an scp entry through the CPU oracle -> float32 [rows, cols]."""
    loc = mats_dir_entry.split(" ")[-1]
    path, off = loc.rsplit(":", 1)
    with open(path, "rb") as f:
        f.seek(int(off))
        assert f.read(2) == b"\0B" and f.read(3) == b"CM "
        return KO.read_compressed_mat(f)


def _decode_batch(batch):
    out = np.empty(batch.shape, dtype=np.float32)
    for i in range(batch.B):
        out[i] = KO.decode(batch.glob[i, 0], batch.glob[i, 1], batch.headers[i], batch.data[i])
    return out


def test_speaker_info_and_reader(tmp_path):
    data, spklist, mats = make_kaldi_dir(tmp_path, num_speakers=4, utts_per_speaker=2, dim=23)
    spk2features, features2spk, spk2index = DL.get_speaker_info(data, spklist)
    assert sorted(spk2index.values()) == [0, 1, 2, 3] and len(features2spk) == 8
    assert all(len(v) == 2 for v in spk2features.values())
    fr = DL.FeatureReader(data)
    assert fr.dim == 23 and fr.get_dim() == 23
    entry = spk2features[2][0]
    utt = entry.split(" ")[0]
    assert fr.utt2num_frames[utt] == mats[utt].shape[0]
    raw, start = fr.read_segment(entry, 50, shuffle=False)
    assert start == 0 and raw.data.shape == (23, 50)
    full = _decode_full(entry)
    seg = KO.decode(raw.globmin, raw.globrange, raw.headers, raw.data)
    assert np.array_equal(seg, full[:50])
    # quantisation error of the synthetic archive is small against the source features
    assert np.abs(full - mats[utt]).max() < 0.1 * np.abs(mats[utt]).max()


def test_random_queue_batches(tmp_path):
    data, spklist, mats = make_kaldi_dir(tmp_path, num_speakers=6, utts_per_speaker=3, dim=30, min_frames=80, max_frames=200)
    q = DL.KaldiDataRandomQueue(data, spklist, num_parallel=2, max_qsize=4, num_speakers=4, num_segments=2, min_len=60,
                                max_len=100, shuffle=True, base_seed=7)
    full = {k: _decode_full(k) for k in q.features2spk}
    q.start()
    try:
        lengths = set()
        for _ in range(6):
            batch, labels = q.fetch()
            assert batch.shape[0] == 8 and batch.shape[2] == 30 and 60 <= batch.shape[1] <= 100
            lengths.add(batch.shape[1])
            assert labels.dtype == np.int32 and labels.shape == (8,)
            # N speakers x K segments: labels come in runs of K, N distinct speakers per batch
            assert all(labels[2 * i] == labels[2 * i + 1] for i in range(4)) and len(set(labels.tolist())) == 4
            x = _decode_batch(batch)
            for i in range(8):        # every segment is a contiguous crop of some utterance of that speaker, bit-exactly
                cands = [m for k, m in full.items() if q.features2spk[k] == labels[i] and m.shape[0] > batch.shape[1]]
                hit = False
                for m in cands:
                    for s in range(m.shape[0] - batch.shape[1] + 1):
                        if np.array_equal(m[s], x[i, 0]) and np.array_equal(m[s:s + batch.shape[1]], x[i]):
                            hit = True
                            break
                    if hit:
                        break
                assert hit, (i, labels[i])
        assert len(lengths) > 1          # one length per batch, drawn anew every batch (data_loader.py:273)
    finally:
        q.stop()


def test_sequential_queue_visits_every_segment_once(tmp_path):
    data, spklist, mats = make_kaldi_dir(tmp_path, num_speakers=5, utts_per_speaker=4, dim=30, min_frames=90, max_frames=150)
    q = DL.KaldiDataSeqQueue(data, spklist, num_parallel=2, max_qsize=4, batch_size=4, min_len=70, max_len=120,
                             shuffle=False, base_seed=3)
    q.start()
    seen = 0
    labels_seen = []
    try:
        while True:
            try:
                batch, labels = q.fetch()
            except DL.DataOutOfRange:
                break
            assert batch.shape[0] == 4 and 70 <= batch.shape[1] <= 120
            seen += 4
            labels_seen += labels.tolist()
    finally:
        q.stop()
    # 20 segments over 2 workers of 10: each worker serves int(10 / 4) = 2 full batches (data_loader.py:445)
    assert seen == 16
    assert set(labels_seen) <= set(range(5))
