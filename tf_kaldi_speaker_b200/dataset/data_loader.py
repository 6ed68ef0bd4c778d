"""Host-side Kaldi data loader with the reference's surface (dataset/data_loader.py:19-56, 229-414, 417-560).

``KaldiDataRandomQueue`` (training: N speakers x K segments per batch, one random length per batch) and
``KaldiDataSeqQueue`` (validation: every segment once) keep the reference's constructor arguments, ``set_batch`` /
``set_length`` / ``start`` / ``fetch`` / ``stop`` protocol, sampling rules and ``DataOutOfRange`` -- so
``Trainer.train(data_dir, spklist, learning_rate)`` builds them exactly like model/trainer.py:472-480 does.

What differs is what travels: the reference's loader processes dequantise every compressed ('CM ') segment with NumPy
and put float32 ``[B, T, D]`` on the queue; here a worker gathers the RAW uint8 crop of every segment (the bytes
``_read_compressed_submat`` reads, dataset/kaldi_io.py:814-868) and ``fetch()`` hands the trainer a
``CompressedSegmentBatch`` in pinned memory: 4x fewer bytes through the queue and over PCIe, dequantised + transposed on
the GPU by ``xv_cm_decode`` (bit-exact with the reference reader, tests/test_cm_decode_gpu.py).  The sampling itself --
``random.Random`` per worker, speakers without a long-enough utterance re-drawn (data_loader.py:273-295) -- is restated
line by line; the py2-only ``rd.jumpahead(seed)`` (data_loader.py:262) becomes a seed offset.
"""
import os
import random
import struct
import time
import multiprocessing

import numpy as np

from .kaldi_io import CompressedFeatureReader


# Worker processes are forked, like the reference's (data_loader.py:381-397): the children only read archives and run
# NumPy, and a fork shares the (large) speaker -> segments tables copy-on-write.  XV_LOADER_START_METHOD=spawn|forkserver
# trades that for a clean child (the tables are then pickled to every worker).
_MP = multiprocessing.get_context(os.environ.get("XV_LOADER_START_METHOD", "fork"))
Event, Process, Queue = _MP.Event, _MP.Process, _MP.Queue


class DataOutOfRange(Exception):
    pass


def get_speaker_info(data, spklist):
    """data_loader.py:19-56: -> (spk2features {spk index: ['utt path:offset']}, features2spk, spk2index)."""
    assert (os.path.isdir(data) and os.path.isfile(spklist))
    spk2index = {}
    with open(spklist, "r") as f:
        for line in f.readlines():
            spk, index = line.strip().split(" ")
            spk2index[spk] = int(index)
    utt2spk = {}
    with open(os.path.join(data, "spk2utt"), "r") as f:
        for line in f.readlines():
            spk, utts = line.strip().split(" ", 1)
            for utt in utts.split(" "):
                utt2spk[utt] = spk2index[spk]
    spk2features = {}
    features2spk = {}
    with open(os.path.join(data, "feats.scp"), "r") as f:
        for line in f.readlines():
            (key, rxfile) = line.strip().split(" ")
            spk = utt2spk[key]
            if spk not in spk2features:
                spk2features[spk] = []
            spk2features[spk].append(key + " " + rxfile)
            features2spk[key + " " + rxfile] = spk
    return spk2features, features2spk, spk2index


class FeatureReader(object):
    """dataset/kaldi_io.py:27-149 for the training path: ``utt2num_frames``, ``dim`` and segment reads from
    ``path:offset`` entries -- returning the undecoded crop (CompressedRaw) instead of a float matrix."""

    def __init__(self, data):
        self.data = data
        self._raw = CompressedFeatureReader()
        self.utt2num_frames = {}
        assert os.path.exists(os.path.join(data, "utt2num_frames")), "[Error] Expect utt2num_frames exists in %s " % data
        with open(os.path.join(data, "utt2num_frames"), "r") as f:
            for line in f.readlines():
                utt, length = line.strip().split(" ")
                self.utt2num_frames[utt] = int(length)
        self.dim = self.get_dim()

    def get_dim(self):
        with open(os.path.join(self.data, "feats.scp"), "r") as f:
            loc = f.readline().strip().split(" ")[-1]
        filename, offset = loc.rsplit(":", 1)
        with open(filename, "rb") as fd:
            fd.seek(int(offset))
            if fd.read(2) != b"\0B":
                raise IOError("Cannot read features from %s" % loc)
            header = fd.read(3).decode()
            if header == "CM ":
                return struct.unpack("<ffii", fd.read(16))[3]
            if header in ("FM ", "DM "):
                return struct.unpack("<bibi", fd.read(10))[3]
            raise IOError("unknown matrix header '%s' in %s" % (header, loc))

    def close(self):
        self._raw.close()

    def read_segment(self, file_or_fd, length=None, shuffle=False, start=None, rd=random):
        """kaldi_io.py:112-149: (raw crop, start).  ``length`` longer than the utterance is clamped; ``shuffle`` draws the
        start frame uniformly."""
        utt = file_or_fd.split(" ")[0]
        if length is not None and start is None:
            num_features = self.utt2num_frames[utt]
            length = num_features if length > num_features else length
            start = rd.randint(0, num_features - length) if shuffle else 0
        elif length is not None:
            assert not shuffle, "The start point is specified, thus shuffling is invalid."
        return self._raw.read_segment(file_or_fd, length, start), start


def _pack(raws, labels):
    """[CompressedRaw] -> the arrays of a CompressedSegmentBatch (what goes through the process queue)."""
    B, D, T = len(raws), raws[0].cols, raws[0].data.shape[1]
    data = np.empty((B, D, T), dtype=np.uint8)
    headers = np.empty((B, D, 4), dtype=np.uint16)
    glob = np.empty((B, 2), dtype=np.float32)
    for i, r in enumerate(raws):
        data[i], headers[i], glob[i, 0], glob[i, 1] = r.data, r.headers, r.globmin, r.globrange
    return data, headers, glob, np.asarray(labels, dtype=np.int32)


def _make_rng(seed, base_seed):
    if base_seed is None:
        return random.Random(int.from_bytes(os.urandom(4), "little") + int(seed))     # data_loader.py:261-262
    return random.Random(int(base_seed) * 1000003 + int(seed))                       # reproducible runs / tests


def batch_random(stop_event, queue, data, spk2features, num_total_speakers, num_speakers=10, num_segments=10,
                 min_len=200, max_len=400, shuffle=True, seed=0, base_seed=None):
    """Worker of KaldiDataRandomQueue (data_loader.py:229-307)."""
    rd = _make_rng(seed, base_seed)
    feature_reader = FeatureReader(data)
    speakers = list(spk2features.keys())
    if num_total_speakers < num_speakers:
        print("[Warning] The number of available speakers are less than the required speaker. Some speakers will be duplicated.")
        speakers = speakers * (int(num_speakers / num_total_speakers) + 1)
    while not stop_event.is_set():
        batch_speakers = rd.sample(speakers, num_speakers)
        batch_length = rd.randint(min_len, max_len)
        raws, labels = [], []
        for i, speaker in enumerate(batch_speakers):
            # The length may be larger than the utterance length: a speaker without a long-enough utterance is re-drawn
            feature_list = []
            spk = speaker
            while len(feature_list) == 0:
                feature_list = []
                for feat in spk2features[spk]:
                    if feature_reader.utt2num_frames[feat.split(" ")[0]] > batch_length:
                        feature_list.append(feat)
                if len(feature_list) == 0:
                    spk = rd.choice(list(set(speakers) - set(batch_speakers)))
                    batch_speakers[i] = spk
            if len(feature_list) < num_segments:
                feature_list = feature_list * (int(num_segments / len(feature_list)) + 1)
            speaker_features = rd.sample(feature_list, num_segments)
            for feat in speaker_features:
                raw, _ = feature_reader.read_segment(feat, batch_length, shuffle=shuffle, rd=rd)
                raws.append(raw)
                labels.append(spk)
        queue.put(_pack(raws, labels))
    time.sleep(0.2)
    while not queue.empty():
        try:
            queue.get(block=False)
        except Exception:
            pass
    return


class _PinnedRing(object):
    """CompressedSegmentBatch objects in pinned memory, re-used round robin per batch shape; a slot is refilled only after
    the H2D copy that read it has completed (the host runs several CUDA-graph replays ahead of the device)."""

    def __init__(self, depth=4):
        self.depth = depth
        self.slots = {}

    def wrap(self, data, headers, glob):
        from .feeder import CompressedSegmentBatch
        B, D, T = data.shape
        ring = self.slots.setdefault((B, T, D), [[], 0])
        if len(ring[0]) < self.depth:
            ring[0].append(CompressedSegmentBatch(B, T, D))
        slot = ring[0][ring[1] % len(ring[0])]
        ring[1] += 1
        slot.wait_reusable()
        slot.data[...] = data
        slot.headers[...] = headers
        slot.glob[...] = glob
        return slot


class KaldiDataRandomQueue(object):
    """data_loader.py:310-414."""

    def __init__(self, data_dir, spklist, num_parallel=1, max_qsize=10, num_speakers=None, num_segments=None, min_len=None,
                 max_len=None, shuffle=True, base_seed=None):
        self.data = data_dir
        self.num_speakers = num_speakers
        self.num_segments = num_segments
        self.min_len = min_len
        self.max_len = max_len
        self.num_parallel_datasets = num_parallel
        self.shuffle = shuffle
        self.base_seed = base_seed
        self.spk2features, self.features2spk, spk2index = get_speaker_info(data_dir, spklist)
        self.num_total_speakers = len(list(spk2index.keys()))
        self.queue = Queue(max_qsize)
        self.stop_event = Event()
        self.processes = []
        self._ring = _PinnedRing()

    def set_batch(self, num_speakers, num_segments):
        self.num_speakers = num_speakers
        self.num_segments = num_segments

    def set_length(self, min_len, max_len):
        self.min_len = min_len
        self.max_len = max_len

    def start(self):
        self.processes = [Process(target=batch_random, args=(self.stop_event, self.queue, self.data, self.spk2features,
                                                             self.num_total_speakers, self.num_speakers, self.num_segments,
                                                             self.min_len, self.max_len, self.shuffle, i, self.base_seed))
                          for i in range(self.num_parallel_datasets)]
        for process in self.processes:
            process.daemon = True
            process.start()

    def fetch(self):
        """-> (features, labels): features is a CompressedSegmentBatch (pinned uint8 crops; Trainer.train_step decodes it
        on the device), labels int32 [B]."""
        data, headers, glob, labels = self.queue.get()
        return self._ring.wrap(data, headers, glob), labels

    def stop(self):
        self.stop_event.set()
        while not self.queue.empty():
            try:
                self.queue.get(block=False)
            except Exception:
                break
        time.sleep(0.3)
        for process in self.processes:
            process.terminate()


def batch_sequence(stop_event, queue, data, feature_list, features2spk, batch_size=128, min_len=200, max_len=400,
                   shuffle=True, seed=0, base_seed=None):
    """Worker of KaldiDataSeqQueue (data_loader.py:417-462): every segment once, one length per batch, clamped to the
    shortest utterance of the batch."""
    rd = _make_rng(seed, base_seed)
    feature_reader = FeatureReader(data)
    num_batches = int(len(feature_list) / batch_size)
    for i in range(num_batches):
        batch_length = rd.randint(min_len, max_len)
        for j in range(batch_size):
            n = feature_reader.utt2num_frames[feature_list[i * batch_size + j].split(" ")[0]]
            if n < batch_length:
                batch_length = n
        raws, labels = [], []
        for j in range(batch_size):
            feat = feature_list[i * batch_size + j]
            raw, _ = feature_reader.read_segment(feat, batch_length, shuffle=shuffle, rd=rd)
            raws.append(raw)
            labels.append(features2spk[feat])
        queue.put(_pack(raws, labels))
    queue.put(None)             # end marker: the queue is FIFO per producer, so everything before it has arrived
    stop_event.set()
    return


class KaldiDataSeqQueue(object):
    """data_loader.py:465-560."""

    def __init__(self, data_dir, spklist, num_parallel=1, max_qsize=10, batch_size=128, min_len=None, max_len=None,
                 shuffle=True, base_seed=None):
        self.data = data_dir
        self.batch_size = batch_size
        self.min_len = min_len
        self.max_len = max_len
        self.num_parallel_datasets = num_parallel
        self.shuffle = shuffle
        self.base_seed = base_seed
        self.spk2features, self.features2spk, spk2index = get_speaker_info(data_dir, spklist)
        self.num_total_speakers = len(list(spk2index.keys()))
        self.feature_list = []
        self.sub_feature_list = []
        for spk in self.spk2features:
            self.feature_list += self.spk2features[spk]
        if shuffle:
            (random if base_seed is None else random.Random(base_seed)).shuffle(self.feature_list)
        num_sub_features = len(self.feature_list) // num_parallel          # py2 integer division (data_loader.py:504)
        for i in range(num_parallel):
            if i == num_parallel - 1:
                self.sub_feature_list.append(self.feature_list[i * num_sub_features:])
            else:
                self.sub_feature_list.append(self.feature_list[i * num_sub_features:(i + 1) * num_sub_features])
        self.queue = Queue(max_qsize)
        self.stop_event = [Event() for _ in range(num_parallel)]
        self.processes = []
        self._ring = _PinnedRing()
        self._finished = 0

    def set_batch(self, batch_size):
        self.batch_size = batch_size

    def set_length(self, min_len, max_len):
        self.min_len = min_len
        self.max_len = max_len

    def start(self):
        self.processes = [Process(target=batch_sequence, args=(self.stop_event[i], self.queue, self.data,
                                                               self.sub_feature_list[i], self.features2spk, self.batch_size,
                                                               self.min_len, self.max_len, self.shuffle, i, self.base_seed))
                          for i in range(self.num_parallel_datasets)]
        for process in self.processes:
            process.daemon = True
            process.start()

    def fetch(self):
        """data_loader.py:541-552: DataOutOfRange once every worker has finished and the queue is drained."""
        while self._finished < self.num_parallel_datasets:
            item = self.queue.get()
            if item is None:
                self._finished += 1
                continue
            data, headers, glob, labels = item
            return self._ring.wrap(data, headers, glob), labels
        raise DataOutOfRange

    def stop(self):
        for process in self.processes:
            process.terminate()
