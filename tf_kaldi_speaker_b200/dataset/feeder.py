"""Host feeder for compressed Kaldi features: raw uint8 segment crops in pinned memory -> float32 [B, T, D] on the GPU.

The reference's loader processes dequantise every segment with NumPy on the host (dataset/kaldi_io.py:784-797, 852-868,
called from data_loader.py:229-307) and feed float32 [B, T, D] through ``feed_dict`` (model/trainer.py:505-508).  Here the
host only gathers the bytes: per segment the uint8 crop [D, T] (column-major, as stored), the four uint16 percentiles of
every column and the matrix' (min, range); ``xv_cm_decode`` applies the reference's three-piece linear map and the
transpose on the device, bit-exactly (tests/test_cm_decode_gpu.py against the reference reader's golden vectors).
PCIe bytes per 128 x 200 x 30 batch: 0.80 MB instead of 3.07 MB."""
import ctypes as C

import numpy as np
import torch

from .. import _lib as L


class CompressedSegmentBatch(object):
    """B equal-length segments of compressed features, staged in (pinned) host memory."""

    def __init__(self, batch, frames, dim, pin=True):
        self.B, self.T, self.D = int(batch), int(frames), int(dim)
        pin = bool(pin) and torch.cuda.is_available()
        self._data = torch.empty((self.B, self.D, self.T), dtype=torch.uint8, pin_memory=pin)
        self._headers = torch.empty((self.B, self.D, 4), dtype=torch.int16, pin_memory=pin)     # uint16 bit patterns
        self._glob = torch.empty((self.B, 2), dtype=torch.float32, pin_memory=pin)
        self.data = self._data.numpy()
        self.headers = self._headers.numpy().view(np.uint16)
        self.glob = self._glob.numpy()
        self._dev = None
        self._copied = None          # event recorded after the H2D copies of the last decode_into()

    @property
    def shape(self):
        return (self.B, self.T, self.D)

    @property
    def h2d_bytes(self):
        return self._data.numel() + self._headers.numel() * 2 + self._glob.numel() * 4

    def set(self, i, raw):
        """Place CompressedRaw ``raw`` (dataset.kaldi_io.read_compressed_raw) as segment i."""
        if raw.cols != self.D or raw.data.shape[1] != self.T:
            raise ValueError("segment %d: expected %d frames x %d dims, got %d x %d"
                             % (i, self.T, self.D, raw.data.shape[1], raw.cols))
        self.data[i] = raw.data
        self.headers[i] = raw.headers
        self.glob[i, 0] = raw.globmin
        self.glob[i, 1] = raw.globrange

    def decode_into(self, out):
        """H2D copy of the raw bytes + xv_cm_decode on the current stream -> ``out`` float32 [B, T, D] (device)."""
        assert out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == self.shape and out.is_contiguous()
        if self._dev is None or self._dev[0].device != out.device:
            self._dev = (torch.empty_like(self._data, device=out.device), torch.empty_like(self._headers, device=out.device),
                         torch.empty_like(self._glob, device=out.device))
        d, h, g = self._dev
        d.copy_(self._data, non_blocking=True)
        h.copy_(self._headers, non_blocking=True)
        g.copy_(self._glob, non_blocking=True)
        if self._copied is None:
            self._copied = torch.cuda.Event()
        self._copied.record()
        L.check(L.load().xv_cm_decode(L.ptr(d), L.ptr(h), L.ptr(g), L.ptr(out), self.B, self.T, self.D,
                                      C.c_int64(self.T), C.c_int64(self.D), L.stream_ptr()))
        return out

    def wait_reusable(self):
        """Block until the pinned arrays may be overwritten (the last upload has read them)."""
        if self._copied is not None:
            self._copied.synchronize()

    def to_device(self, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        return self.decode_into(torch.empty(self.shape, dtype=torch.float32, device=dev))
