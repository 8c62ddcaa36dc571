"""heimdall_core.camera -- frame feed: camera frames -> detector input.

Mirrors the value types and the two conversion helpers of the reference's `heimdall-camera` crate
(rust/heimdall-camera/src/lib.rs): `PixelFormat` (:34-47), `CameraFrame` (:111-132), `to_ndarray` (:260-278) and the
cv2.cvtColor conversions named by `to_opencv_mat` (:203-257).  The conversions of the mosaic and YUV formats run on
the GPU (csrc/k_pixfmt.cu); `Detector.submit_frames` feeds whole batches of raw frames to the detector without ever
materialising the RGB image.
"""
from __future__ import annotations

import enum
import time
from dataclasses import dataclass, field
from typing import Dict

import numpy as np

from . import _abi as A


class PixelFormat(enum.IntEnum):
    """rust/heimdall-camera/src/lib.rs:34-47 (same order)."""
    Mono8 = A.HV_PIX_MONO8
    Mono16 = A.HV_PIX_MONO16
    RGB8 = A.HV_PIX_RGB8
    BGR8 = A.HV_PIX_BGR8
    RGBA8 = A.HV_PIX_RGBA8
    BGRA8 = A.HV_PIX_BGRA8
    YUV422 = A.HV_PIX_YUV422
    YUV422Packed = A.HV_PIX_YUV422_PACKED
    BayerRG8 = A.HV_PIX_BAYER_RG8
    BayerGB8 = A.HV_PIX_BAYER_GB8
    BayerGR8 = A.HV_PIX_BAYER_GR8
    BayerBG8 = A.HV_PIX_BAYER_BG8


RAW_BYTES_PER_PIXEL = {PixelFormat.Mono8: 1, PixelFormat.Mono16: 2, PixelFormat.RGB8: 3, PixelFormat.BGR8: 3,
                       PixelFormat.RGBA8: 4, PixelFormat.BGRA8: 4, PixelFormat.YUV422: 2, PixelFormat.YUV422Packed: 2,
                       PixelFormat.BayerRG8: 1, PixelFormat.BayerGB8: 1, PixelFormat.BayerGR8: 1,
                       PixelFormat.BayerBG8: 1}


class ConversionError(ValueError):
    """CameraError::ConversionError (lib.rs:25-26)."""


@dataclass
class CameraFrame:
    """lib.rs:111-132.  `data` is the raw frame as a flat uint8 array."""
    data: np.ndarray
    width: int
    height: int
    pixel_format: PixelFormat
    timestamp: float = field(default_factory=time.time)
    frame_id: int = 0
    metadata: Dict[str, str] = field(default_factory=dict)
    camera: int = 0

    def as_abi(self) -> A.hv_camera_frame:
        d = np.ascontiguousarray(self.data, dtype=np.uint8).reshape(-1)
        self.data = d  # keep the contiguous buffer alive as long as the frame
        return A.hv_camera_frame(d.ctypes.data, d.size, int(self.width), int(self.height), int(self.pixel_format),
                                 int(self.camera), int(self.frame_id), int(self.timestamp * 1e9))


def frame_channels(fmt: PixelFormat) -> int:
    return int(A.lib.hv_frame_channels(int(fmt)))


def to_ndarray(frame: CameraFrame) -> np.ndarray:
    """lib.rs:260-278: Mono8 / RGB8 / BGR8 / RGBA8 / BGRA8 only, bytes unchanged, shape (h, w, channels)."""
    ch = {PixelFormat.Mono8: 1, PixelFormat.RGB8: 3, PixelFormat.BGR8: 3, PixelFormat.RGBA8: 4,
          PixelFormat.BGRA8: 4}.get(PixelFormat(frame.pixel_format))
    if ch is None:
        raise ConversionError("Erreur de conversion d'image: Format de pixel non supporté pour la conversion ndarray: "
                              f"{PixelFormat(frame.pixel_format).name}")
    d = np.asarray(frame.data, dtype=np.uint8).reshape(-1)
    if d.size != frame.height * frame.width * ch:
        raise ConversionError("Erreur de conversion d'image: Erreur de conversion en ndarray: shape mismatch")
    return d.reshape(frame.height, frame.width, ch).copy()


def to_image(frame: CameraFrame, detector=None) -> np.ndarray:
    """The (h, w, c) image the detector sees for this frame: `to_ndarray` for the pass-through formats, the
    cv2.cvtColor conversion named by `to_opencv_mat` (lib.rs:226-252) for Bayer*8 / YUV422*, computed on the GPU."""
    from .batch import default_detector
    det = detector or default_detector()
    return det.convert_frame(frame)


class SyncMode(enum.IntEnum):
    """rust/heimdall-gige/src/sync.rs:18-27."""
    Freerun = A.HV_SYNC_FREERUN
    Software = A.HV_SYNC_SOFTWARE
    Hardware = A.HV_SYNC_HARDWARE


class FrameSetBatcher:
    """N2: groups frames arriving camera by camera into `FrameSet`s (rust/heimdall-gige/src/frame.rs:127-185) and feeds
    `sets_per_batch` complete sets at a time to the detector (set-major, camera-minor).  `detector=None` gives the
    dry mode: the batching rules run, nothing is submitted (CPU tests)."""

    def __init__(self, detector, n_cameras: int, sets_per_batch: int = 1, sync_mode: SyncMode = SyncMode.Hardware,
                 max_pending_sets: int = 0, params=None, zero_copy: bool = False):
        import ctypes as C
        self._C = C
        self._det = detector
        self._params = params
        self.n_cameras, self.sets_per_batch = int(n_cameras), int(sets_per_batch)
        # zero_copy: frames that already lie in page-locked memory (Detector.host_alloc) are referenced, not copied; the
        # caller keeps them unchanged until wait() has returned for their batch
        cfg = A.hv_frameset_config(self.n_cameras, self.sets_per_batch, int(sync_mode), int(max_pending_sets),
                                   A.HV_FRAMESET_ZERO_COPY if zero_copy else 0, 0)
        h = C.c_void_p()
        st = A.lib.hv_frameset_create(detector._ctx if detector is not None else None, C.byref(cfg), C.byref(h))
        if st != A.HV_OK:
            raise ValueError(f"hv_frameset_create: {A.lib.hv_status_string(st).decode()}")
        self._h = h

    def close(self) -> None:
        if self._h:
            A.lib.hv_frameset_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def push(self, frame: "CameraFrame") -> int:
        """Returns the ticket of the batch this frame completed, else 0 (dry mode: -1, -2, ...)."""
        C = self._C
        t = C.c_int64(0)
        fa = frame.as_abi()
        p = self._params
        st = A.lib.hv_frameset_push(self._h, C.byref(fa), C.byref(p) if p is not None else None, C.byref(t))
        if st != A.HV_OK:
            self._raise(st)
        return t.value

    def flush(self) -> int:
        """Retry the submission of a batch that push() could not hand over (HeimdallCudaError, status HV_ERR_CAPACITY:
        collect an earlier ticket with wait() first).  Returns the ticket, or 0 when nothing is queued."""
        C = self._C
        t = C.c_int64(0)
        p = self._params
        st = A.lib.hv_frameset_flush(self._h, C.byref(p) if p is not None else None, C.byref(t))
        if st != A.HV_OK:
            self._raise(st)
        return t.value

    def _raise(self, st: int):
        from .batch import HeimdallCudaError
        msg = A.lib.hv_frameset_last_error(self._h).decode() or A.lib.hv_status_string(st).decode()
        if st == A.HV_ERR_CAPACITY:
            raise HeimdallCudaError(st, msg)
        raise ValueError(msg)

    def batch_ids(self, ticket: int):
        C = self._C
        ids = (C.c_uint64 * self.sets_per_batch)()
        n = C.c_int32(0)
        st = A.lib.hv_frameset_batch_ids(self._h, ticket, ids, self.sets_per_batch, C.byref(n))
        if st != A.HV_OK:
            raise ValueError(A.lib.hv_frameset_last_error(self._h).decode())
        return [int(ids[k]) for k in range(n.value)]

    def wait(self, ticket: int):
        """BatchResult of a submitted batch: frame f = set f // n_cameras (see batch_ids), camera f % n_cameras."""
        from .batch import BatchResult
        C = self._C
        n = self.n_cameras * self.sets_per_batch
        res, dfx, cap = self._det._out_arrays(n, None)
        total = C.c_size_t(0)
        st = A.lib.hv_frameset_wait(self._h, ticket, res.ctypes.data_as(C.POINTER(A.hv_frame_result)),
                                    dfx.ctypes.data_as(C.POINTER(A.hv_defect)), cap, C.byref(total))
        if st not in (A.HV_OK, A.HV_ERR_CAPACITY):
            raise ValueError(A.lib.hv_frameset_last_error(self._h).decode() or A.lib.hv_status_string(st).decode())
        return BatchResult(res, dfx[:total.value], st)

    def stats(self) -> Dict[str, int]:
        s = A.hv_frameset_stats()
        A.lib.hv_frameset_get_stats(self._h, self._C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in A.hv_frameset_stats._fields_}
