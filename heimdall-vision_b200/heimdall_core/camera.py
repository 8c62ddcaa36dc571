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
