"""Mirror of HoniHelper.hs for replayed / synthetic OpenNI depth streams.

takeDepthSnapshot :: IO (Either String (Vector Word16, (Int, Int)))     (HoniHelper.hs:20-36)
withHoni          :: IO a -> IO a                                       (HoniHelper.hs:50-56)

The OpenNI2 device I/O itself is out of scope (no device in a GPU box); the frame format is kept: raw
`frameData` bytes reinterpreted as host-endian uint16, row-major, plus (width, height) (HoniHelper.hs:34-36,45-46).
A frame source is any iterator of (bytes | ndarray, (w, h)); `Left err` is returned as ("Left", err)."""
from __future__ import annotations

import numpy as np

_source = None


def bsToVector16Bits(frame_data) -> np.ndarray:
    """ByteString -> Vector Word16, zero-copy reinterpretation (HoniHelper.hs:45-46)."""
    if isinstance(frame_data, np.ndarray):
        return frame_data.view(np.uint16).reshape(-1)
    return np.frombuffer(frame_data, dtype=np.uint16)


def setFrameSource(it):
    """Install the replayed stream that stands in for the first OpenNI2 device."""
    global _source
    _source = iter(it) if it is not None else None


def takeDepthSnapshot():
    if _source is None:
        return ("Left", "No depth device")  # HoniHelper.hs:28
    try:
        data, (w, h) = next(_source)
    except StopIteration:
        return ("Left", "streamReadFrame: end of replayed stream")
    vec = bsToVector16Bits(data)
    if vec.size != w * h:
        return ("Left", "frame size does not match (width, height)")
    return ("Right", (vec, (w, h)))


def withHoni(action):
    return action()


def addDevicePointCloud(ctx=None):
    """Main.addDevicePointCloud (Main.hs:1282-1313) on the GPU: returns (points, mask) or raises on `Left`."""
    from . import default_context

    s = takeDepthSnapshot()
    if s[0] == "Left":
        raise RuntimeError("WARNING: " + s[1])  # Main.hs:1289
    vec, (w, h) = s[1]
    return (ctx or default_context()).backproject_ref(vec, w, h)
