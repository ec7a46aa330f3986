"""Wire layouts of the hot path and Unity.Mathematics half conversions (numpy).

The structured dtypes below are byte-for-byte the C# sequential layouts the
reference jobs read (every member is 2 bytes, so there is no padding):

* ``AABB_DT``   20 B  <- Assets/C# Scripts/DataTypes/Collider Structs/ColliderAABBStruct.cs:8-14
* ``OBB_DT``    26 B  <- .../ColliderOBBStruct.cs:8-24 (+ DataTypes/halfQuaternion.cs:7-11)
* ``SPHERE_DT`` 16 B  <- .../ColliderSphereStruct.cs:8-14
* ``SETTINGS_DT`` 24 B <- DataTypes/AudioTargetRTSettings.cs:8-24

They are what crosses the C ABI (include/audiort.h ArtAABB / ArtOBB / ArtSphere /
ArtTargetSettings).
"""
from __future__ import annotations

import numpy as np

_MAT = [("absorption", "<u2"), ("density", "<u2"), ("echo", "<u2")]

AABB_DT = np.dtype([("center", "<u2", (3,)), ("size", "<u2", (3,))] + _MAT + [("audioTargetId", "<i2")])
OBB_DT = np.dtype([("center", "<u2", (3,)), ("size", "<u2", (3,)), ("rot", "<u2", (3,))] + _MAT + [("audioTargetId", "<i2")])
SPHERE_DT = np.dtype([("center", "<u2", (3,)), ("radius", "<u2")] + _MAT + [("audioTargetId", "<i2")])
SETTINGS_DT = np.dtype([("muffleStrength", "<f4"), ("reverbStrength", "<f4"), ("reverbVolume", "<f4"),
                        ("percievedAudioPosition", "<f4", (3,))])

assert AABB_DT.itemsize == 20 and OBB_DT.itemsize == 26 and SPHERE_DT.itemsize == 16 and SETTINGS_DT.itemsize == 24

# Enums/ColliderType.cs: None, AABB, OBB, Sphere  (hit-id extension output = type << 30 | index)
TYPE_NONE, TYPE_AABB, TYPE_OBB, TYPE_SPHERE = 0, 1, 2, 3


def f32tof16(x) -> np.ndarray:
    """Unity.Mathematics 1.3.2 ``math.f32tof16`` on an array of float32 -> uint16 bits.

    Truncate the low 12 mantissa bits, rescale by 2**-112, clamp, add 0x1000,
    shift right 13: round-to-nearest with ties AWAY from zero (IEEE RNE differs
    on exact ties). NaN -> 0x7e00, Inf -> 0x7c00.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    ux = x.view(np.uint32)
    msk = np.uint32(0x7FFFF000)
    uux = ux & msk
    with np.errstate(under="ignore", over="ignore", invalid="ignore"):
        scaled = uux.view(np.float32) * np.float32(1.92592994e-34)
    sb = np.minimum(scaled.view(np.uint32), np.uint32(0x0F7FF000))
    h = (sb + np.uint32(0x1000)) >> np.uint32(13)
    inf32 = np.uint32(255 << 23)
    h = np.where(uux >= inf32, np.where(uux > inf32, np.uint32(0x7E00), np.uint32(0x7C00)), h)
    return (h | ((ux & ~msk) >> np.uint32(16))).astype(np.uint16)


def f16tof32(h) -> np.ndarray:
    """``math.f16tof32``: exact, identical to IEEE binary16 -> binary32."""
    return np.ascontiguousarray(h, dtype=np.uint16).view(np.float16).astype(np.float32)


def half_round(x) -> np.ndarray:
    """float32 -> Unity half -> float32 (the value the jobs will actually see)."""
    return f16tof32(f32tof16(x))
