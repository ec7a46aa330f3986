"""Host-side mirror of the reference's job-scheduling seam (Python twin of the C# side).

The reference fills three job structs field by field and schedules them every frame
(Assets/C# Scripts/Audio/AudioRayTracer.cs:161-237); results are consumed one frame later after
``mainJobHandle.Complete()`` (ART:95-107). The classes below keep the reference's field names so that
the parity tests read like the reference's own scheduling code; ``AudioRayTracer.OnUpdate`` does what
ART:92-238 does, with the three ``Schedule()`` calls replaced by one ``art_trace_schedule``.
Nothing here computes: the work is done by libaudiort_cuda (see native.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import native
from .layouts import AABB_DT, OBB_DT, SPHERE_DT
from .scenes import Scene


@dataclass
class AudioRaytracerJobBatched:
    """Field-for-field mirror of RT:12-52."""
    RayOrigin: np.ndarray
    RayDirections: np.ndarray                     # half3[N] as uint16 [N,3]
    AABBColliders: np.ndarray
    AABBColliderCount: int
    OBBColliders: np.ndarray
    OBBColliderCount: int
    SphereColliders: np.ndarray
    SphereColliderCount: int
    AudioTargetPositions: np.ndarray              # float3[Na]
    TotalAudioTargets: int
    MaxRayLife: float
    MaxHitsPerRay: int
    MaxMuffleHitDistance: float
    # outputs (RT:34-50)
    RayHitResults: Optional[np.ndarray] = None
    RayHitResultCounts: Optional[np.ndarray] = None
    EchoRayDistances: Optional[np.ndarray] = None
    MuffleRayHits: Optional[np.ndarray] = None


@dataclass
class AudioPermeationJobBatched:
    """Mirror of PM:10-27 (shares inputs with the ray tracer job)."""
    PermeationStrengthPerRay: float
    PermeationPowerRemains: Optional[np.ndarray] = None


@dataclass
class ProcessAudioDataJob:
    """Mirror of PA:10-28."""
    MuffleEffectiveness: float
    PermeationEffectiveness: float
    MaxReverbDistance: float
    AudioTargetSettings: Optional[np.ndarray] = None


class JobHandle:
    """≙ Unity's JobHandle for the combined frame (ART:46, 237)."""

    def __init__(self, ctx: native.Context, handle: int):
        self._ctx, self._h = ctx, handle
        self.result: Optional[native.FrameResult] = None

    @property
    def IsCompleted(self) -> bool:          # ART:95
        return self._ctx.is_completed(self._h)

    def Complete(self) -> native.FrameResult:   # ART:97
        if self.result is None:
            self.result = self._ctx.complete(self._h)
        return self.result


@dataclass
class AudioRayTracer:
    """Mirror of the MonoBehaviour's inspector fields (ART:9-35) + the per-frame scheduling block."""
    rayOrigin: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.65, 0.0], np.float32))
    rayCount: int = 1000
    maxBounces: int = 3
    maxRayLife: float = 10.0
    maxMuffleHitDistance: float = 10.0
    muffleEffectiveness: float = 1.0
    mufflePermeationEffectiveness: float = 0.5
    permeationStrengthPerRay: float = 1.0
    maxReverbDistance: float = 20.0
    toUseThreadCount: int = 3                  # AudioRaytracingManager.ToUseThreadCount (ARM:19)
    device: int = 0

    def __post_init__(self):
        self._ctx = native.Context(self.device)
        self._handle: Optional[JobHandle] = None
        # InitializeAudioRaytraceSystem (ART:66-87): Fibonacci directions generated on the device (FIB:15-35)
        self._ctx.generate_fibonacci_rays(self.rayCount)

    @property
    def MaxHitsPerRay(self) -> int:            # ART:16
        return self.maxBounces + 1

    def set_colliders(self, aabbs, obbs, spheres):
        """≙ AudioColliderManager.UpdateJobBatch (ART:155): hand the baked struct arrays to the plugin."""
        self._ctx.set_scene(np.asarray(aabbs, AABB_DT), np.asarray(obbs, OBB_DT), np.asarray(spheres, SPHERE_DT))

    def OnUpdate(self, transform_position, audio_target_positions) -> Optional[native.FrameResult]:
        """ART:92-238. Returns last frame's results (consumed one frame late, ART:107) or None while running."""
        last = None
        if self._handle is not None:
            if not self._handle.IsCompleted:                       # ART:95 async poll
                return None
            last = self._handle.Complete()                         # ART:97
        targets = np.asarray(audio_target_positions, np.float32).reshape(-1, 3)
        if len(targets) == 0:                                      # ART:95 AudioTargetCount_NextBatch == 0
            return last
        scene = Scene(aabbs=np.zeros(0, AABB_DT), obbs=np.zeros(0, OBB_DT), spheres=np.zeros(0, SPHERE_DT),
                      targets=targets, ray_directions=np.zeros((self.rayCount, 3), np.uint16),
                      ray_origin=np.asarray(transform_position, np.float32) + self.rayOrigin,   # ART:165
                      max_ray_life=self.maxRayLife, max_hits_per_ray=self.MaxHitsPerRay,
                      max_muffle_hit_distance=self.maxMuffleHitDistance,
                      permeation_strength_per_ray=self.permeationStrengthPerRay,
                      muffle_effectiveness=self.muffleEffectiveness,
                      permeation_effectiveness=self.mufflePermeationEffectiveness,
                      max_reverb_distance=self.maxReverbDistance, batch_count=self.toUseThreadCount)
        self._handle = JobHandle(self._ctx, self._ctx.schedule(scene))   # ART:191 + 213 + 237
        return last

    def OnDestroy(self):                        # ART:241-254
        if self._handle is not None:
            self._handle.Complete()
        self._ctx.close()
