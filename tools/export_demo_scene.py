#!/usr/bin/env python
"""Export the reference's demo level (BASELINE config 1) to the job-input wire layout.

    python tools/export_demo_scene.py [--reference /root/reference] [--out tests/golden/c1_demo_scene.npz]

Runs only where the reference tree is mounted (this container); the GPU box gets the committed fixture. It is an
OFFLINE restatement of what Unity does between scene load and the first `AudioRayTracer.OnUpdate`:

* parse `Assets/Scenes/Sample Scene.unity` (Unity YAML: GameObject / Transform / MonoBehaviour documents);
* compose each Transform's world position / rotation / lossy scale through `m_Father` (TRS matrices; Unity does this
  in its native transform system, so this is the published semantics, not code from the tree);
* bake every enabled Audio{AABB,OBB,Sphere}Collider exactly as `GetBakedColliderStruct` does
  (Audio/Colliders/AudioAABBCollider.cs:27-50, AudioOBBCollider.cs:31-66, AudioSphereCollider.cs:27-57):
  FP32 add / multiply, then Unity `(half)` rounding; the OBB's private `rotation` field is not serialized, so it starts
  as identity, is multiplied by `transform.rotation` (UnityEngine `Quaternion * Quaternion`), squeezed through
  halfQuaternion (3 halves, w >= 0), optionally multiplied by `Quaternion.Euler(rotationEulerOffset)`, inverted
  (`math.inverse`) and squeezed again;
* material = the referenced ScriptableObject's three raw halves (ScriptableObjects/AudioMaterials/*.asset) or
  `AudioMaterialProperties.Default` (0, 1, 1);
* AudioTargetRT components get ids 0.. in registration order, colliders on the same GameObject inherit the id
  (AudioCollider.cs:29-36, AudioTargetManager.cs:49-57), everything else -1;
* job parameters from `Prefabs/Player.prefab` + the scene's PrefabInstance overrides (ART:9-35):
  RayOrigin = (float3)transform.position + rayOrigin (ART:165).

ASSUMPTION (engine behaviour, not in the tree): OnEnable order == document order of the scene file. It only affects
the order of colliders inside each typed array (tie-breaking and hit ids) and which MusicBox gets which target id.
"""
from __future__ import annotations

import argparse
import os
import re
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_raytracer_b200.layouts import AABB_DT, OBB_DT, SPHERE_DT, f16tof32, f32tof16  # noqa: E402

GUID_AABB = "60dde53da7d8d1c4a9d5adc1fd2f51b0"
GUID_OBB = "3556f4f37fe1b374bb2efc2590368bb9"
GUID_SPHERE = "6933d97b3db32fb4e9006d8727c6fad8"
GUID_TARGET = "7aca81e69a4ea0d40a7c10992aeac491"
GUID_RAYTRACER = "d14e0eebbfc9f5b4685365d70167ab2b"
GUID_RT_MANAGER = "91af001a7a3c2e44bac1966bbbe2f07c"

f32 = np.float32


# ---------------------------------------------------------------------------------------------
# Unity YAML
# ---------------------------------------------------------------------------------------------
def parse_unity_yaml(path):
    """-> list of (classId, fileId, stripped, body dict) in document order."""
    text = open(path, encoding="utf-8-sig").read()
    docs = []
    for m in re.finditer(r"^--- !u!(\d+) &(\d+)( stripped)?\n(.*?)(?=^--- |\Z)", text, re.S | re.M):
        body = yaml.safe_load(m.group(4))
        docs.append((int(m.group(1)), int(m.group(2)), bool(m.group(3)), body))
    return docs


def vec(d, keys="xyz"):
    return np.array([float(d[k]) for k in keys], dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# quaternion helpers (x, y, z, w), FP32 like the managed code
# ---------------------------------------------------------------------------------------------
def qmul(a, b):
    """UnityEngine.Quaternion operator* (Hamilton product), FP32 products and sums in the documented order."""
    ax, ay, az, aw = [f32(v) for v in a]
    bx, by, bz, bw = [f32(v) for v in b]
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], dtype=np.float32)


def qrot(q, v):
    """UnityEngine.Quaternion * Vector3."""
    x, y, z, w = [f32(c) for c in q]
    px, py, pz = [f32(c) for c in v]
    x2, y2, z2 = x * f32(2), y * f32(2), z * f32(2)
    xx, yy, zz = x * x2, y * y2, z * z2
    xy, xz, yz = x * y2, x * z2, y * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    return np.array([(f32(1) - (yy + zz)) * px + (xy - wz) * py + (xz + wy) * pz,
                     (xy + wz) * px + (f32(1) - (xx + zz)) * py + (yz - wx) * pz,
                     (xz - wy) * px + (yz + wx) * py + (f32(1) - (xx + yy)) * pz], dtype=np.float32)


def qeuler(deg):
    """UnityEngine.Quaternion.Euler: rotate z, then x, then y  ->  q = qy * qx * qz."""
    hx, hy, hz = [np.deg2rad(float(d)) * 0.5 for d in deg]
    qx = np.array([np.sin(hx), 0, 0, np.cos(hx)], dtype=np.float32)
    qy = np.array([0, np.sin(hy), 0, np.cos(hy)], dtype=np.float32)
    qz = np.array([0, 0, np.sin(hz), np.cos(hz)], dtype=np.float32)
    return qmul(qmul(qy, qx), qz)


def qmat(q):
    x, y, z, w = [float(c) for c in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def half_quat_store(q):
    """halfQuaternion setter (DataTypes/halfQuaternion.cs:47-61): keep x,y,z as halves with w >= 0."""
    q = np.asarray(q, dtype=np.float32)
    xyz = -q[:3] if q[3] < 0 else q[:3]
    return f32tof16(xyz.astype(np.float32))


def half_quat_load(bits):
    """halfQuaternion getter (DataTypes/halfQuaternion.cs:34-46): w = sqrt(max(0, 1 - |xyz|^2)), math.normalize."""
    x, y, z = [f32(v) for v in f16tof32(bits)]
    w2 = f32(1) - (x * x + y * y + z * z)
    w = np.sqrt(w2) if w2 > 0 else f32(0)
    q = np.array([x, y, z, w], dtype=np.float32)
    dot = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]
    return (f32(1) / np.sqrt(dot)) * q


def qinverse(q):
    """Unity.Mathematics math.inverse(quaternion): rcp(dot(q,q)) * q * (-1,-1,-1,1)."""
    q = np.asarray(q, dtype=np.float32)
    dot = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]
    return (f32(1) / dot) * q * np.array([-1, -1, -1, 1], dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# scene graph
# ---------------------------------------------------------------------------------------------
class SceneGraph:
    def __init__(self, docs):
        self.gameobjects = {fid: b["GameObject"] for c, fid, s, b in docs if c == 1 and not s}
        self.transforms = {fid: b[next(iter(b))] for c, fid, s, b in docs if c in (4, 224) and not s}
        self.go_transform = {t["m_GameObject"]["fileID"]: fid for fid, t in self.transforms.items()}
        self._world = {}

    def world(self, tid):
        """-> (4x4 world matrix fp64, world rotation quaternion fp32)."""
        if tid in self._world:
            return self._world[tid]
        t = self.transforms[tid]
        lp, ls = vec(t["m_LocalPosition"]), vec(t["m_LocalScale"])
        lq = np.array([float(t["m_LocalRotation"][k]) for k in "xyzw"], dtype=np.float32)
        M = np.eye(4)
        M[:3, :3] = qmat(lq) @ np.diag(ls)
        M[:3, 3] = lp
        father = t["m_Father"]["fileID"]
        if father:
            PM, pq = self.world(father)
            M = PM @ M
            q = qmul(pq, lq)
        else:
            q = lq
        self._world[tid] = (M, q)
        return self._world[tid]

    def position(self, tid):
        return self.world(tid)[0][:3, 3].astype(np.float32)

    def rotation(self, tid):
        return self.world(tid)[1]

    def lossy_scale(self, tid):
        M, q = self.world(tid)
        return np.diag(qmat(q).T @ M[:3, :3]).astype(np.float32)

    def active_in_hierarchy(self, go_id):
        tid = self.go_transform[go_id]
        while tid:
            t = self.transforms[tid]
            if not self.gameobjects[t["m_GameObject"]["fileID"]].get("m_IsActive", 1):
                return False
            tid = t["m_Father"]["fileID"]
        return True


def half3_bits(d):
    return np.array([d["x"]["value"], d["y"]["value"], d["z"]["value"]], dtype=np.uint16)


def load_materials(ref):
    out = {}
    mdir = os.path.join(ref, "Assets", "ScriptableObjects", "AudioMaterials")
    for fn in sorted(os.listdir(mdir)):
        if not fn.endswith(".asset"):
            continue
        guid = re.search(r"guid: (\w+)", open(os.path.join(mdir, fn + ".meta")).read()).group(1)
        body = parse_unity_yaml(os.path.join(mdir, fn))[0][3]["MonoBehaviour"]
        mp = body["MaterialProperties"]
        out[guid] = (fn[:-6], np.array([mp["Absorption"]["value"], mp["Density"]["value"], mp["Echo"]["value"]], dtype=np.uint16))
    return out


DEFAULT_MATERIAL = np.array([0x0000, 0x3C00, 0x3C00], dtype=np.uint16)   # AudioMaterialProperties.Default


def export(ref):
    scene_path = os.path.join(ref, "Assets", "Scenes", "Sample Scene.unity")
    docs = parse_unity_yaml(scene_path)
    g = SceneGraph(docs)
    materials = load_materials(ref)

    behaviours = [(fid, b["MonoBehaviour"]) for c, fid, s, b in docs if c == 114 and not s and "MonoBehaviour" in b]

    def live(mb):
        return mb.get("m_Enabled", 1) and g.active_in_hierarchy(mb["m_GameObject"]["fileID"])

    # ---- audio targets (registration order = document order) --------------------------------
    target_of_go, targets = {}, []
    for fid, mb in behaviours:
        if mb["m_Script"].get("guid") == GUID_TARGET and live(mb):
            go = mb["m_GameObject"]["fileID"]
            target_of_go[go] = len(targets)
            targets.append(g.position(g.go_transform[go]))

    aabbs, obbs, spheres, names = [], [], [], {"aabb": [], "obb": [], "sphere": []}
    for fid, mb in behaviours:
        guid = mb["m_Script"].get("guid")
        if guid not in (GUID_AABB, GUID_OBB, GUID_SPHERE) or not live(mb):
            continue
        go = mb["m_GameObject"]["fileID"]
        tid = g.go_transform[go]
        pos, rot, scale = g.position(tid), g.rotation(tid), g.lossy_scale(tid)
        cs = mb["colliderStruct"]
        so = mb.get("AudioMaterialPropertiesSO") or {}
        mat = materials[so["guid"]][1] if so.get("guid") in materials else DEFAULT_MATERIAL
        owner = target_of_go.get(go, -1)
        center = f16tof32(half3_bits(cs["Center"]))
        name = g.gameobjects[go]["m_Name"]
        if guid == GUID_AABB:
            rec = np.zeros(1, AABB_DT)
            rec["center"] = f32tof16(center + pos)                                   # Half3.Add(float3, float3)
            rec["size"] = f32tof16(f16tof32(half3_bits(cs["Size"])) * scale)         # Half3.Multiply(float3, float3)
            names["aabb"].append(name)
            dst = aabbs
        elif guid == GUID_OBB:
            rec = np.zeros(1, OBB_DT)
            rbits = np.zeros(3, dtype=np.uint16)                                     # private field: not serialized
            if mb.get("includeGameObjectRotation", 1):
                rbits = half_quat_store(qmul(half_quat_load(rbits), rot))            # Rotation *= transform.rotation
            off = vec(mb.get("rotationEulerOffset", {"x": 0, "y": 0, "z": 0}))
            if np.any(off != 0):
                rbits = half_quat_store(qmul(half_quat_load(rbits), qeuler(off)))
            rec["center"] = f32tof16(qrot(rot, center) + pos)
            rec["size"] = f32tof16(f16tof32(half3_bits(cs["Size"])) * scale)
            rec["rot"] = half_quat_store(qinverse(half_quat_load(rbits)))            # AudioOBBCollider.cs:59
            names["obb"].append(name)
            dst = obbs
        else:
            rec = np.zeros(1, SPHERE_DT)
            rec["center"] = f32tof16(center + pos)
            largest = f16tof32(f32tof16(np.array([max(scale[0], max(scale[1], scale[2]))], dtype=np.float32)))
            rec["radius"] = f32tof16(f16tof32(np.array([cs["Radius"]["value"]], dtype=np.uint16)) * largest)[0]
            names["sphere"].append(name)
            dst = spheres
        rec["absorption"], rec["density"], rec["echo"] = mat
        rec["audioTargetId"] = owner
        dst.append(rec)

    # ---- AudioRayTracer parameters: prefab defaults + scene overrides ------------------------
    pdocs = parse_unity_yaml(os.path.join(ref, "Assets", "Prefabs", "Player.prefab"))
    rt_fid, rt = next((fid, b["MonoBehaviour"]) for c, fid, s, b in pdocs
                      if c == 114 and b["MonoBehaviour"]["m_Script"].get("guid") == GUID_RAYTRACER)
    root_tid = next(fid for c, fid, s, b in pdocs if c == 4 and b["Transform"]["m_GameObject"]["fileID"] == rt["m_GameObject"]["fileID"])
    root_tf = next(b["Transform"] for c, fid, s, b in pdocs if c == 4 and fid == root_tid)
    params = {k: rt[k] for k in ("rayCount", "maxBounces", "maxRayLife", "maxMuffleHitDistance", "muffleEffectiveness",
                                 "mufflePermeationEffectiveness", "permeationStrengthPerRay", "maxReverbDistance")}
    ray_off = vec(rt["rayOrigin"])
    player_pos = vec(root_tf["m_LocalPosition"])
    for c, fid, s, b in docs:
        if c != 1001:
            continue
        for mod in b["PrefabInstance"]["m_Modification"]["m_Modifications"]:
            tgt, path, val = mod["target"]["fileID"], mod["propertyPath"], mod["value"]
            if tgt == rt_fid:
                if path in params:
                    params[path] = float(val)
                elif path.startswith("rayOrigin."):
                    ray_off["xyz".index(path[-1])] = float(val)
            elif tgt == root_tid and path.startswith("m_LocalPosition."):
                player_pos["xyz".index(path[-1])] = float(val)
    ray_origin = player_pos.astype(np.float32) + ray_off.astype(np.float32)          # ART:165

    threads = 1
    for fid, mb in behaviours:
        if mb["m_Script"].get("guid") == GUID_RT_MANAGER:
            threads = int(mb.get("maxThreadCount", 1))                                # ARM:16-19 (min with worker count)

    cat = lambda lst, dt: np.concatenate(lst).astype(dt) if lst else np.zeros(0, dt)
    return dict(aabbs=cat(aabbs, AABB_DT), obbs=cat(obbs, OBB_DT), spheres=cat(spheres, SPHERE_DT),
                targets=np.array(targets, dtype=np.float32).reshape(-1, 3), ray_origin=ray_origin,
                params=params, threads=threads, names=names)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c1_demo_scene.npz"))
    a = ap.parse_args()
    e = export(a.reference)
    p = e["params"]
    print(f"AABB {len(e['aabbs'])}  OBB {len(e['obbs'])}  spheres {len(e['spheres'])}  targets {len(e['targets'])}")
    print("ray origin", e["ray_origin"], "params", p, "threads", e["threads"])
    np.savez_compressed(
        a.out, aabbs=e["aabbs"].view(np.uint8), obbs=e["obbs"].view(np.uint8), spheres=e["spheres"].view(np.uint8),
        targets=e["targets"], ray_origin=e["ray_origin"],
        params=np.array([p["maxRayLife"], int(p["maxBounces"]) + 1, p["maxMuffleHitDistance"], p["permeationStrengthPerRay"],
                         p["muffleEffectiveness"], p["mufflePermeationEffectiveness"], p["maxReverbDistance"], e["threads"]],
                        dtype=np.float64),
        ray_count=np.array([int(p["rayCount"])], dtype=np.int64))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
