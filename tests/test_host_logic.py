"""Host-side logic that needs no GPU: partial-result blobs (merge / finalize = ProcessAudioDataJob formulas,
slot-row mapping, quirks Q2/Q5/Q6/Q8) exercised through the C ABI's context-free entry points, and the
ray-shard index mapping."""
import ctypes as C

import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes
from helpers import BLOB_MAGIC, blob_dtype, echo_fixed_sums, load_golden


def blob_from_oracle(scene, frame, oracle, rays=None, last_hit_owner=True):
    """Build the blob a context would export for `rays` (global indices, default all) from oracle outputs."""
    Na, T, N, H = scene.n_targets, scene.batch_count, scene.n_rays, scene.max_hits_per_ray
    b = np.zeros(1, blob_dtype(Na, T))
    b["magic"], b["nTargets"], b["batchCount"], b["shards"] = BLOB_MAGIC, Na, T, 1
    rays = np.arange(N) if rays is None else np.asarray(rays)
    echo = frame.echo.reshape(N, H)[rays].ravel()
    lo, hi, zeros = echo_fixed_sums(echo)
    b["fixedLo"], b["fixedHi"], b["zeros"], b["entries"] = lo, hi, zeros, echo.size
    return b, rays


def test_blob_size_matches_library(art_lib):
    for Na, T in [(1, 1), (3, 2), (64, 8), (256, 5), (7, 3)]:
        assert blob_dtype(Na, T).itemsize == art_lib.art_partials_size(Na, T)


def test_finalize_reproduces_process_audio_data_job(art_lib, oracle):
    """art_finalize == PA:32-76 on the oracle's own arrays (exact reduction within 1e-6 of the FP64 evaluation,
    sequential-FP32 mode bit-exact with the reference-faithful evaluation)."""
    for name in ("c2_n768_t3", "c3_n48_t2", "c2_n512_gated"):
        s, g = load_golden(name)
        f = oracle.run_frame(s)
        Na, T, N, H = s.n_targets, s.batch_count, s.n_rays, s.max_hits_per_ray
        b, _ = blob_from_oracle(s, f, oracle)
        # per-batch muffle counts: batch k == slot row k for these sizes (RT:63-64)
        bs = oracle.batch_size(N, T)
        rows = [(k * bs * T) // N for k in range((N + bs - 1) // bs)]
        assert rows == list(range(len(rows)))
        b["muffleCounts"][0, :] = f.muffle.astype(np.uint32)
        # canonical permeation (Q5/Q6): slot row 0 holds the values of the last hitting ray of the LAST batch
        b["lastHitRay"][0, :] = -1
        b["lastHitRay"][0, len(rows) - 1] = N - 1
        b["permLast"][0, (len(rows) - 1) * Na:(len(rows)) * Na] = f.permeation[:Na]
        # sequential FP32 sum exactly as PA:40-48 (numpy cumsum on float32 is a sequential running sum)
        e = f.echo.view(np.float16).astype(np.float32)
        b["seqTotal"] = np.cumsum(e[e != 0], dtype=np.float32)[-1] if (e != 0).any() else 0.0
        b["seqZeros"] = np.float32(min(int((e == 0).sum()), 1 << 24))
        b["seqValid"] = 1
        blob = b.view(np.uint8).ravel()
        r = native.finalize(blob, s, N)
        np.testing.assert_array_equal(r.muffle, f.muffle)
        np.testing.assert_array_equal(r.permeation.view(np.uint32), f.permeation.view(np.uint32))
        for k in ("muffleStrength", "reverbStrength", "reverbVolume"):
            np.testing.assert_allclose(r.settings[k], f.settings_fp64[k], rtol=0, atol=1e-6)
        r2 = native.finalize(blob, s, N, flags=native.FRAME_REVERB_SEQ_FP32)
        np.testing.assert_array_equal(r2.settings.view(np.uint8), f.settings.view(np.uint8))


def test_merge_is_exact_and_order_independent(art_lib):
    rng = np.random.default_rng(3)
    Na, T = 5, 4
    blobs = []
    for i in range(4):
        b = np.zeros(1, blob_dtype(Na, T))
        b["magic"], b["nTargets"], b["batchCount"], b["shards"] = BLOB_MAGIC, Na, T, 1
        b["fixedLo"], b["fixedHi"] = rng.integers(-10**9, 10**9), rng.integers(0, 10**12)
        b["zeros"], b["entries"] = rng.integers(0, 10**6), 10**6
        b["muffleCounts"][0] = rng.integers(0, 10**6, T * Na)
        b["lastHitRay"][0] = rng.integers(-1, 1000, T)
        b["permLast"][0] = rng.normal(size=T * Na)
        b["permSumInt"][0] = rng.integers(0, 10**9, Na)
        b["permSumFrac"][0] = rng.integers(0, 2**36, Na)
        blobs.append(b.view(np.uint8).ravel())
    m1 = native.merge_partials(blobs).view(blob_dtype(Na, T))
    m2 = native.merge_partials(blobs[::-1]).view(blob_dtype(Na, T))
    views = [x.view(blob_dtype(Na, T)) for x in blobs]
    assert m1["shards"][0] == 4
    assert m1["fixedHi"][0] == sum(int(v["fixedHi"][0]) for v in views)
    np.testing.assert_array_equal(m1["muffleCounts"][0], sum(v["muffleCounts"][0].astype(np.uint64) for v in views))
    # max-by-ray-index select of the last hitting ray per batch
    for k in range(T):
        owner = int(np.argmax([v["lastHitRay"][0][k] for v in views]))
        assert m1["lastHitRay"][0][k] == views[owner]["lastHitRay"][0][k]
        if views[owner]["lastHitRay"][0][k] >= 0:
            np.testing.assert_array_equal(m1["permLast"][0][k * Na:(k + 1) * Na], views[owner]["permLast"][0][k * Na:(k + 1) * Na])
    ties = any(len({int(v["lastHitRay"][0][k]) for v in views}) < 4 for k in range(T))
    if not ties:
        np.testing.assert_array_equal(m1.view(np.uint8), m2.view(np.uint8))
    assert m1["seqValid"][0] == 0          # a sequential FP32 sum cannot be merged


def test_merge_rejects_mismatched_blobs(art_lib):
    a = np.zeros(1, blob_dtype(2, 2)); a["magic"], a["nTargets"], a["batchCount"] = BLOB_MAGIC, 2, 2
    b = np.zeros(1, blob_dtype(3, 2)); b["magic"], b["nTargets"], b["batchCount"] = BLOB_MAGIC, 3, 2
    with pytest.raises(native.ArtError):
        native.merge_partials([a.view(np.uint8).ravel(), b.view(np.uint8).ravel()[:a.itemsize]])


def test_u16_wrap_and_slot_rows_in_finalize(art_lib):
    """Q8: the u16 table holds count mod 65536 and PA sums the WRAPPED values; Q6: permeation goes to slot row 0."""
    s = scenes.make_config("c2", n_rays=64, batch_count=2)
    Na, T, N = 1, 2, 64
    b = np.zeros(1, blob_dtype(Na, T))
    b["magic"], b["nTargets"], b["batchCount"], b["shards"] = BLOB_MAGIC, Na, T, 1
    b["muffleCounts"][0] = [70000, 5]
    b["lastHitRay"][0] = [31, 63]
    b["permLast"][0] = [11.0, 22.0]
    b["zeros"], b["entries"] = 64 * 8, 64 * 8
    r = native.finalize(b.view(np.uint8).ravel(), s, N)
    assert list(r.muffle) == [70000 % 65536, 5]
    assert list(r.muffle_totals) == [70005]
    assert list(r.permeation) == [22.0, 0.0]            # both batches write slot row 0; the last batch wins (Q5/Q6)
    expect = 1.0 - np.float32(70000 % 65536 + 5) / np.float32(64 * 8) * np.float32(1.0)
    perm = np.float32(22.0) / np.float32(64) / np.float32(1.0) * np.float32(0.5)
    assert r.settings["muffleStrength"][0] == np.float32(max(0.0, min(1.0, np.float32(expect) - perm)))


def test_shard_index_mapping_matches_header_formula():
    """art_set_ray_shard: local j -> global ((j / chunk) * G + r) * chunk + j % chunk covers every ray once."""
    N, G, chunk = 1000, 3, 64
    seen = np.zeros(N, np.int32)
    for r in range(G):
        n_chunks = (N + chunk - 1) // chunk
        n_local = sum(min(chunk, N - c * chunk) for c in range(r, n_chunks, G))
        j = np.arange(n_local)
        g = ((j // chunk) * G + r) * chunk + j % chunk
        assert g.max() < N
        seen[g] += 1
    assert (seen == 1).all()
