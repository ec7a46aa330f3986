"""The C oracle against a second, independently written restatement of the reference (oracle/independent.py, numpy float32
on whole ray x collider planes): both must agree BIT FOR BIT on the committed golden scenes. The reference itself cannot run
here or on the GPU box (C#/Unity, no dotnet / mono: profiles/r02_dotnet_probe.txt), so this does not pin the oracle to the
reference -- it removes the risk of a transcription slip in either restatement (loop order, strict '<', the quirks Q3/Q4/Q5/Q7,
operation order of every FP32 expression)."""
import numpy as np
import pytest

from helpers import load_golden


def canonical_permeation(scene, hit, vals):
    """PM:36-46, 85 in canonical serial batch order: every batch zeroes its slot row, every hitting ray overwrites it."""
    N, Na, T = scene.n_rays, scene.n_targets, scene.batch_count
    out = np.zeros(T * Na, np.float32)
    b = int(max(1.0, np.ceil(np.float32(N) / np.float32(T))))
    ray_of = np.nonzero(hit)[0]
    for start in range(0, N, b):
        total = min(b, N - start)
        pbc = (T * Na) // total // Na                      # PermeationPowerRemains.Length / totalRays / TotalAudioTargets
        row = (start * pbc) // N
        out[row * Na:(row + 1) * Na] = 0
        sel = np.nonzero((ray_of >= start) & (ray_of < start + total))[0]
        if sel.size:
            out[row * Na:(row + 1) * Na] = vals[sel[-1]]
    return out


@pytest.mark.parametrize("name", ["c2_n768_t3", "c3_n48_t2", "c4_n24_t1", "c2_n512_gated", "c1_demo", "c1_demo_1src_t3"])
def test_independent_restatement_equals_the_c_oracle(oracle, name):
    from oracle import independent as ind
    s, g = load_golden(name)
    f = oracle.run_frame(s)
    t = ind.trace(s)
    np.testing.assert_array_equal(t["hit_ids"], f.hit_ids)
    np.testing.assert_array_equal(t["hit_counts"], f.hit_counts)
    np.testing.assert_array_equal(t["echo"], f.echo)
    a, b = t["hit_points"].copy(), f.hit_points.copy()
    a[a == 0x8000] = 0; b[b == 0x8000] = 0                 # (-0 and +0 half)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(t["muffle"], f.muffle)
    np.testing.assert_array_equal(t["muffle_totals"], f.muffle_totals)
    assert t["segments"] == f.counters["segments"]
    hit, vals = ind.permeation(s)
    assert int(hit.sum()) == f.counters["perm_hit_rays"]
    np.testing.assert_array_equal(canonical_permeation(s, hit, vals).view(np.uint32), f.permeation.view(np.uint32))
    sums = np.cumsum(vals.astype(np.float64), axis=0)[-1] if len(vals) else np.zeros(s.n_targets)
    np.testing.assert_array_equal(sums, f.permeation_sum)   # sequential double sum over rays, as the C oracle accumulates it


def test_f32tof16_independent_vs_c(oracle):
    from oracle import independent as ind
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * np.float32(300.0),
                        (rng.integers(0, 0x7F800000, 20000, dtype=np.uint32)).view(np.float32),
                        np.float32([0.0, -0.0, 1.0 + 2.0 ** -11, 65504.0, 65519.9, 65520.0, 1e9, np.inf, -np.inf, 6e-8, 5.96e-8, 2.98e-8])])
    got = ind.f32_to_f16(x)
    want = np.array([oracle.f32tof16(float(v)) for v in x], dtype=np.uint16)
    np.testing.assert_array_equal(got, want)
    # the overflow range follows the package's COMMENT ("clamp to signed infinity"): |x| >= 65520 -> +-Inf
    assert ind.f32_to_f16(np.float32([65520.0, 1e9]))[0] == 0x7C00 and ind.f32_to_f16(np.float32([65520.0, 1e9]))[1] == 0x7C00
    back = ind.f16_to_f32(got[:20000])
    assert np.all(np.abs(back - x[:20000]) <= np.abs(x[:20000]) * 2.0 ** -11 + 1e-7)
