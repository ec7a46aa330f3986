"""Golden fixtures: frozen oracle outputs (tests/golden/make_golden.py). CPU leg pins the oracle; GPU leg checks
the CUDA path against the same files without needing the oracle at all."""
import numpy as np
import pytest

from helpers import load_golden

CASES = ["c2_n768_t3", "c3_n48_t2", "c4_n24_t1", "c2_n512_gated", "c1_demo", "c1_demo_1src_t3"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle, name):
    s, g = load_golden(name)
    f = oracle.run_frame(s)
    for k in ("echo", "hit_points", "hit_counts", "hit_ids", "muffle", "muffle_totals"):
        np.testing.assert_array_equal(getattr(f, k), g[k], err_msg=k)
    np.testing.assert_array_equal(f.permeation.view(np.uint32), g["permeation"].view(np.uint32))
    np.testing.assert_array_equal(f.settings.view(np.uint8), g["settings"].view(np.uint8))
    c = f.counters
    got = [c["segments"], c["segment_hits"]] + c["trace_tests"] + c["echo_tests"] + c["muffle_tests"]
    np.testing.assert_array_equal(np.array(got, dtype=np.uint64), g["counters"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_reproduces_golden(gpu_ctx, name):
    from audio_raytracer_b200 import native
    s, g = load_golden(name)
    native.upload(gpu_ctx, s)
    r = gpu_ctx.run_frame(s, flags=native.FRAME_COUNTERS)
    for k in ("echo", "hit_counts", "hit_ids", "muffle", "muffle_totals"):
        np.testing.assert_array_equal(getattr(r, k), g[k], err_msg=k)
    a, b = r.hit_points.copy(), g["hit_points"].copy()
    a[a == 0x8000] = 0
    b[b == 0x8000] = 0
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(r.permeation.view(np.uint32), g["permeation"].view(np.uint32))
    for f in ("muffleStrength", "reverbStrength", "reverbVolume"):
        np.testing.assert_allclose(r.settings[f], g["settings_fp64"][f], rtol=0, atol=1e-6)
    c = r.counters
    got = [c["segments"], c["segmentHits"]] + c["traceTests"] + c["echoTests"] + c["muffleTests"]
    np.testing.assert_array_equal(np.array(got, dtype=np.uint64), g["counters"])
    # ART_FRAME_REVERB_SEQ_FP32: the reference's own FP32 rounding, bit for bit
    r2 = gpu_ctx.run_frame(s, flags=native.FRAME_REVERB_SEQ_FP32)
    np.testing.assert_array_equal(r2.settings.view(np.uint8), g["settings"].view(np.uint8))
