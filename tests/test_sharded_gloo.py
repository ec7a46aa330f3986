"""world_size-2 on CPU (gloo): the N>1 host path of bench.py -- every rank builds the partial blob of its
interleaved ray shard, blobs are all-gathered, merged and finalised; the result must equal the single-rank frame.
The per-shard inputs come from the oracle here (no GPU in this test); on the GPU box the same exchange runs over
NCCL with blobs produced by libaudiort_cuda (tests/test_gpu_parity.py::test_sharded_contexts...)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_raytracer_b200 import native, scenes
    from helpers import BLOB_MAGIC, blob_dtype, echo_fixed_sums
    from oracle import oracle as orc
    s = scenes.make_config("c2", n_rays=512, batch_count=world)
    Na, T, N, H = s.n_targets, s.batch_count, s.n_rays, s.max_hits_per_ray
    chunk = 64
    bs = orc.batch_size(N, T)
    b = np.zeros(1, blob_dtype(Na, T))
    b["magic"], b["nTargets"], b["batchCount"], b["shards"] = BLOB_MAGIC, Na, T, 1
    b["lastHitRay"][0] = -1
    lo = hi = zeros = entries = 0
    for c in range(rank, (N + chunk - 1) // chunk, world):        # this rank's interleaved chunks
        first, count = c * chunk, min(chunk, N - c * chunk)
        f = orc.trace_range(s, first, count)
        k = first // bs
        b["muffleCounts"][0][k * Na:(k + 1) * Na] += f.muffle_totals
        e = f.echo.reshape(N, H)[first:first + count].ravel()
        l, h, z = echo_fixed_sums(e)
        lo, hi, zeros, entries = lo + l, hi + h, zeros + z, entries + e.size
        if first + count == N:                                     # owner of the last ray: canonical permeation values
            full = orc.run_frame(s, jobs=orc.JOB_PM)
            b["lastHitRay"][0][T - 1] = N - 1
            b["permLast"][0][(T - 1) * Na:T * Na] = full.permeation[:Na]
    b["fixedLo"], b["fixedHi"], b["zeros"], b["entries"] = lo, hi, zeros, entries
    mine = torch.from_numpy(b.view(np.uint8).ravel().copy())
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    merged = native.merge_partials([g.numpy() for g in gathered])
    r = native.finalize(merged, s, N)
    if rank == 0:
        np.savez(out_path, muffle=r.muffle, permeation=r.permeation, settings=r.settings.view(np.uint8),
                 totals=r.muffle_totals)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_matches_single_rank(tmp_path, oracle, art_lib):
    from audio_raytracer_b200 import scenes
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    s = scenes.make_config("c2", n_rays=512, batch_count=2)
    f = oracle.run_frame(s)
    np.testing.assert_array_equal(z["muffle"], f.muffle)
    np.testing.assert_array_equal(z["totals"], f.muffle_totals)
    np.testing.assert_array_equal(z["permeation"].view(np.uint32), f.permeation.view(np.uint32))
    from audio_raytracer_b200.layouts import SETTINGS_DT
    st = z["settings"].view(SETTINGS_DT)
    for k in ("muffleStrength", "reverbStrength", "reverbVolume"):
        np.testing.assert_allclose(st[k], f.settings_fp64[k], rtol=0, atol=1e-6)
