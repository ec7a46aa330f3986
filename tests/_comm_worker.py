"""One rank of the multi-process sharding test (tests/test_gpu_multi_device.py): joins the library's communicator
(art_comm_init), traces its shard and writes what art_complete returned.   python _comm_worker.py RANK WORLD IDFILE OUT"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_raytracer_b200 import native, scenes  # noqa: E402


def main():
    rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    if rank == 0:
        uid = native.comm_unique_id()
        with open(idfile + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise SystemExit("no communicator id")
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    s = scenes.make_config("c3", n_rays=40000, batch_count=world)
    with native.Context(device=rank) as ctx:
        ctx.comm_init(uid, rank, world, 256)
        native.upload(ctx, s)
        r = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)          # merged and finalised inside the library, on every rank
        r2 = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
        assert r.counters["devicesUsed"] == world
        np.savez(out, muffle=r.muffle, muffle_totals=r.muffle_totals, permeation=r.permeation, permeation_sum=r.permeation_sum,
                 settings=r.settings.view(np.uint8), echo=r.echo, hit_counts=r.hit_counts, segments=r.counters["segments"],
                 exchange_ms=r2.counters["exchangeMs"], muffle2=r2.muffle, n_local=ctx.local_ray_count())


if __name__ == "__main__":
    main()
