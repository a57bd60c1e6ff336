"""N>1 host logic on CPU: two gloo ranks each encode their GOP range (with the oracle standing
in for the per-GPU encoder) and rank 0 concatenates; the result must equal the unsharded stream."""
import os
import sys

import pytest
import torch.multiprocessing as mp

from video_codec_pipeline_b200.shard import gop_ranges


def test_gop_ranges_partition():
    for nframes, gop, world in ((300, 60, 2), (300, 60, 8), (3840, 60, 8), (61, 60, 4), (5, 60, 2), (600, 60, 4)):
        r = gop_ranges(nframes, gop, world)
        assert len(r) == world
        assert sum(x[1] for x in r) == nframes
        pos = 0
        for f0, n, g0 in r:
            if n:
                assert f0 == pos and f0 % gop == 0 and g0 == f0 // gop
            pos += n
        counts = [(x[1] + gop - 1) // gop for x in r]
        assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from oracle import pyoracle
    from video_codec_pipeline_b200 import synth
    from video_codec_pipeline_b200.shard import gather_streams, gop_ranges
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, n, gop = 96, 64, 10, 3
    clip = synth.make_clip(w, h, n, seed=9)
    f0, cnt, g0 = gop_ranges(n, gop, world)[rank]
    local = b""
    if cnt:
        p = pyoracle.make_params(w, h, gop=gop, qp_i=26, qp_p=28, first_gop=g0)
        local = pyoracle.encode(p, clip[f0:f0 + cnt], want_recon=False)["stream"]
    whole = gather_streams(local, rank, world)
    if rank == 0:
        ref = pyoracle.encode(pyoracle.make_params(w, h, gop=gop, qp_i=26, qp_p=28), clip, want_recon=False)["stream"]
        q.put(whole == ref)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gop_sharding_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
    assert ok
    assert all(p.exitcode == 0 for p in procs)
