"""World-size-2 `gloo` tests of the multi-GPU host logic (no GPU): the time-shard planner, the halo
windows and the one collective of the path (all-reduce MAX of the per-shard peaks).  The per-shard
arithmetic is stood in for by the numpy restatement of the reference (tests/util.py), so the test
checks exactly what the sharding layer adds: partitioning, halos, and the exchange."""
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
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, src, dst, interp, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aukit_b200.sharding import allreduce_max_, plan_time_shards, shard_clips
        from util import ref_resample_window
        rng = np.random.default_rng(123)                    # same signal on every rank
        n_in = 50_000
        x = rng.uniform(-1, 1, (2, n_in))
        shards = plan_time_shards(n_in, src, dst, interp, world)
        sh = shards[rank]
        # contiguous, disjoint, complete output partition
        assert shards[0].out_first == 0 and all(shards[i].out_first + shards[i].n_out == shards[i + 1].out_first for i in range(world - 1))
        window = x[:, sh.in_first: sh.in_first + sh.in_count]          # this rank holds ONLY its shard + halo
        local = ref_resample_window(window, sh.in_first, n_in, src, dst, sh.out_first, sh.n_out, interp)
        mono = (local[0] + local[1]) / 2
        peak = torch.tensor([np.max(np.abs(mono))], dtype=torch.float64)
        allreduce_max_(peak)                                            # the path's only collective
        out = np.clip(mono * (0.8 / peak.item()), -1, 1)
        # clip sharding: disjoint cover, greedy balance
        mine = shard_clips(11, world, rank)
        sized = shard_clips(6, world, rank, sizes=[9, 1, 1, 1, 4, 4])
        q.put((rank, sh.out_first, out, peak.item(), mine, sized))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("src,dst,interp", [(44100, 48000, "cubic"), (96000, 44100, "linear"), (44100, 48000, "none")])
def test_time_sharded_normalize_over_gloo(O, src, dst, interp):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, src, dst, interp, q)) for r in range(world)]
    [p.start() for p in procs]
    results = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process reference: the C oracle on the whole buffer
    rng = np.random.default_rng(123)
    x = rng.uniform(-1, 1, (2, 50_000))
    whole = O.normalize(O.mono(O.resample(x, src, dst, interp)), 0.8)[0]
    stitched = np.concatenate([r[2] for r in results])
    assert stitched.shape == whole.shape
    assert np.max(np.abs(stitched - whole)) <= 1e-14          # same fp64 expressions; pow() vs ** last-ulp only
    assert results[0][3] == results[1][3]                      # both ranks ended with the global peak
    assert sorted(results[0][4] + results[1][4]) == list(range(11))
    assert sorted(results[0][5] + results[1][5]) == list(range(6))
    loads = [sum([9, 1, 1, 1, 4, 4][i] for i in r[5]) for r in results]
    assert max(loads) - min(loads) <= 2


def test_shard_planners_are_aligned_and_complete(O):
    """plan_time_shards: interior boundaries on the fused kernels' warp tile (5120 outputs at 44.1 -> 48 kHz), windows
    padded but always covering the needed halo; plan_block_shards / aukit_block_shard: contiguous, complete, near-equal."""
    import ctypes as C
    from aukit_b200 import _lib
    from aukit_b200.sharding import plan_block_shards, plan_time_shards, shard_alignment
    lib = _lib.load()
    assert shard_alignment(44100, 48000) == 5120 and shard_alignment(22050, 48000) == 32 * 320
    assert shard_alignment(96000, 48000) == 4 and shard_alignment(44056.5, 48000) == 4
    n_in = 8 * 3600 * 44100
    for world in (1, 2, 4, 8):
        sh = plan_time_shards(n_in, 44100, 48000, "cubic", world)
        n_out = int(lib.aukit_resample_out_len(n_in, 44100.0, 48000.0))
        assert sh[0].out_first == 0 and sh[-1].out_first + sh[-1].n_out == n_out
        for a, b in zip(sh, sh[1:]):
            assert a.out_first + a.n_out == b.out_first and b.out_first % 5120 == 0
        for s in sh:
            f, c = C.c_uint64(0), C.c_uint64(0)
            assert lib.aukit_resample_window(n_in, 44100.0, 48000.0, 2, s.out_first, s.n_out, C.byref(f), C.byref(c)) == 0
            assert s.in_first <= f.value and s.in_first + s.in_count >= f.value + c.value      # halo covered
            assert s.in_first + s.in_count <= n_in and (s.in_first % 4 == 0 or s.in_first == f.value)
            assert f.value - s.in_first <= 12 and s.in_first + s.in_count - (f.value + c.value) <= 8
    for nb, world in ((778_236, 8), (779_765, 3), (5, 8), (0, 2)):
        parts = plan_block_shards(nb, world)
        assert parts[0][0] == 0 and sum(c for _, c in parts) == nb
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
