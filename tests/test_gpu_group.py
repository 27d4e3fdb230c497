"""Multi-GPU from ONE host thread (aukit_cuda_group_*, csrc/comm.cu): the MAX exchange over peer-mapped memory and the
time-sharded fused chain must give the bits of a single GPU.  A group may list the same device several times (each
member has its own context and stream), so the exchange kernel, its epoch ring and the shard planner are exercised on a
one-GPU box too; with two or more GPUs visible the same tests also run across devices."""
import ctypes as C

import numpy as np
import pytest

from util import TOL, f32_equal_bits, tone_s16

pytestmark = pytest.mark.gpu


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.mark.parametrize("kind", ["tone", "noise"])
def test_group_preload_equals_single_gpu(ak, O, kind):
    n = 6 * 44100 + 1234
    pcm = tone_s16(n, 2, 44100, seed=3) if kind == "tone" else np.random.default_rng(3).integers(-32768, 32768, (n, 2)).astype(np.int16)
    whole = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8)
    assert np.max(np.abs(whole[0] - ref)) <= TOL
    for devs in _device_lists():
        g = ak.Group(devs)
        try:
            for _ in range(3):                                   # several exchanges: the epoch ring wraps
                got = g.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)
                assert f32_equal_bits(got, whole), devs
            a = g.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8, owner=ak.context())
            assert a.sampleRate == 48000 and f32_equal_bits(a.numpy(), whole), devs     # gathered device to device
            del a
        finally:
            g.close()


def test_group_preload_other_shapes(ak, O):
    rng = np.random.default_rng(4)
    g = ak.Group([0, 0, 0])
    try:
        x = (rng.standard_normal((50001, 8)) * 0.3).astype("<f4")                # config 5 / 5': f32 x 8, not mono
        for dst in (48000, 44100):
            got = g.preload(x.tobytes(), 32, "float", 8, 96000, dst, "cubic", False, 1.0)
            assert f32_equal_bits(got, ak.preload(x.tobytes(), 32, "float", 8, 96000, dst, "cubic", False, 1.0))
        s = rng.integers(0, 256, 30000 * 6, dtype=np.uint8)                      # s24 big-endian stereo, linear
        got = g.preload(s.tobytes(), 24, "signed", 2, 22050, 48000, "linear", True, 0.5, True)
        assert f32_equal_bits(got, ak.preload(s.tobytes(), 24, "signed", 2, 22050, 48000, "linear", True, 0.5, True))
        tiny = rng.integers(-3000, 3000, (7, 2)).astype("<i2")                   # fewer outputs than members: empty shards
        got = g.preload(tiny.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)
        assert f32_equal_bits(got, ak.preload(tiny.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8))
    finally:
        g.close()


@pytest.mark.parametrize("independent", [False, True])
def test_group_normalize_on_sharded_audio(ak, O, independent):
    """effects.normalize (A:3431) where the Audio is held as time shards, one per member."""
    rng = np.random.default_rng(6)
    x = (rng.uniform(-1, 1, (3, 90001)) * np.array([[0.2], [0.7], [0.4]])).astype(np.float32)
    ref = O.normalize(x.astype(np.float64), 0.9, independent)
    for devs in _device_lists():
        g = ak.Group(devs)
        try:
            cuts = [x.shape[1] * i // g.size for i in range(g.size + 1)]
            shards = []
            for i, ctx in enumerate(g.contexts):
                ctx.make_current()
                shards.append(ak.Audio.from_numpy(x[:, cuts[i]: cuts[i + 1]], 48000, ctx))
            g.normalize(shards, 0.9, independent)
            parts = []
            for ctx, a in zip(g.contexts, shards):
                ctx.make_current()
                parts.append(a.numpy())
            got = np.concatenate(parts, axis=1)
            assert np.max(np.abs(got - ref)) <= TOL
            ak.context().make_current()          # several devices in one process: the context in use is made current (include/aukit_cuda.h)
            one = ak.effects.normalize(ak.Audio.from_numpy(x, 48000), 0.9, independent).numpy()
            assert f32_equal_bits(got, one)
            del shards
        finally:
            ak.context().make_current()
            g.close()
