"""aukit_cuda_preloader_* (pipelined host path): same results as the synchronous end-to-end call and
as the oracle, for clips of different sizes rotating through the device slots."""
import numpy as np
import pytest

from util import TOL, tone_s16

pytestmark = pytest.mark.gpu


def _clip(frames, seed):
    rng = np.random.default_rng(seed)
    x = tone_s16(frames, 2, seed=seed).astype(np.int32) + rng.integers(-300, 300, (frames, 2))
    return np.clip(x, -32768, 32767).astype("<i2")


def test_preloader_matches_sync_call_and_oracle(ak, O):
    sizes = [44100, 1000, 30011, 7, 44100, 12345]
    clips = [_clip(n, i) for i, n in enumerate(sizes)]
    max_in = max(c.nbytes for c in clips)
    max_out = max(int(ak.context().lib.aukit_resample_out_len(n, 44100.0, 48000.0)) for n in sizes)
    pl = ak.Preloader(max_in, max_out, slots=2)
    ins, outs = [], []
    for c in clips:
        h = ak.Preloader.pinned((c.nbytes,), np.uint8)
        h[:] = c.reshape(-1).view(np.uint8)
        n_out = int(ak.context().lib.aukit_resample_out_len(len(c), 44100.0, 48000.0))
        o = ak.Preloader.pinned((1, n_out), np.float32)
        o[:] = np.nan
        ins.append(h)
        outs.append(o)
        pl.submit(h, o, peakAmplitude=0.8, sampleRate=44100, targetRate=48000, interpolation="cubic", mono=True)
    pl.drain()
    for c, o in zip(clips, outs):
        sync = ak.preload(c.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)
        assert np.array_equal(o, sync)                                  # same kernels, same bits
    c = clips[2]
    dec = O.pcm(c.tobytes(), 16, "signed", 2, True, False)
    ref = O.normalize(O.mono(O.resample(dec, 44100, 48000, "cubic")), 0.8, False)
    assert float(np.max(np.abs(outs[2][0] - ref[0]))) <= TOL
    pl.close()


def test_preloader_split_phases_and_errors(ak):
    c = _clip(5000, 3)
    pl = ak.Preloader(c.nbytes, 2 * 6000, slots=2)        # samples = output channels x frames
    h = ak.Preloader.pinned((c.nbytes,), np.uint8)
    h[:] = c.reshape(-1).view(np.uint8)
    d = pl.describe(c.nbytes, sampleRate=44100, targetRate=48000, interpolation="linear", mono=False)
    out = ak.Preloader.pinned((2, d.n_out), np.float32)
    k = pl.begin(h, d)
    assert pl.peak_ptr(k) != 0 and pl.stream != 0
    pl.finish(k, out, 0.5)
    pl.drain()
    sync = ak.preload(c.tobytes(), 16, "signed", 2, 44100, 48000, "linear", False, 0.5)
    assert np.array_equal(out, sync)
    with pytest.raises(ak.AukitError, match="larger than the preloader"):
        big = np.zeros(c.nbytes * 2, dtype=np.uint8)
        pl.submit(big, out, sampleRate=44100)
    with pytest.raises(ak.AukitError, match="was not begun"):
        pl.finish(1, out)
    pl.close()
