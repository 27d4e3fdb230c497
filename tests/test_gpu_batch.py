"""BASELINE config 3 on the device: a batch of clips through the fused decode -> resample -> amplify kernel (K14,
csrc/pipeline_tile.cu) must equal, per clip, aukit.pcm (A:1049) -> Audio:resample (A:653) -> effects.amplify (A:3356):
within 2^-20 of the oracle's double-precision chain, and bit for bit what the three separate device kernels give."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from util import TOL, f32_equal_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_chain(O, raw, bits, dt, ch, be, src, dst, interp, mult):
    d = O.pcm(raw, bits, dt, ch, True, be)
    return O.amplify(O.resample(d, src, dst, interp), mult)


def test_config3_reference_vectors_through_the_fused_batch(ak):
    """The reference's own outputs (tests/golden, chain3_*): three 0.05 s s24-BE stereo clips, one per rate."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    manifest = json.loads(z["manifest"].tobytes().decode())
    cases = [(i, m) for i, m in enumerate(manifest) if m["op"] == "chain3" and m["args"]["interpolation"] == "cubic"]
    assert len(cases) == 3
    clips = [z["c%d/in" % i].tobytes() for i, _ in cases]
    rates = [m["args"]["sampleRate"] for _, m in cases]
    outs = ak.preload_clips(clips, rates, 24, "signed", 2, 48000, "cubic", 0.5, True)
    for (i, m), a in zip(cases, outs):
        got = a.numpy()
        for c in range(2):
            ref = z["c%d/out%d" % (i, c)]
            assert got[c].shape == ref.shape
            assert np.max(np.abs(got[c] - ref)) <= TOL, m["name"]
        assert a.sampleRate == 48000


@pytest.mark.parametrize("bits,dt,be,ch", [(24, "signed", True, 2), (16, "signed", False, 2), (24, "signed", False, 1), (8, "unsigned", False, 2),
                                           (16, "signed", True, 1), (32, "signed", False, 2), (8, "signed", False, 1)])
@pytest.mark.parametrize("interp,mult", [("cubic", 0.5), ("linear", 1.7), ("cubic", 0.3)])
def test_batch_matches_oracle_and_the_unfused_kernels(ak, O, bits, dt, be, ch, interp, mult):
    """Clips long enough for interior (bulk-copied) tiles, ragged lengths, every rate class of config 3 plus 11.025 kHz;
    full-scale noise so that the clamp of A:668 acts."""
    rng = np.random.default_rng(bits * 100 + ch)
    rates = [22050, 44100, 96000, 44100, 11025, 96000, 22050]
    frames = [41017, 90001, 200003, 3, 30000, 1, 50000]
    B = bits // 8
    clips = [rng.integers(0, 256, n * ch * B, dtype=np.uint8).tobytes() for n in frames]
    outs = ak.preload_clips(clips, rates, bits, dt, ch, 48000, interp, mult, be)
    assert len(outs) == len(clips)
    for raw, src, a in zip(clips, rates, outs):
        ref = _oracle_chain(O, raw, bits, dt, ch, be, src, 48000, interp, mult)
        got = a.numpy()
        assert got.shape == ref.shape, (src, got.shape, ref.shape)
        if ref.size:
            assert np.max(np.abs(got - ref)) <= TOL, (src, len(raw))
        unf = ak.effects.amplify(ak.pcm(raw, bits, dt, ch, src, True, be).resample(48000, interp), mult).numpy()
        assert f32_equal_bits(got, unf), "fused batch differs from pcm -> resample -> amplify on the device (rate %d)" % src


def test_batch_device_api_unaligned_clips_and_caller_layout(ak, O):
    """aukit_cuda_dev_batch_resample_amplify on caller-owned device memory: clip starts at odd byte offsets (the bulk
    copies round down to 16 bytes and carry the shift), caller-chosen output rows."""
    torch = pytest.importorskip("torch")
    ctx = ak.context()
    rng = np.random.default_rng(5)
    rates = [44100, 22050, 96000, 44100]
    frames = [70001, 33333, 150001, 25000]
    pad = [6, 2, 11, 0]
    blob, clips = bytearray(), (ak.Clip * 4)()
    raws = []
    for k in range(4):
        blob += bytes(pad[k])
        raw = rng.integers(0, 256, frames[k] * 6, dtype=np.uint8).tobytes()
        raws.append(raw)
        clips[k].in_offset, clips[k].frames, clips[k].srcRate = len(blob), frames[k], float(rates[k])
        blob += raw
    total = int(ctx.lib.aukit_batch_plan(clips, 4, 2, 48000.0))
    # move clip 2's output somewhere of the caller's choosing (rows 100 floats further apart)
    clips[2].out_offset, clips[2].out_stride = total, clips[2].out_stride + 100
    total += int(clips[2].out_stride) * 2
    d_in = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    d_out = torch.full((total,), 7.0, dtype=torch.float32, device="cuda")
    ctx.use_torch_stream()
    try:
        ak._lib.check(ctx.lib.aukit_cuda_dev_batch_resample_amplify(ctx.handle, clips, 4, 24, 0, 2, 1, 48000.0, 2, 0.5, d_in.data_ptr(),
                                                                    d_out.data_ptr()))
        torch.cuda.synchronize()
    finally:
        ctx.set_stream(None)
    out = d_out.cpu().numpy()
    for k in range(4):
        ref = _oracle_chain(O, raws[k], 24, "signed", 2, True, rates[k], 48000, "cubic", 0.5)
        n, st, off = int(clips[k].n_out), int(clips[k].out_stride), int(clips[k].out_offset)
        assert n == ref.shape[1]
        for c in range(2):
            assert np.max(np.abs(out[off + c * st: off + c * st + n] - ref[c])) <= TOL, (k, c)
            assert np.all(out[off + c * st + n: off + (c + 1) * st] == 7.0), "wrote past the end of a row"


def test_batch_falls_back_per_clip_where_the_tile_kernel_does_not_apply(ak, O):
    """Float input (unbounded), 'none' interpolation and non-integer rates run clip by clip through the general kernels."""
    rng = np.random.default_rng(9)
    f = (rng.standard_normal(5000 * 2) * 0.6).astype("<f4").tobytes()
    outs = ak.preload_clips([f, f], [44100, 44056.5], 32, "float", 2, 48000, "cubic", 0.9, False)
    for src, a in zip((44100, 44056.5), outs):
        ref = _oracle_chain(O, f, 32, "float", 2, False, src, 48000, "cubic", 0.9)
        assert a.numpy().shape == ref.shape and np.max(np.abs(a.numpy() - ref)) <= TOL
    s = rng.integers(0, 256, 4000 * 4, dtype=np.uint8).tobytes()
    a = ak.preload_clips([s], 44100, 16, "signed", 2, 48000, "none", 0.5)[0]
    ref = _oracle_chain(O, s, 16, "signed", 2, False, 44100, 48000, "none", 0.5)
    assert f32_equal_bits(a.numpy(), ref.astype(np.float32))
    with pytest.raises(ak.AukitError, match="uneven amount of data per channel"):
        ak.preload_clips([b"\0" * 7], 44100, 24, "signed", 2)
    assert ak.preload_clips([], 44100) == []


def test_s24_conversion_is_exact_for_every_value(ak):
    """conv_sample<S24>: fma(max(lo, 0), RN(1 / (2^23 - 1)), lo) must equal (float)((double)s / 8388607) (A:1133) for all
    2^24 values -- through the batch kernel at ratio 1/2 (every second frame is copied unchanged), multiplier 1."""
    s = np.arange(-(1 << 23), 1 << 23, dtype=np.int64)
    frames = np.zeros((s.size, 2), dtype=np.int64)          # mono frames doubled: outputs pick frames 0, 2, 4, ...
    frames[:, 0] = s
    v = frames.reshape(-1)
    raw = np.stack([(v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF], axis=1).astype(np.uint8).tobytes()
    got = ak.preload_clips([raw], 96000, 24, "signed", 1, 48000, "cubic", 1.0, True)[0].numpy()[0]
    ref = np.where(s < 0, s / 8388608.0, s / 8388607.0).astype(np.float32)
    assert f32_equal_bits(got, ref)
