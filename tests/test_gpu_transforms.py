"""Parity of the float stages (resample / mono / amplify / normalize / fused chain) against the
oracle.  Tolerance: |f32 - ref_double| <= 2^-20 (BASELINE.json:north_star); lengths exact;
+-1 LSB after the reference's 8/16-bit requantisation formula (A:874)."""
import ctypes as C

import numpy as np
import pytest

from util import TOL, f32_equal_bits, tone_s16

pytestmark = pytest.mark.gpu

RATES = [(44100, 48000), (22050, 48000), (96000, 48000), (11025, 48000), (48000, 44100), (8000, 48000),
         (44100, 44100), (48000, 8000), (44056, 48000), (96000, 44100)]


def _signal(n, ch, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (ch, n)).astype(np.float32)          # full-scale noise: cubic overshoots -> clamp path
    x[:, : n // 2] *= np.float32(0.3)
    return x


def _requant_ok(got, ref, bits):
    mx = 2.0 ** (bits - 1)
    q = lambda d: np.where(d < 0, d * mx, d * (mx - 1))          # A:874, un-rounded
    return np.nanmax(np.abs(np.round(q(got.astype(np.float64))) - np.round(q(ref)))) <= 1


@pytest.mark.parametrize("src,dst", RATES)
@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
def test_resample_matches_oracle(ak, O, src, dst, interp):
    x = _signal(20011, 2, src + dst)
    a = ak.Audio.from_numpy(x, src)
    r = a.resample(dst, interp)
    ref = O.resample(x.astype(np.float64), src, dst, interp)
    got = r.numpy()
    assert got.shape == ref.shape and r.sampleRate == dst
    if interp == "none":
        assert f32_equal_bits(got, ref.astype(np.float32))       # selects samples: must be identical (finding 5)
    else:
        assert np.max(np.abs(got - ref)) <= TOL
        assert _requant_ok(got, ref, 16) and _requant_ok(got, ref, 8)


def test_resample_exact_hits_unclamped_and_edges(ak, O):
    x = np.array([[2.0, -3.0, 0.5, 0.25, -0.75]], dtype=np.float32)   # out-of-range float PCM (A:667 vs A:668)
    for interp in ("none", "linear", "cubic"):
        for src, dst in ((1, 2), (1, 3), (2, 1), (3, 7)):
            got = ak.Audio.from_numpy(x, src).resample(dst, interp).numpy()
            ref = O.resample(x.astype(np.float64), src, dst, interp)
            assert got.shape == ref.shape
            assert np.max(np.abs(got - ref)) <= TOL, (interp, src, dst)
    one = ak.Audio.from_numpy(np.array([[0.5]], dtype=np.float32), 8000).resample(48000, "cubic").numpy()
    assert np.max(np.abs(one - O.resample(np.array([[0.5]]), 8000, 48000, "cubic"))) <= TOL
    empty = ak.pcm(b"", 16).resample(48000)
    assert empty.frames == 0
    with pytest.raises(ak.AukitError, match="invalid interpolation type"):
        ak.Audio.from_numpy(x, 1).resample(2, "bogus")


def test_resample_default_interpolation_and_metadata_copy(ak, O):
    x = _signal(1000, 1, 5)
    a = ak.Audio.from_numpy(x, 44100)
    a.metadata["title"] = "t"
    r = a.resample(48000)                                        # aukit.defaultInterpolation = "linear" (A:99, A:655)
    assert np.max(np.abs(r.numpy() - O.resample(x.astype(np.float64), 44100, 48000, "linear"))) <= TOL
    assert r.metadata == {"title": "t"} and r.metadata is not a.metadata


def test_resample_index_parity_none_mode_long(ak, O):
    """floor(x) / exact-hit decisions over 2e6 outputs at 44.1k -> 48k: a ramp input makes the
    selected index directly visible in 'none' mode (SURVEY finding 5: 2172 of 3000 rational
    hits land one sample lower in the reference)."""
    n = 1_900_000
    ramp = (np.arange(n, dtype=np.float64) % 65521) / 65536.0
    x = ramp.astype(np.float32)[None, :]
    got = ak.Audio.from_numpy(x, 44100).resample(48000, "none").numpy()
    ref = O.resample(x.astype(np.float64), 44100, 48000, "none")
    assert f32_equal_bits(got, ref.astype(np.float32))


@pytest.mark.parametrize("ch", [1, 2, 3, 8])
def test_mono(ak, O, ch):
    x = _signal(10007, ch, ch)
    m = ak.Audio.from_numpy(x, 48000).mono()
    ref = O.mono(x.astype(np.float64))
    assert m.channels() == 1 and m.frames == 10007
    assert f32_equal_bits(m.numpy(), ref.astype(np.float32))      # fp64 sum, one rounding


def test_amplify_and_normalize_in_place(ak, O):
    x = _signal(30011, 2, 9) * np.float32(0.6)
    a = ak.Audio.from_numpy(x, 48000)
    assert ak.effects.amplify(a, 1) is a and f32_equal_bits(a.numpy(), x)             # m == 1: untouched (A:3359)
    assert ak.effects.amplify(a, 2.5) is a                                            # mutates and returns the argument
    assert f32_equal_bits(a.numpy(), O.amplify(x.astype(np.float64), 2.5).astype(np.float32))
    for peak, indep in ((1.0, False), (0.8, False), (0.8, True), (None, None)):
        b = ak.Audio.from_numpy(x, 48000)
        args = [] if peak is None else [peak, indep]
        assert ak.effects.normalize(b, *args) is b
        ref = O.normalize(x.astype(np.float64), 1.0 if peak is None else peak, bool(indep))
        got = b.numpy()
        assert np.max(np.abs(got - ref)) <= TOL
        assert np.max(np.abs(got)) == pytest.approx(1.0 if peak is None else peak, abs=1e-6)


def test_normalize_silence_nan_and_nan_input(ak, O):
    z = ak.Audio.from_numpy(np.zeros((1, 100), dtype=np.float32), 48000)
    assert np.isnan(ak.effects.normalize(z, 0.8).numpy()).all()                       # 0 * inf (A:3444, A:3455)
    x = np.array([[np.nan, 0.5, -0.25]], dtype=np.float32)
    got = ak.effects.normalize(ak.Audio.from_numpy(x, 48000)).numpy()
    assert np.isnan(got[0, 0]) and got[0, 1] == 1.0 and got[0, 2] == -0.5             # math.max ignores NaN


@pytest.mark.parametrize("interp", ["linear", "cubic", "none"])
def test_auplay_chain_unfused_vs_oracle(ak, O, interp):
    """BASELINE config 1: 10 s 44.1 kHz stereo s16 WAV -> aukit.wav -> :resample -> :mono -> normalize(0.8)."""
    from util import wav_pcm
    pcm = tone_s16(441000, 2, 44100, seed=1)
    a = ak.wav(wav_pcm(pcm.tobytes(), 2, 44100, 16))
    out = ak.effects.normalize(a.resample(48000, interp).mono(), 0.8)
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, interp, 0.8)
    got = out.numpy()[0]
    assert got.shape == ref.shape == (480000,)
    assert np.max(np.abs(got - ref)) <= TOL
    assert _requant_ok(got, ref, 16)


@pytest.mark.parametrize("interp", ["linear", "cubic", "none"])
@pytest.mark.parametrize("kind", ["tone", "noise"])
def test_fused_pipeline_vs_oracle(ak, O, interp, kind):
    n = 300007
    pcm = tone_s16(n, 2, 44100, seed=2) if kind == "tone" else \
        np.random.default_rng(2).integers(-32768, 32768, (n, 2)).astype(np.int16)
    got = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, interp, True, 0.8)[0]
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, interp, 0.8)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL
    assert _requant_ok(got, ref, 16) and _requant_ok(got, ref, 8)


@pytest.mark.parametrize("bits,dtype,be,ch,src", [(24, "signed", True, 2, 22050), (8, "unsigned", False, 1, 11025),
                                                  (32, "float", False, 3, 96000), (16, "signed", True, 2, 48000),
                                                  (32, "signed", False, 2, 44100)])
def test_fused_pipeline_other_formats(ak, O, bits, dtype, be, ch, src):
    rng = np.random.default_rng(bits + ch)
    n = 50021
    raw = rng.integers(0, 256, n * ch * bits // 8, dtype=np.uint8)
    if dtype == "float":
        raw = (rng.standard_normal(n * ch) * 0.4).astype("<f4").view(np.uint8)
    for mono in (True, False):
        got = ak.preload(raw, bits, dtype, ch, src, 48000, "cubic", mono, 1.0, be)
        x = O.resample(O.pcm(raw, bits, dtype, ch, True, be), src, 48000, "cubic")
        ref = O.normalize(O.mono(x) if mono else x, 1.0)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= TOL


def test_time_sharded_resample_is_bitwise_identical(ak, O):
    """Sharding the OUTPUT range (each shard carrying its halo) reproduces the single-pass result
    bit for bit, because positions come from the global index."""
    lib = ak._lib.load()
    ctx = ak.context()
    x = _signal(100003, 2, 21)
    for src, dst, interp in ((96000, 44100, "cubic"), (44100, 48000, "linear"), (96000, 48000, "cubic"), (44100, 48000, "none")):
        mode = {"none": 0, "linear": 1, "cubic": 2}[interp]
        whole = ak.Audio.from_numpy(x, src).resample(dst, interp).numpy()
        n_out = whole.shape[1]
        parts = []
        for r in range(4):
            o0, o1 = n_out * r // 4, n_out * (r + 1) // 4
            f, c = C.c_uint64(), C.c_uint64()
            assert lib.aukit_resample_window(x.shape[1], src, dst, mode, o0, o1 - o0, C.byref(f), C.byref(c)) == 0
            shard_in = ak.Audio.from_numpy(x[:, f.value: f.value + c.value], src)      # shard + halo only
            out = ak.Audio.from_numpy(np.zeros((2, o1 - o0), dtype=np.float32), dst)
            ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, shard_in.data_ptr, shard_in.stride, 2, x.shape[1], f.value,
                                                      c.value, float(src), float(dst), mode, o0, o1 - o0, out.data_ptr, out.stride))
            parts.append(out.numpy())
        assert f32_equal_bits(np.concatenate(parts, axis=1), whole)


def test_time_sharded_fused_pipeline_matches_single_pass(ak, O):
    import torch
    lib = ak._lib.load()
    ctx = ak.context()
    n = 200003
    pcm = np.random.default_rng(5).integers(-20000, 20000, (n, 2)).astype(np.int16)
    whole = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)
    n_out = whole.shape[1]
    ctx.use_torch_stream()
    dmax = torch.zeros(1, device="cuda")
    descs, ins = [], []
    for r in range(3):
        o0, o1 = n_out * r // 3, n_out * (r + 1) // 3
        f, c = C.c_uint64(), C.c_uint64()
        lib.aukit_resample_window(n, 44100.0, 48000.0, 2, o0, o1 - o0, C.byref(f), C.byref(c))
        t = torch.from_numpy(pcm[f.value: f.value + c.value].copy()).cuda()
        d = ak.PipelineDesc(16, 0, 2, 0, 44100.0, 48000.0, 2, 1, n, f.value, c.value, o0, o1 - o0)
        descs.append(d); ins.append(t)
    torch.cuda.synchronize()
    for d, t in zip(descs, ins):      # local peaks max-combine into one value (the allreduce-max on one GPU)
        ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(d), t.data_ptr(), dmax.data_ptr()))
    ctx.synchronize()
    outs = []
    for d, t in zip(descs, ins):
        o = torch.empty(d.n_out, device="cuda")
        ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d), t.data_ptr(), 0.8, dmax.data_ptr(), o.data_ptr(), d.n_out))
        ctx.synchronize()
        outs.append(o.cpu().numpy())
    ctx.set_stream(None)
    assert f32_equal_bits(np.concatenate(outs)[None, :], whole)


def test_concat_joins_blocks(ak):
    a = ak.Audio.from_numpy(np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float32) / 8, 8000)
    b = ak.Audio.from_numpy(np.array([[7, 8]], dtype=np.float32) / 8, 8000)
    c = a.concat(b)
    assert (c.numpy() * 8).tolist() == [[1, 2, 3, 7, 8], [4, 5, 6, 0, 0]]       # missing channel -> silence (A:713)
    assert a.len() == 3 / 8000 and c.channels() == 2


@pytest.mark.parametrize("src", [44100, 22050, 96000, 48000, 32000, 8000, 88200])
@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
def test_fused_polyphase_rates_and_exact_hits(ak, O, src, interp):
    """The rational-ratio fast path against the oracle, with float input that exceeds [-1, 1] so the
    exact-hit (copied unclamped, A:667) vs near-hit (clamped, A:668; `none` picks the previous
    sample) decisions are visible in the output."""
    rng = np.random.default_rng(src)
    n = 40013
    x = (rng.standard_normal((n, 2)) * 0.9).astype("<f4")
    got = ak.preload(x.tobytes(), 32, "float", 2, src, 48000, interp, False, 1.0)
    r = O.resample(O.pcm(x, 32, "float", 2), src, 48000, interp)
    ref = O.normalize(r, 1.0)
    assert got.shape == ref.shape
    if interp == "none":
        # same selected samples, same clamp decisions: only the final scale is rounded in f32
        assert np.max(np.abs(got - ref)) <= 2.0 ** -22
    else:
        assert np.max(np.abs(got - ref)) <= TOL * 4      # |values| reach ~4: tolerance scales with magnitude
    s16 = rng.integers(-32768, 32768, (n, 2)).astype("<i2")
    got = ak.preload(s16.tobytes(), 16, "signed", 2, src, 48000, interp, True, 0.8)[0]
    ref = O.chain_s16(s16.tobytes(), 2, src, 48000, interp, 0.8)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= TOL


def test_fused_polyphase_mono_input_and_8ch(ak, O):
    rng = np.random.default_rng(77)
    for ch, src in ((1, 44100), (8, 96000), (8, 44100), (4, 22050)):
        x = rng.integers(-32768, 32768, (30011, ch)).astype("<i2")
        for mono in (True, False):
            got = ak.preload(x.tobytes(), 16, "signed", ch, src, 48000, "cubic", mono, 0.9)
            r = O.resample(O.pcm(x, 16, "signed", ch), src, 48000, "cubic")
            ref = O.normalize(O.mono(r) if mono else r, 0.9)
            assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= TOL


@pytest.mark.parametrize("src,dst", [(44100, 48000), (96000, 44100), (32000, 48000)])
@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
@pytest.mark.parametrize("n_total", [6_000_000_011, 1_400_000_003])
def test_huge_global_positions(ak, src, dst, interp, n_total):
    """Shards near the END of a buffer of several 10^9 frames: positions are >= 2^28, where the
    reference's fp64 rounding of x shows up in the interpolated value (x's ulp is ~2^-21).  Covers
    the fused path's exact-position mode and the standalone resample kernel."""
    import torch
    from util import ref_resample_window
    lib = ak._lib.load()
    ctx = ak.context()
    ctx.use_torch_stream()
    mode = {"none": 0, "linear": 1, "cubic": 2}[interp]
    total_out = int(lib.aukit_resample_out_len(n_total, float(src), float(dst)))
    rng = np.random.default_rng(src + mode)
    starts = [total_out - 150_000, int(total_out * 0.37)]
    for pos in (2 ** 28 / 1.25, 2 ** 28, 2 ** 29, 2 ** 30, 2 ** 30.5):          # shards straddling every mode / drift-segment boundary
        o = int(pos * dst / src) - 70_000
        if 0 < o < total_out - 150_000:
            starts.append(o)
    for o0 in starts:
        cnt = 150_000 if o0 + 150_000 <= total_out else total_out - o0
        f, c = C.c_uint64(), C.c_uint64()
        assert lib.aukit_resample_window(n_total, float(src), float(dst), mode, o0, cnt, C.byref(f), C.byref(c)) == 0
        pcm = rng.integers(-32768, 32768, (int(c.value), 2)).astype(np.int16)
        dec = np.where(pcm < 0, pcm / 32768.0, pcm / 32767.0).T                    # A:1133
        ref = ref_resample_window(dec, int(f.value), n_total, src, dst, o0, cnt, interp)
        # fused pipeline (mono, normalize 0.8)
        mono = (ref[0] + ref[1]) / 2
        want = np.clip(mono * (0.8 / np.max(np.abs(mono))), -1, 1)
        t = torch.from_numpy(pcm).cuda()
        dmax = torch.zeros(1, device="cuda")
        out = torch.empty(cnt, device="cuda")
        d = ak.PipelineDesc(16, 0, 2, 0, float(src), float(dst), mode, 1, n_total, f.value, c.value, o0, cnt)
        ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(d), t.data_ptr(), dmax.data_ptr()))
        ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d), t.data_ptr(), 0.8, dmax.data_ptr(), out.data_ptr(), cnt))
        got = out.cpu().numpy()
        assert np.max(np.abs(got - want)) <= TOL, (o0, float(np.max(np.abs(got - want))))
        # standalone resample kernel on the decoded floats
        fin = torch.from_numpy(dec.astype(np.float32)).cuda().contiguous()
        fout = torch.empty((2, cnt), device="cuda")
        ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, fin.data_ptr(), fin.shape[1], 2, n_total, f.value, c.value,
                                                  float(src), float(dst), mode, o0, cnt, fout.data_ptr(), cnt))
        got2 = fout.cpu().numpy()
        if interp == "none":
            assert f32_equal_bits(got2, ref.astype(np.float32))
        else:
            assert np.max(np.abs(got2 - ref)) <= TOL
    ctx.set_stream(None)


@pytest.mark.parametrize("src", [44100, 22050, 11025])
def test_run_per_lane_kernel_matches_oracle_and_polyphase(ak, O, src, monkeypatch):
    """Large enough for the run-per-lane kernel (interior warp tiles) with the polyphase kernels on the
    head and tail: against the oracle within tolerance, and against the polyphase-only result."""
    n = 1_300_003 * src // 44100
    pcm = np.random.default_rng(src).integers(-32768, 32768, (n, 2)).astype(np.int16)
    got = ak.preload(pcm.tobytes(), 16, "signed", 2, src, 48000, "cubic", True, 0.8)[0]
    ref = O.chain_s16(pcm.tobytes(), 2, src, 48000, "cubic", 0.8)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL


def test_polyphase_only_paths_in_subprocess(O):
    """The kernel-selection switches are read once per process: exercise the polyphase-only apply pass
    (AUKIT_RUN_APPLY=0) and the polyphase-only pipeline (AUKIT_DISABLE_RUN=1) in child processes."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import aukit_b200 as ak
from oracle import oracle as O
pcm = np.random.default_rng(9).integers(-32768, 32768, (900011, 2)).astype(np.int16)
got = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)[0]
ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8)
assert got.shape == ref.shape
err = float(np.max(np.abs(got - ref)))
assert err <= 2.0 ** -20, err
print("ok", err)
''' % root
    for extra in ({"AUKIT_RUN_APPLY": "0"}, {"AUKIT_DISABLE_RUN": "1"}, {"AUKIT_DISABLE_POLY": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **extra), timeout=600)
        assert r.returncode == 0 and "ok" in r.stdout, (extra, r.stdout + r.stderr)


@pytest.mark.parametrize("n,ch,freq,rate", [(1_000_003, 2, 24000.0, 48000), (300_001, 3, 200.0, 44100), (2_000_000, 1, 5.0, 48000),
                                            (4096, 1, 1000.0, 48000), (4097, 2, 1000.0, 48000), (8192, 1, 1000.0, 48000), (8193, 2, 150.0, 48000), (17, 1, 3000.0, 8000),
                                            (700_000, 1, 0.05, 48000)])
def test_lowpass_matches_oracle(ak, O, n, ch, freq, rate):
    """effects.lowpass (A:3586): chained-tile scan vs the sequential reference recurrence, incl. cut-offs
    whose memory spans hundreds of 4096-sample tiles (5 Hz, 0.05 Hz: the look-back cannot stop early)."""
    rng = np.random.default_rng(n)
    x = (rng.uniform(-1, 1, (ch, n)) + 0.3).astype(np.float32)
    a = ak.Audio.from_numpy(x, rate)
    assert ak.effects.lowpass(a, freq) is a
    got = a.numpy()
    ref = O.lowpass(x.astype(np.float64), freq, rate)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 0], x[:, 0])                              # d[1] is untouched (A:3591)
    assert float(np.max(np.abs(got - ref))) <= TOL


def test_lowpass_after_normalize_like_auplay(ak, O):
    """auplay.lua:27-30: normalize(0.8) then lowpass(sampleRate / 2) on the mono mixdown."""
    pcm = tone_s16(50_000, 2, seed=4)
    a = ak.pcm(pcm.tobytes(), 16, "signed", 2, 44100).resample(48000, "cubic").mono()
    ak.effects.normalize(a, 0.8)
    ak.effects.lowpass(a, a.sampleRate / 2)
    dec = O.pcm(pcm.tobytes(), 16, "signed", 2, True, False)
    ref = O.lowpass(O.normalize(O.mono(O.resample(dec, 44100, 48000, "cubic")), 0.8, False), 24000.0, 48000.0)
    assert float(np.max(np.abs(a.numpy() - ref))) <= 2 * TOL


@pytest.mark.parametrize("bits,dtype", [(8, "signed"), (8, "unsigned"), (16, "signed"), (24, "signed"), (24, "unsigned"),
                                        (32, "signed"), (32, "unsigned"), (32, "float")])
@pytest.mark.parametrize("ch,interleaved", [(1, True), (2, True), (2, False), (3, True)])
def test_audio_pcm_values_bit_exact_and_bytes(ak, O, bits, dtype, ch, interleaved):
    """Audio:pcm (A:901) / encodePCM (A:868): the un-rounded values are bit-identical to the oracle's doubles
    (f32 sample x 32-bit scale is exact in fp64); the packed bytes follow the stated rounding mode."""
    n = 100_003
    rng = np.random.default_rng(bits * 10 + ch)
    x = rng.uniform(-1, 1, (ch, n)).astype(np.float32)
    x[:, :5] = np.array([-1.0, 1.0, 0.0, -0.0, 0.5], dtype=np.float32)
    if bits <= 24:
        # adversarial: products that land on or next to integers and exact ties (k, k + 1/2, +- a few ulps)
        smax = float(2 ** (bits - 1))
        k = rng.integers(-2 ** (bits - 1), 2 ** (bits - 1), 4000).astype(np.float64)
        for j, (off, sc) in enumerate(((0.0, smax), (0.5, smax), (0.0, smax - 1), (0.5, smax - 1))):
            base = ((k[j * 1000:(j + 1) * 1000] + off) / sc).astype(np.float32)
            for u, tweak in enumerate((0, 1, -1, 2)):
                seg = np.nextafter(base, np.float32(np.inf if tweak > 0 else -np.inf)) if tweak else base
                if abs(tweak) == 2:
                    seg = np.nextafter(seg, np.float32(np.inf))
                x[0, 100 + (j * 4 + u) * 1000: 100 + (j * 4 + u + 1) * 1000] = np.clip(seg, -1, 1)
    a = ak.Audio.from_numpy(x, 48000)
    got = a.pcm(bits, dtype, interleaved)
    ref = O.audio_pcm(x.astype(np.float64), bits, dtype, interleaved)
    assert got.dtype == np.float64 and np.array_equal(got, ref) and np.array_equal(np.signbit(got), np.signbit(ref))
    B = bits // 8
    for mode, fn in (("truncate", np.trunc), ("floor", np.floor), ("nearest", np.rint)):
        raw = np.frombuffer(a.pcm_bytes(bits, dtype, interleaved, mode), dtype=np.uint8).reshape(-1, B)
        if dtype == "float":
            order = x.T.reshape(-1) if interleaved else x.reshape(-1)
            assert np.array_equal(raw.reshape(-1).view("<f4"), order)
            continue
        q = fn(ref).astype(np.int64)
        lo, hi = (0, 2 ** bits - 1) if dtype == "unsigned" else (-2 ** (bits - 1), 2 ** (bits - 1) - 1)
        q = np.clip(q, lo, hi)
        want = np.zeros((q.size, B), dtype=np.uint8)
        for k in range(B):
            want[:, k] = (q >> (8 * k)) & 0xFF
        assert np.array_equal(raw, want), mode


@pytest.mark.parametrize("n", [16, 4096, 100_000, 100_003, 7])
@pytest.mark.parametrize("bits,dtype", [(8, "signed"), (8, "unsigned"), (16, "signed"), (16, "unsigned")])
def test_audio_pcm_bytes_wide_kernel_shapes(ak, O, bits, dtype, n):
    """The 16-bytes-per-thread requantisation kernel (mono, planar rows, interleaved stereo) at lengths that are and are
    not multiples of its group size: same bytes as the sample-by-sample definition."""
    rng = np.random.default_rng(n + bits)
    for ch, interleaved in ((1, True), (2, True), (2, False), (4, False)):
        x = rng.uniform(-1.2, 1.2, (ch, n)).astype(np.float32)
        ref = O.audio_pcm(x.astype(np.float64), bits, dtype, interleaved)
        lo, hi = (0, 2 ** bits - 1) if dtype == "unsigned" else (-2 ** (bits - 1), 2 ** (bits - 1) - 1)
        q = np.clip(np.floor(ref).astype(np.int64), lo, hi)
        want = q.astype("<u%d" % (bits // 8) if dtype == "unsigned" else "<i%d" % (bits // 8)).tobytes()
        assert ak.Audio.from_numpy(x, 48000).pcm_bytes(bits, dtype, interleaved, "floor") == want, (ch, interleaved)


def test_audio_pcm_requantisation_within_one_lsb_of_reference_chain(ak, O):
    """north_star's statement for the whole path: after requantisation to 8-bit signed (what speaker.playAudio
    takes, auplay.lua:34) the CUDA chain is within 1 LSB of the reference chain."""
    pcm = tone_s16(60_000, 2, seed=6)
    a = ak.pcm(pcm.tobytes(), 16, "signed", 2, 44100).resample(48000, "cubic").mono()
    ak.effects.normalize(a, 0.8)
    got = np.floor(a.pcm(8, "signed", True))
    ref = np.floor(O.audio_pcm(O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8), 8, "signed", True))
    assert np.max(np.abs(got - ref)) <= 1
    assert np.mean(got != ref) < 1e-3


@pytest.mark.parametrize("peak", [1.0, 1.5, 0.25])
def test_run_per_lane_kernel_other_peaks(ak, O, peak):
    """peakAmplitude >= 1 keeps the final clamp of A:3455 (CLAMP1 variant; 1.5 makes it act), < 1 drops it."""
    n = 1_200_011
    pcm = np.random.default_rng(int(peak * 100)).integers(-32768, 32768, (n, 2)).astype(np.int16)
    got = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, peak)[0]
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", peak)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= TOL * max(1.0, peak)
    if peak > 1:
        assert np.max(got) == 1.0 and np.min(got) == -1.0


def test_fused_chain_on_silence_is_all_nan_like_the_reference(ak, O):
    """max == 0: the reference multiplies every sample by peak / 0 = inf, 0 * inf = NaN, and its clamp lets NaN
    through (A:228, A:3444-3455) -- also on the interior tiles that the run-per-lane kernel takes."""
    n = 1_000_000
    got = ak.preload(bytes(4 * n), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)[0]
    ref = O.chain_s16(bytes(4 * 4000), 2, 44100, 48000, "cubic", 0.8)
    assert np.all(np.isnan(ref)) and np.all(np.isnan(got))


@pytest.mark.parametrize("src,dst,ch", [(44100, 48000, 2), (48000, 8000, 1), (11025, 48000, 3), (44100, 44100, 1)])
def test_resample_sinc_matches_oracle_and_shards(ak, O, src, dst, ch):
    """interpolate.sinc (A:267-281) through Audio:resample, whole and as two output shards that only hold
    their +-10 frame window (aukit_resample_window)."""
    lib = ak._lib.load()
    ctx = ak.context()
    n = 20_011
    x = np.random.default_rng(src + dst).uniform(-1, 1, (ch, n)).astype(np.float32)
    got = ak.Audio.from_numpy(x, src).resample(dst, "sinc").numpy()
    ref = O.resample(x.astype(np.float64), src, dst, "sinc")
    assert got.shape == ref.shape and float(np.max(np.abs(got - ref))) <= 2 * TOL
    n_out = got.shape[1]
    parts = []
    for o0, o1 in ((0, n_out // 2), (n_out // 2, n_out)):
        f, c = C.c_uint64(), C.c_uint64()
        assert lib.aukit_resample_window(n, float(src), float(dst), 3, o0, o1 - o0, C.byref(f), C.byref(c)) == 0
        shard_in = ak.Audio.from_numpy(np.ascontiguousarray(x[:, f.value: f.value + c.value]), src)
        out = ak.Audio.from_numpy(np.zeros((ch, o1 - o0), dtype=np.float32), dst)
        ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, shard_in.data_ptr, shard_in.stride, ch, n, f.value, c.value,
                                                  float(src), float(dst), 3, o0, o1 - o0, out.data_ptr, out.stride))
        parts.append(out.numpy())
    assert f32_equal_bits(np.concatenate(parts, axis=1), got)
    with pytest.raises(ak.AukitError, match="Audio:resample only"):
        ak.preload(bytes(4000), 16, "signed", 2, 44100, 48000, "sinc", True, 0.8)


def test_more_effects_match_oracle_at_size(ak, O):
    """effects.invert / fade / delay / center (A:3392-3513) on buffers that span many CTAs and centre blocks."""
    rate, n = 8000, 1_000_003
    x = (np.random.default_rng(77).uniform(-1, 1, (2, n)) * 0.9 + 0.05).astype(np.float32)
    xd = x.astype(np.float64)
    a = ak.Audio.from_numpy(x, rate)
    assert ak.effects.invert(a) is a and np.array_equal(a.numpy(), -x)
    a = ak.Audio.from_numpy(x, rate)
    ak.effects.fade(a, 1.0, 0.25, 100.0, 1.75)
    assert float(np.max(np.abs(a.numpy() - O.fade(xd, rate, 1.0, 0.25, 100.0, 1.75)))) <= TOL
    a = ak.Audio.from_numpy(x, rate)
    ak.effects.delay(a, 0.37, 0.8)
    assert float(np.max(np.abs(a.numpy() - O.delay(xd, rate, 0.37, 0.8)))) <= TOL
    a = ak.Audio.from_numpy(x, rate)
    ak.effects.center(a)
    assert float(np.max(np.abs(a.numpy() - O.center(xd, rate)))) <= TOL
    with pytest.raises(ak.AukitError, match="arithmetic on a nil value"):
        ak.effects.fade(ak.Audio.from_numpy(x, rate), 0.00001, 0.5, 1.0, 1.0)
    with pytest.raises(ak.AukitError, match="arithmetic on a nil value"):
        ak.effects.delay(ak.Audio.from_numpy(x, rate), -1.0)


@pytest.mark.parametrize("n,ch,freq,rate", [(1_000_003, 2, 200.0, 48000), (300_001, 3, 8000.0, 44100), (2_000_000, 1, 0.5, 48000),
                                            (4097, 2, 1000.0, 48000), (5, 1, 3000.0, 8000)])
def test_highpass_matches_oracle(ak, O, n, ch, freq, rate):
    """effects.highpass (A:3605): the lowpass scan with ratio a and input a (x[i] - x[i-1]); 0.5 Hz keeps the
    state alive across hundreds of tiles, and the tile boundaries need the saved original x[i-1]."""
    x = (np.random.default_rng(n + 1).uniform(-1, 1, (ch, n)) + 0.3).astype(np.float32)
    a = ak.Audio.from_numpy(x, rate)
    assert ak.effects.highpass(a, freq) is a
    got = a.numpy()
    ref = O.highpass(x.astype(np.float64), freq, rate)
    assert np.array_equal(got[:, 0], x[:, 0])
    assert float(np.max(np.abs(got - ref))) <= 2 * TOL


@pytest.mark.parametrize("kind,n,ch,freq,rate", [("low", 10_200_003, 3, 200.0, 44100), ("low", 30_000_001, 1, 24000.0, 48000),
                                                  ("high", 10_200_003, 3, 200.0, 48000), ("high", 15_000_000, 2, 8000.0, 44100),
                                                  ("low", 30_000_001, 1, 30.0, 48000), ("high", 15_000_000, 2, 30.0, 48000)])
def test_lowpass_highpass_blocked_chunks_at_size(ak, O, kind, n, ch, freq, rate):
    """Buffers of thousands of tiles with an ordinary cut-off take the blocked variant (csrc/lowpass.cu: one contiguous
    chunk of tiles per CTA, state carried in a register, chunk-start states from the pre-pass): chunks that straddle
    channel boundaries (1246 / 1832 tiles per channel in chunks of 9), a ragged last tile, both effects, and 30 Hz cut-offs
    whose memory is two tiles (two-tile warm-up in the pre-pass) -- against the sequential reference recurrence."""
    x = (np.random.default_rng(n + ch).uniform(-1, 1, (ch, n)) + 0.3).astype(np.float32)
    a = ak.Audio.from_numpy(x, rate)
    fx = ak.effects.lowpass if kind == "low" else ak.effects.highpass
    assert fx(a, freq) is a
    got = a.numpy()
    ref = (O.lowpass if kind == "low" else O.highpass)(x.astype(np.float64), freq, rate)
    assert got.shape == ref.shape
    assert np.array_equal(got[:, 0], x[:, 0])
    assert float(np.max(np.abs(got - ref))) <= (TOL if kind == "low" else 2 * TOL)


@pytest.mark.parametrize("n,at", [(300_001, 101_234), (300_001, 8192 * 5 - 1), (10_100_003, 3_367_901), (10_100_003, 8192 * 700 - 1)])
def test_lowpass_highpass_non_finite_input_poisons_the_rest_of_the_channel(ak, O, n, at):
    """A:3592-3595 / A:3613-3615: once the state is non-finite it stays so -- every later sample of THAT channel, however
    many tiles (or chunks) away; the other channels are untouched.  Small buffers take the look-back, large ones the
    blocked chunks; both rely on lp_poison_fix beyond their one-tile / 2^-80 memory.  An infinite input as the LAST sample
    of a tile leaves an infinite state behind that must not be taken for a lasting one."""
    rng = np.random.default_rng(n)
    x = rng.uniform(-0.9, 0.9, (3, n)).astype(np.float32)
    x[1, at] = np.nan                                   # mid-tile, or the last sample of a tile
    a = ak.Audio.from_numpy(x, 48000)
    ak.effects.lowpass(a, 2000.0)
    got = a.numpy()
    ref = O.lowpass(x.astype(np.float64), 2000.0, 48000)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.isnan(got[1, at:]).all() and np.isfinite(got[1, :at]).all() and np.isfinite(got[[0, 2]]).all()
    fin = np.isfinite(ref)
    assert float(np.max(np.abs(got[fin] - ref[fin]))) <= TOL
    x[1, at] = np.inf                                   # highpass: +Inf once, then (Inf + x) - Inf = NaN for good
    a = ak.Audio.from_numpy(x, 48000)
    ak.effects.highpass(a, 300.0)
    got = a.numpy()
    ref = O.highpass(x.astype(np.float64), 300.0, 48000)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(np.isinf(got), np.isinf(ref))
    fin = np.isfinite(ref)
    assert float(np.max(np.abs(got[fin] - ref[fin]))) <= 2 * TOL


def _quiet_with_bursts(n, bursts, seed=77, base=12000):
    """Moderate-level stereo noise (no cubic overshoot can reach +-1) with full-scale noise in `bursts`."""
    rng = np.random.default_rng(seed)
    pcm = rng.integers(-base, base + 1, (n, 2)).astype(np.int16)
    for lo, hi in bursts:
        pcm[lo:hi] = rng.integers(-32768, 32768, (hi - lo, 2)).astype(np.int16)
    return pcm


@pytest.mark.parametrize("bursts", [(), ((400_000, 400_700), (1_100_000, 1_100_040)), ((0, 50), (1_499_000, 1_500_007))])
def test_static_run_kernel_checked_clamp_and_redo(ak, O, bursts, tmp_path):
    """44.1 -> 48 kHz takes the straight-line kernel, which checks the per-channel clamp of A:668 instead of applying
    it and redoes a tile with the clamping twin when it acted: a signal where it never acts, one where it acts in a few
    interior tiles, one where it acts only in the polyphase edges.  Within tolerance of the oracle, and bit-identical
    to the scripted kernel and to the polyphase-only pipeline (what keeps time shards equal to a single pass)."""
    import os
    import subprocess
    import sys
    n = 1_500_007
    pcm = _quiet_with_bursts(n, bursts)
    got = ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)[0]
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    inp, outp = str(tmp_path / "in.npy"), str(tmp_path / "out.npy")
    np.save(inp, pcm)
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import aukit_b200 as ak
pcm = np.load(%r)
np.save(%r, ak.preload(pcm.tobytes(), 16, "signed", 2, 44100, 48000, "cubic", True, 0.8)[0])
''' % (root, inp, outp)
    for extra in ({"AUKIT_RUN_STATIC": "0"}, {"AUKIT_DISABLE_RUN": "1"}):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, **extra), timeout=600)
        assert r.returncode == 0, (extra, r.stdout + r.stderr)
        assert f32_equal_bits(np.load(outp), got), extra


def test_far_shards_do_not_depend_on_the_sharding(ak):
    """Positions of several 10^8 frames (hours 3-8 of a time-sharded buffer): the drift baked into the run-per-lane
    kernel's weight table comes from a fixed GLOBAL grid of segments, so a range computed as one shard, as two shards,
    or as part of a longer shard that starts elsewhere gives the same bits."""
    import torch
    from aukit_b200.sharding import padded_window
    lib, ctx = ak._lib.load(), ak.context()
    n_total, src, dst, TILE = 8 * 3600 * 44100, 44100, 48000, 5120
    rng = np.random.default_rng(11)
    ctx.use_torch_stream()
    try:
        for pos in (3.3e8, 2 ** 28 / 1.25 * 1.25 ** 3, 9.1e8):          # inside a segment, across a segment boundary, far out
            o_mid = int(pos * dst / src) // TILE * TILE
            o0, o1 = o_mid - 40 * TILE, o_mid + 40 * TILE
            f_all, c_all = padded_window(n_total, src, dst, "cubic", o0 - 30 * TILE, (o1 - o0) + 30 * TILE)
            pcm = rng.integers(-32768, 32768, (c_all, 2)).astype(np.int16)
            dmax = torch.full((1,), 0.9, device="cuda")

            def run(a0, a1):
                f, c = padded_window(n_total, src, dst, "cubic", a0, a1 - a0)
                t = torch.from_numpy(pcm[f - f_all: f - f_all + c].copy()).cuda()
                out = torch.empty(a1 - a0, device="cuda")
                d = ak.PipelineDesc(16, 0, 2, 0, float(src), float(dst), 2, 1, n_total, f, c, a0, a1 - a0)
                ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d), t.data_ptr(), 0.8, dmax.data_ptr(), out.data_ptr(), a1 - a0))
                torch.cuda.synchronize()
                return out.cpu().numpy()
            whole = run(o0, o1)
            two = np.concatenate([run(o0, o_mid + 7 * TILE), run(o_mid + 7 * TILE, o1)])
            longer = run(o0 - 30 * TILE, o1)[30 * TILE:]
            assert f32_equal_bits(two, whole) and f32_equal_bits(longer, whole), pos
    finally:
        ctx.set_stream(None)


@pytest.mark.parametrize("src,dst", [(44100, 48000), (48000, 44100), (22050, 48000), (8000, 48000), (32000, 48000), (44100, 22050 * 3)])
@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
def test_planar_tma_resample_equals_polyphase_kernel(ak, O, monkeypatch, src, dst, interp):
    """Audio:resample's interior tiles (resample_planar.cu: TMA-staged, double-buffered) against the polyphase kernel
    that still does the first / last tiles, bit for bit, on data that leaves [-1, 1] and holds NaN / Inf (the exact-hit
    copy of A:667 is unclamped, the clamp of A:668 lets NaN through), and against the oracle."""
    rng = np.random.default_rng(src + dst)
    for ch, n in ((1, 150_001), (2, 260_003), (3, 90_000)):
        x = rng.uniform(-1.6, 1.6, (ch, n)).astype(np.float32)
        x[0, 5000] = np.nan
        if ch != 2:                 # (the reference's polynomial form A:265 turns an infinite tap into NaN, the weight form
            x[ch - 1, 70_000] = np.inf      # into +-1 after the clamp: infinities are checked between the kernels only)
            x[0, 12_345] = -np.inf
        a = ak.Audio.from_numpy(x, src)
        got = a.resample(dst, interp).numpy()
        monkeypatch.setenv("AUKIT_DISABLE_PLANAR", "1")
        other = a.resample(dst, interp).numpy()
        monkeypatch.delenv("AUKIT_DISABLE_PLANAR")
        assert got.shape == other.shape and f32_equal_bits(got, other), (ch, n)
        if ch == 2:
            ref = O.resample(x.astype(np.float64), src, dst, interp)
            # (an output one ulp beside an input frame takes that frame: a non-finite NEIGHBOUR does not reach it as it
            # does through the reference's ~1e-12 weights -- DESIGN.md "known deviations"; everything else must agree)
            fin = np.isfinite(ref) & np.isfinite(got)
            assert np.count_nonzero(np.isnan(ref) != np.isnan(got)) <= 12
            assert np.max(np.abs(got[fin] - ref[fin])) <= 2 * TOL          # |values| reach 1.6 before the clamp


@pytest.mark.parametrize("src,dst,ch", [(44100, 48000, 2), (48000, 44100, 1), (22050, 44100, 3), (8000, 48000, 2)])
def test_planar_sinc_kernel_at_size(ak, O, monkeypatch, src, dst, ch):
    """interpolate.sinc through the polyphase TMA kernel over many tiles: against the oracle, against the one-thread-per-
    frame kernel (same table-plus-correction weights: a few ulps apart), and sharded at an arbitrary output index with
    the shards holding only their +-10 frame windows -- bit-identical to the unsharded call."""
    lib, ctx = ak._lib.load(), ak.context()
    n = 180_017
    x = np.random.default_rng(src + dst + ch).uniform(-1, 1, (ch, n)).astype(np.float32)
    a = ak.Audio.from_numpy(x, src)
    got = a.resample(dst, "sinc").numpy()
    ref = O.resample(x.astype(np.float64), src, dst, "sinc")
    assert got.shape == ref.shape and float(np.max(np.abs(got - ref))) <= 2 * TOL
    monkeypatch.setenv("AUKIT_DISABLE_PLANAR", "1")
    other = a.resample(dst, "sinc").numpy()
    monkeypatch.delenv("AUKIT_DISABLE_PLANAR")
    assert float(np.max(np.abs(got - other))) <= 2.0 ** -21
    n_out = got.shape[1]
    cut = n_out // 3 + 7
    parts = []
    for o0, o1 in ((0, cut), (cut, n_out)):
        f, c = C.c_uint64(), C.c_uint64()
        assert lib.aukit_resample_window(n, float(src), float(dst), 3, o0, o1 - o0, C.byref(f), C.byref(c)) == 0
        shard_in = ak.Audio.from_numpy(np.ascontiguousarray(x[:, f.value: f.value + c.value]), src)
        out = ak.Audio.from_numpy(np.zeros((ch, o1 - o0), dtype=np.float32), dst)
        ak._lib.check(lib.aukit_cuda_dev_resample(ctx.handle, shard_in.data_ptr, shard_in.stride, ch, n, f.value, c.value,
                                                  float(src), float(dst), 3, o0, o1 - o0, out.data_ptr, out.stride))
        parts.append(out.numpy())
    assert f32_equal_bits(np.concatenate(parts, axis=1), got)
