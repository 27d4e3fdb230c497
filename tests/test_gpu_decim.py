"""K15 (csrc/pipeline_decim.cu): ratios 1 / 2^k, where the reference's position is an exact integer for every output
and the sample is copied unclamped (A:666-667) -- BASELINE config 5's 96 -> 48 kHz.  Fused chain and Audio:resample,
against the oracle, the reference's own chain5 vector, and shard-vs-whole bit equality."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from util import TOL, f32_equal_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config5_reference_vector_through_the_fused_chain(ak):
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    manifest = json.loads(z["manifest"].tobytes().decode())
    for name in ("chain5_f32x8_96000_48000", "chain5_f32x8_96000_44100"):
        i = [m["name"] for m in manifest].index(name)
        m = manifest[i]
        got = ak.preload(z["c%d/in" % i].tobytes(), 32, "float", 8, 96000, m["args"]["targetRate"], "cubic", False, 1.0)
        for c in range(8):
            ref = z["c%d/out%d" % (i, c)]
            assert got[c].shape == ref.shape and np.max(np.abs(got[c] - ref)) <= TOL, (name, c)


@pytest.mark.parametrize("ch", [8, 4, 2, 3])
@pytest.mark.parametrize("src", [96000, 192000, 48000])
@pytest.mark.parametrize("interp", ["cubic", "none"])
def test_power_of_two_ratios_float_input(ak, O, ch, src, interp):
    rng = np.random.default_rng(ch * 7 + src // 48000)
    n = 70001
    x = (rng.standard_normal((n, ch)) * 0.5).astype("<f4")          # exceeds [-1, 1]: copied unclamped, then normalized
    for mono in (False, True):
        got = ak.preload(x.tobytes(), 32, "float", ch, src, 48000, interp, mono, 0.9)
        r = O.resample(O.pcm(x, 32, "float", ch), src, 48000, interp)
        ref = O.normalize(O.mono(r) if mono else r, 0.9)
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= TOL
    # standalone Audio:resample: the selected samples, bit for bit
    a = ak.pcm(x.tobytes(), 32, "float", ch, src).resample(48000, interp).numpy()
    assert f32_equal_bits(a, O.resample(O.pcm(x, 32, "float", ch), src, 48000, interp).astype(np.float32))


def test_decimation_nan_inf_and_silence(ak, O):
    x = (np.random.default_rng(1).standard_normal((4001, 8)) * 0.3).astype("<f4")
    x[10, 3] = np.nan            # frame 10 is selected (even): Lua's math.max skips NaN (A:3441), the product stays NaN
    x[12, 0] = np.inf            # max = inf -> mult = 0: finite samples become 0, inf * 0 = NaN
    got = ak.preload(x.tobytes(), 32, "float", 8, 96000, 48000, "cubic", False, 1.0)
    ref = O.normalize(O.resample(O.pcm(x, 32, "float", 8), 96000, 48000, "cubic"), 1.0)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    fin = ~np.isnan(ref)
    assert np.max(np.abs(got[fin] - ref[fin])) <= TOL
    z = np.zeros((1000, 8), dtype="<f4")
    assert np.all(np.isnan(ak.preload(z.tobytes(), 32, "float", 8, 96000, 48000, "cubic", False, 1.0)))   # 0 * inf (A:3444)


def test_config5_time_shards_equal_the_single_pass(ak):
    """8-channel f32 96 -> 48 kHz over 3 shards with a MAX-combined peak == one pass, bit for bit."""
    import torch
    lib, ctx = ak._lib.load(), ak.context()
    n = 300007
    x = (np.random.default_rng(8).standard_normal((n, 8)) * 0.25).astype("<f4")
    whole = ak.preload(x.tobytes(), 32, "float", 8, 96000, 48000, "cubic", False, 1.0)
    n_out = whole.shape[1]
    ctx.use_torch_stream()
    try:
        dmax = torch.zeros(1, device="cuda")
        descs, ins = [], []
        for r in range(3):
            o0, o1 = n_out * r // 3, n_out * (r + 1) // 3
            f, c = C.c_uint64(), C.c_uint64()
            lib.aukit_resample_window(n, 96000.0, 48000.0, 2, o0, o1 - o0, C.byref(f), C.byref(c))
            ins.append(torch.from_numpy(x[f.value: f.value + c.value].copy()).cuda())
            descs.append(ak.PipelineDesc(32, 2, 8, 0, 96000.0, 48000.0, 2, 0, n, f.value, c.value, o0, o1 - o0))
        for d, t in zip(descs, ins):
            ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(d), t.data_ptr(), dmax.data_ptr()))
        outs = []
        for d, t in zip(descs, ins):
            o = torch.empty((8, d.n_out), device="cuda")
            ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d), t.data_ptr(), 1.0, dmax.data_ptr(), o.data_ptr(), d.n_out))
            outs.append(o)
        torch.cuda.synchronize()
        got = np.concatenate([o.cpu().numpy() for o in outs], axis=1)
    finally:
        ctx.set_stream(None)
    assert f32_equal_bits(got, whole)


# ------------------------------------------------------------------ K16: wide float frames (config 5': 96 -> 44.1 kHz, 8 channels)
@pytest.mark.parametrize("ch", [8, 4])
@pytest.mark.parametrize("src,dst", [(96000, 44100), (44100, 48000), (48000, 8000), (44056.5, 48000)])
@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
def test_wide_float_frames_against_oracle(ak, O, ch, src, dst, interp):
    rng = np.random.default_rng(ch + int(src))
    n = 40009
    x = (rng.standard_normal((n, ch)) * 0.6).astype("<f4")          # exceeds [-1, 1]: hit (unclamped) vs near hit (clamped) is visible
    for mono in (False, True):
        got = ak.preload(x.tobytes(), 32, "float", ch, src, dst, interp, mono, 0.9)
        r = O.resample(O.pcm(x, 32, "float", ch), src, dst, interp)
        ref = O.normalize(O.mono(r) if mono else r, 0.9)
        assert got.shape == ref.shape
        assert np.max(np.abs(got - ref)) <= (2.0 ** -22 if interp == "none" else TOL * 4)   # |values| reach ~4 before normalize


@pytest.mark.parametrize("interp", ["none", "linear", "cubic"])
def test_wide_float_frames_at_huge_positions_and_shard_equality(ak, interp):
    """Shards of a 24 h 96 kHz buffer (positions up to 2^33): the kernel evaluates the reference's fp64 position itself,
    so it must follow the numpy restatement of A:653-673 anywhere, and two abutting shards must equal one shard."""
    import torch
    from util import ref_resample_window
    lib, ctx = ak._lib.load(), ak.context()
    n_total, src, dst, CH = 8_294_400_000, 96000, 44100, 8
    mode = {"none": 0, "linear": 1, "cubic": 2}[interp]
    total_out = int(lib.aukit_resample_out_len(n_total, float(src), float(dst)))
    rng = np.random.default_rng(mode)
    ctx.use_torch_stream()
    try:
        for o0 in (total_out - 60_000, int(total_out * 0.613), int(2 ** 30.5 * dst / src) - 30_000, 0):
            cnt = 60_000
            f, c = C.c_uint64(), C.c_uint64()
            assert lib.aukit_resample_window(n_total, float(src), float(dst), mode, o0, cnt, C.byref(f), C.byref(c)) == 0
            x = (rng.standard_normal((int(c.value), CH)) * 0.3).astype(np.float32)
            ref = ref_resample_window(x.T.astype(np.float64), int(f.value), n_total, src, dst, o0, cnt, interp)
            want = np.clip(ref * (1.0 / np.max(np.abs(ref))), -1, 1)
            t = torch.from_numpy(x).cuda()
            dmax = torch.zeros(1, device="cuda")
            out = torch.empty((CH, cnt), device="cuda")
            d = ak.PipelineDesc(32, 2, CH, 0, float(src), float(dst), mode, 0, n_total, f.value, c.value, o0, cnt)
            ak._lib.check(lib.aukit_cuda_dev_pipeline_peak(ctx.handle, C.byref(d), t.data_ptr(), dmax.data_ptr()))
            ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d), t.data_ptr(), 1.0, dmax.data_ptr(), out.data_ptr(), cnt))
            got = out.cpu().numpy()
            assert np.max(np.abs(got - want)) <= TOL, (o0, float(np.max(np.abs(got - want))))
            # the same range as two shards (each with its own window) == one shard, bit for bit
            parts = []
            for a0, a1 in ((o0, o0 + 25_001), (o0 + 25_001, o0 + cnt)):
                f2, c2 = C.c_uint64(), C.c_uint64()
                lib.aukit_resample_window(n_total, float(src), float(dst), mode, a0, a1 - a0, C.byref(f2), C.byref(c2))
                t2 = torch.from_numpy(x[f2.value - f.value: f2.value - f.value + c2.value].copy()).cuda()
                o2 = torch.empty((CH, a1 - a0), device="cuda")
                d2 = ak.PipelineDesc(32, 2, CH, 0, float(src), float(dst), mode, 0, n_total, f2.value, c2.value, a0, a1 - a0)
                ak._lib.check(lib.aukit_cuda_dev_pipeline_apply(ctx.handle, C.byref(d2), t2.data_ptr(), 1.0, dmax.data_ptr(), o2.data_ptr(), a1 - a0))
                parts.append(o2.cpu().numpy())
            assert f32_equal_bits(np.concatenate(parts, axis=1), got)
    finally:
        ctx.set_stream(None)
