"""Golden vectors produced by the reference ITSELF: tests/golden/reference_vectors.npz holds the
outputs of the unmodified /root/reference/aukit.lua (executed in oracle/luavm by
tests/golden/generate.py) for 150+ seeded calls covering every hot-path function.

  * CPU (`-m "not gpu"`): the C oracle must reproduce every vector BIT-EXACTLY in float64 and raise
    the reference's error where the reference raised -- this is what pins the oracle.
  * GPU (`-m gpu`): the CUDA path must match the same vectors (decode: f32 == (float)ref; float
    stages: |f32 - ref| <= 2^-20), through the Python mirror of the reference API.
"""
import json
import os
import re

import numpy as np
import pytest

from util import TOL, f32_equal_bits

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz")
_Z = np.load(GOLDEN)
MANIFEST = json.loads(_Z["manifest"].tobytes().decode())
IDS = [m["name"] for m in MANIFEST]


def blob(i, key):
    k = "c%d/%s" % (i, key)
    return _Z[k] if k in _Z.files else None


def expected(i, m):
    return [_Z["c%d/out%d" % (i, c)] for c in range(m["channels"])]


def same_f64(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b)))) and \
        bool(np.all(np.signbit(a) == np.signbit(b)))


# ------------------------------------------------------------------ oracle replay (CPU)
def oracle_run(O, m, i):
    A, op = m["args"], m["op"]
    data = blob(i, "in")
    raw = data.tobytes() if data is not None else None
    x = blob(i, "x")
    if op == "pcm":
        return list(O.pcm(raw, A["bitDepth"], A["dataType"], A["channels"], A["interleaved"], A["bigEndian"]))
    if op == "g711":
        return O.g711(raw, A["ulaw"], A["channels"])
    if op == "adpcm":
        return list(O.adpcm(raw, A["channels"], A["topFirst"], A["interleaved"], A.get("predictor"), A.get("step_index")))
    if op == "msadpcm":
        return list(O.msadpcm(raw, A["blockAlign"], A["channels"], A.get("coefficients")))
    if op == "wav":
        if A.get("head"):
            info = O.wav_parse(raw)
            return [np.zeros(0)] * info["channels"]
        out, _ = O.wav(raw)
        return list(out)
    if op in ("au", "aiff"):
        out, _ = (O.au(raw) if op == "au" else O.aiff(raw, bool(A.get("head"))))
        return list(out)
    if op == "resample":
        interp = A["interpolation"] or "linear"                   # aukit.defaultInterpolation (A:99)
        return list(O.resample(x, A["sampleRate"], A["targetRate"], interp))
    if op == "mono":
        return list(O.mono(x))
    if op == "amplify":
        return list(O.amplify(x, A["multiplier"]))
    if op == "normalize":
        return list(O.normalize(x, 1.0 if A["peak"] is None else A["peak"], bool(A["independent"])))
    if op == "lowpass":
        return list(O.lowpass(x, A["frequency"], A["sampleRate"]))
    if op == "highpass":
        return list(O.highpass(x, A["frequency"], A["sampleRate"]))
    if op == "invert":
        return list(O.invert(x))
    if op == "fade":
        return list(O.fade(x, A["sampleRate"], A["startTime"], A["startAmplitude"], A["endTime"], A["endAmplitude"]))
    if op == "delay":
        return list(O.delay(x, A["sampleRate"], A["delay"], 0.5 if A["multiplier"] is None else A["multiplier"]))
    if op == "center":
        return list(O.center(x, A["sampleRate"]))
    if op == "wav_out":
        return [np.frombuffer(O.wav_out(x, A["sampleRate"], A.get("bitDepth"), A.get("metadata")), dtype=np.uint8).astype(np.float64)]
    if op == "pcm_out":
        return [O.audio_pcm(x, 8 if A["bitDepth"] is None else A["bitDepth"], A["dataType"] or "signed",
                            True if A["interleaved"] is None else A["interleaved"])]
    if op == "chain":
        out, info = O.wav(raw)
        r = O.resample(out, info["sampleRate"], A["targetRate"], A["interpolation"])
        return list(O.normalize(O.mono(r), A["peak"], False))
    if op == "chain3":
        d = O.pcm(raw, A["bitDepth"], A["dataType"], A["channels"], True, A["bigEndian"])
        return list(O.amplify(O.resample(d, A["sampleRate"], A["targetRate"], A["interpolation"]), A["multiplier"]))
    if op == "chain5":
        d = O.pcm(raw, 32, "float", A["channels"], True, False)
        r = O.resample(d, A["sampleRate"], A["targetRate"], A["interpolation"])
        return list(O.normalize(r, 1.0 if A["peak"] is None else A["peak"], bool(A["independent"])))
    if op == "stream_adpcm":
        return list(O.stream_adpcm_48k(raw, A["blockAlign"], A["channels"]))
    if op == "stream_out":
        steps, total = O.audio_stream(x, A["sampleRate"], A["chunkSize"], A["bitDepth"], A["dataType"])
        return stream_carrier(steps, total, x.shape[0])
    raise AssertionError(op)


def stream_carrier(steps, total, nch):
    """The layout tests/golden/generate.py stores an Audio:stream run in: the channels' concatenated chunks, then the
    chunk positions, then [total length, chunk sizes...]."""
    chans = [np.concatenate([s[0][c] for s in steps]) if steps else np.zeros(0) for c in range(nch)]
    return chans + [np.array([s[1] for s in steps], dtype=np.float64),
                    np.array([total] + [float(len(s[0][0])) for s in steps], dtype=np.float64)]


@pytest.mark.parametrize("i", range(len(MANIFEST)), ids=IDS)
def test_oracle_reproduces_reference_bit_exactly(O, i):
    m = MANIFEST[i]
    if "error" in m:
        with pytest.raises(O.OracleError) as ei:
            oracle_run(O, m, i)
        want = re.sub(r"^aukit\.lua:\d+: ", "", m["error"])       # runtime errors carry the VM's "file:line:" prefix
        # cc.expect.range prints the number the Lua way ("120"), everything else is verbatim
        assert want.split(" (expected")[0] in str(ei.value) or want in str(ei.value), (want, str(ei.value))
        return
    got = oracle_run(O, m, i)
    exp = expected(i, m)
    assert len(got) == len(exp)
    for c, (g, e) in enumerate(zip(got, exp)):
        assert same_f64(g, e), "channel %d differs from the reference" % c


def test_reference_index_quirk_counts():
    """SURVEY finding 5 seen in the reference's own output: over the first 16000 outputs at
    44.1 -> 48 kHz the rational positions are integers 100 times; the reference's floor(x) is one
    lower for some of them (the ramp makes the selected index visible)."""
    i = IDS.index("resample_index_quirk_none")
    out = _Z["c%d/out0" % i]
    ramp = _Z["c%d/x" % i][0]
    n = np.arange(16000)
    exact = (n * 147) % 160 == 0
    rational = ramp[(n * 147) // 160]
    lower = out != rational
    assert exact.sum() == 100 and lower.sum() > 0 and np.all(exact[lower])     # only rational hits can differ
    assert np.array_equal(out[lower], ramp[(n * 147) // 160 - 1][lower])


# ------------------------------------------------------------------ CUDA replay (GPU)
def cuda_run(ak, m, i):
    A, op = m["args"], m["op"]
    data = blob(i, "in")
    raw = data.tobytes() if data is not None else None
    x = blob(i, "x")
    if op == "pcm":
        return ak.pcm(raw, A["bitDepth"], A["dataType"], A["channels"], A["sampleRate"], A["interleaved"], A["bigEndian"])
    if op == "g711":
        return ak.g711(raw, A["ulaw"], A["channels"], A.get("sampleRate"))
    if op == "adpcm":
        return ak.adpcm(raw, A["channels"], A["sampleRate"], A["topFirst"], A["interleaved"], A.get("predictor"), A.get("step_index"))
    if op == "msadpcm":
        return ak.msadpcm(raw, A["blockAlign"], A["channels"], A["sampleRate"], A.get("coefficients"))
    if op == "wav":
        return ak.wav(raw, bool(A.get("head")))
    if op == "au":
        return ak.au(raw)
    if op == "aiff":
        return ak.aiff(raw, bool(A.get("head")))
    if op == "wav_out":
        a = ak.Audio.from_numpy(x.astype(np.float32), A["sampleRate"])
        a.metadata = dict(A.get("metadata") or {})
        return np.frombuffer(a.wav(A.get("bitDepth"), "floor", ak.DIALECT_LITERAL), dtype=np.uint8).astype(np.float64)
    if op == "pcm_out":
        return ak.Audio.from_numpy(x.astype(np.float32), A["sampleRate"]).pcm(A["bitDepth"], A["dataType"], A["interleaved"])
    if op == "chain3":
        a = ak.pcm(raw, A["bitDepth"], A["dataType"], A["channels"], A["sampleRate"], True, A["bigEndian"])
        return ak.effects.amplify(a.resample(A["targetRate"], A["interpolation"]), A["multiplier"])
    if op == "chain5":
        a = ak.pcm(raw, 32, "float", A["channels"], A["sampleRate"], True, False).resample(A["targetRate"], A["interpolation"])
        args = [] if A.get("peak") is None and A.get("independent") is None else [A.get("peak"), A.get("independent")]
        return ak.effects.normalize(a, *args)
    if op == "stream_adpcm":
        import ctypes as C
        ctx = ak.context()
        buf = np.frombuffer(raw, dtype=np.uint8)
        out = C.c_void_p()
        ak._lib.check(ctx.lib.aukit_cuda_ima_adpcm_wav(ctx.handle, C.c_void_p(buf.ctypes.data), buf.size, A["blockAlign"], A["channels"],
                                                       48000.0, ak.DIALECT_GENERAL, C.byref(out)))
        v = ak.Audio(ctx, out).numpy().astype(np.float64)
        p = np.where(v < 0, np.rint(v * 32768.0), np.rint(v * 32767.0))
        return np.clip(np.floor(np.where(p < 0, p / 128.0, p / 127.0)), -128, 127)[:, : v.shape[1] - 8]
    if op == "stream_out":
        a = ak.Audio.from_numpy(x.astype(np.float32), A["sampleRate"])
        it, total = a.stream(A["chunkSize"], A["bitDepth"], A["dataType"])
        return stream_carrier(list(it), total, x.shape[0])
    a = ak.wav(raw) if op == "chain" else ak.Audio.from_numpy(x.astype(np.float32), A["sampleRate"])
    if op in ("resample", "chain"):
        a = a.resample(A["targetRate"], A.get("interpolation"))
    if op in ("mono", "chain"):
        a = a.mono()
    if op == "amplify":
        assert ak.effects.amplify(a, A["multiplier"]) is a
    if op == "lowpass":
        assert ak.effects.lowpass(a, A["frequency"]) is a
    if op == "highpass":
        assert ak.effects.highpass(a, A["frequency"]) is a
    if op == "invert":
        assert ak.effects.invert(a) is a
    if op == "fade":
        assert ak.effects.fade(a, A["startTime"], A["startAmplitude"], A["endTime"], A["endAmplitude"]) is a
    if op == "delay":
        assert ak.effects.delay(a, A["delay"], A["multiplier"]) is a
    if op == "center":
        assert ak.effects.center(a) is a
    if op in ("normalize", "chain"):
        args = [] if A.get("peak") is None and A.get("independent") is None else [A.get("peak"), A.get("independent")]
        assert ak.effects.normalize(a, *args) is a
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(MANIFEST)), ids=IDS)
def test_cuda_matches_reference(ak, i):
    m = MANIFEST[i]
    if "error" in m:
        with pytest.raises(ak.AukitError) as ei:
            a = cuda_run(ak, m, i)
            a.numpy()                                   # device-detected errors surface at the first sync
        want = re.sub(r"^aukit\.lua:\d+: ", "", m["error"]).split(" (expected")[0]
        assert want in str(ei.value), (m["error"], str(ei.value))
        return
    a = cuda_run(ak, m, i)
    exp = expected(i, m)
    if m["op"] == "wav_out":
        assert np.array_equal(a, exp[0]), "Audio:wav bytes differ from the reference's"
        return
    if m["op"] == "pcm_out":
        # the inputs are f32-representable, so the fp64 products are the reference's own numbers
        assert a.dtype == np.float64 and same_f64(a, exp[0])
        return
    if m["op"] in ("stream_adpcm", "stream_out"):
        # integer predictors / fp64 products of f32-representable inputs: the reference's own numbers, bit for bit
        assert len(a) == len(exp)
        for g, e in zip(a, exp):
            assert same_f64(g, e), m["name"]
        return
    assert a.channels() == m["channels"]
    assert float(a.sampleRate) == float(m["sampleRate"])
    decode = m["op"] in ("pcm", "g711", "adpcm", "msadpcm", "wav", "au", "aiff")
    x = blob(i, "x")
    for c, e in enumerate(exp):
        g = a.data[c]
        assert g.shape == e.shape
        if decode:
            assert f32_equal_bits(g, e.astype(np.float32)), "decode must be bit-exact"
        else:
            assert np.array_equal(np.isnan(g), np.isnan(e))
            fin = np.isfinite(e)
            assert np.array_equal(g[~fin & ~np.isnan(e)], e[~fin & ~np.isnan(e)])          # infinities keep their sign
            scale = max(1.0, float(np.max(np.abs(e[fin])))) if fin.any() else 1.0
            err = float(np.max(np.abs(g[fin] - e[fin]))) if fin.any() else 0.0
            # north_star's bound for the float stages is 2^-20 absolute (per unit of signal scale where an effect
            # leaves [-1, 1]).  Where the test hands float64 inputs to the device as float32, that narrowing alone
            # moves the reference's own result by up to 2^-25 |x| gain (gain <= 2: Catmull-Rom weights sum to 1.25 in
            # magnitude, normalize / fade / delay scale by <= 2 here); it is added explicitly instead of doubling TOL.
            narrowed = x is not None and np.issubdtype(np.asarray(x).dtype, np.floating)
            xs = float(np.max(np.abs(np.asarray(x)[np.isfinite(np.asarray(x))]))) if narrowed else 0.0
            bound = TOL * scale + ((2.0 ** -24) * max(xs, scale) if narrowed else 0.0)
            assert err <= bound, (m["name"], float(err), bound)
    if m["op"] in ("au", "aiff"):
        got_meta = {k: (v.decode("latin-1") if isinstance(v, bytes) else v) for k, v in a.metadata.items()}
        assert got_meta == m["metadata"] and a.info == m["info"]
    if m["op"] == "wav" and not m["args"].get("head"):
        want_meta = {k: v for k, v in m["metadata"].items()}
        got_meta = {k: (v.decode("latin-1") if isinstance(v, bytes) else v) for k, v in a.metadata.items()}
        assert got_meta == want_meta
        assert a.info.get("dataType") == m["info"].get("dataType")
    if m["op"] == "pcm":
        assert a.info == {"bitDepth": m["info"]["bitDepth"], "dataType": m["info"]["dataType"]}
