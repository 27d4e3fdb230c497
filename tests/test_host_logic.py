"""Host-side logic of the product (no GPU needed): the aukit.wav container walk, the resample
length / position / halo-window helpers, and argument validation -- each against the oracle."""
import struct

import numpy as np
import pytest

from util import fmt_chunk, riff, wav_pcm


def _cmp_info(ak, O, blob):
    got, ref = ak.wav_info(blob), O.wav_parse(blob)
    for k in ("dataType", "channels", "sampleRate", "blockAlign", "coefficients", "data_off", "data_size"):
        assert got[k] == ref[k], k
    return got, ref


def test_wav_parse_pcm_variants(ak, O):
    for bits, fmt, dt in ((8, 1, "unsigned"), (16, 1, "signed"), (24, 1, "signed"), (32, 3, "float"), (8, 6, "alaw"), (8, 7, "ulaw")):
        got, _ = _cmp_info(ak, O, wav_pcm(bytes(24), 2, 22050, bits, fmt))
        assert got["dataType"] == dt and got["bitDepth"] == bits and got["data_off"] == 44


def test_wav_parse_last_data_chunk_wins_and_no_pad_byte(ak, O):
    blob = riff([(b"fmt ", fmt_chunk(1, 1, 8000, 1, 8)), (b"junk", b"abc"), (b"data", b"\1\2\3"), (b"fact", b"\0" * 4),
                 (b"data", b"\4\5")])
    got, _ = _cmp_info(ak, O, blob)
    assert got["data_size"] == 2 and blob[got["data_off"]: got["data_off"] + 2] == b"\4\5"


def test_wav_parse_msadpcm_coefficients_and_extensible(ak, O):
    coefs = [(256, 0), (512, -256), (0, 0), (192, 64), (240, 0), (460, -208), (392, -232), (100, -100)]
    extra = struct.pack("<HHH", 2 + 2 + 4 * len(coefs), 500, len(coefs)) + b"".join(struct.pack("<hh", *c) for c in coefs)
    blob = riff([(b"fmt ", fmt_chunk(2, 2, 22050, 256, 4, extra)), (b"data", bytes(256))])
    got, _ = _cmp_info(ak, O, blob)
    assert got["coefficients"] == [[c[0] for c in coefs], [c[1] for c in coefs]]
    guid_tail = bytes.fromhex("000010008000" "00aa00389b71")
    for code, dt in ((1, "signed"), (3, "float"), (6, "alaw"), (7, "ulaw"), (0x11, "adpcm"), (2, "msadpcm")):
        ext = struct.pack("<HHI", 22, 24, 3) + bytes([code, 0, 0, 0]) + guid_tail
        blob = riff([(b"fmt ", fmt_chunk(0xFFFE, 2, 48000, 6, 32, ext)), (b"data", bytes(12))])
        got, _ = _cmp_info(ak, O, blob)
        assert got["dataType"] == dt and got["bitDepth"] == 24       # valid-bits field replaces bitDepth (A:1494)


def test_wav_parse_info_tags(ak, O):
    tags = b"INFO" + b"INAM" + struct.pack("<I", 5) + b"Song\0" + b"\0" + b"ITRK" + struct.pack("<I", 2) + b"7\0" \
        + b"XXXX" + struct.pack("<I", 2) + b"zz"
    blob = riff([(b"LIST", tags), (b"fmt ", fmt_chunk(1, 1, 8000, 2, 16)), (b"data", bytes(4))])
    got, ref = _cmp_info(ak, O, blob)
    assert [t[0] for t in ref["tags"]] == ["INAM", "ITRK", "XXXX"]
    assert got["metadata"] == {"title": b"Song\0", "trackNumber": b"7\0"}   # "7\0" is not a Lua numeral


def test_wav_parse_errors(ak, O):
    cases = [
        (b"RIFX" + bytes(40), "not a WAV file"),
        (b"RIFF\0\0\0\0WAVX" + bytes(20), "not a WAV file"),
        (riff([(b"fmt ", fmt_chunk(85, 2, 44100, 4, 16)), (b"data", bytes(4))]), "unsupported WAV file"),
        (riff([(b"fmt ", fmt_chunk(1, 2, 44100, 4, 16))]), "invalid WAV file"),
        (riff([(b"fmt ", fmt_chunk(1, 2, 44100, 4, 16)), (b"data", bytes(4))])[:-2], "invalid WAV file"),
    ]
    for blob, msg in cases:
        with pytest.raises(ak.AukitError, match=msg):
            ak.wav_info(blob)
        with pytest.raises(O.OracleError, match=msg):
            O.wav_parse(blob)


def test_resample_len_and_position_match_oracle(ak, O):
    lib = ak._lib.load()
    rng = np.random.default_rng(7)
    pairs = [(44100, 48000), (22050, 48000), (96000, 48000), (11025, 48000), (48000, 44100), (8000, 48000),
             (44100, 44100), (48000, 8000), (44056, 48000), (32000, 44100)]
    for src, dst in pairs:
        for n in [0, 1, 2, 3, 147, 441000, 158760000, 8294400000] + rng.integers(1, 10**9, 20).tolist():
            assert lib.aukit_resample_out_len(n, src, dst) == O.resample_len(n, src, dst)
        for i in [1, 2, 160, 161, 480000, 2**33] + rng.integers(1, 2**34, 200).tolist():
            assert lib.aukit_resample_position(i, src, dst) == O.resample_pos(i, src, dst)


def test_resample_window_covers_every_tap(ak, O):
    import ctypes as C
    lib = ak._lib.load()
    rng = np.random.default_rng(8)
    halo = {0: (0, 0), 1: (0, 1), 2: (-1, 2)}
    for src, dst in [(44100, 48000), (96000, 44100), (22050, 48000), (48000, 44100), (96000, 48000)]:
        n_in = 100000
        n_out = O.resample_len(n_in, src, dst)
        for mode in (0, 1, 2):
            for _ in range(10):
                o0 = int(rng.integers(0, n_out - 1))
                cnt = int(rng.integers(1, min(5000, n_out - o0) + 1))
                f, c = C.c_uint64(), C.c_uint64()
                assert lib.aukit_resample_window(n_in, src, dst, mode, o0, cnt, C.byref(f), C.byref(c)) == 0
                lo, hi = halo[mode]
                for i in (o0 + 1, o0 + cnt):      # positions are monotone: the ends bound the window
                    fl = int(np.floor(O.resample_pos(i, src, dst)))
                    need_lo = max(1, fl + lo) - 1
                    need_hi = min(n_in, fl + hi) - 1
                    assert f.value <= need_lo and need_hi <= f.value + c.value - 1
                assert c.value <= cnt * src / dst + 6


def test_frame_count_helpers_match_oracle(ak, O):
    lib = ak._lib.load()
    Ol = O.lib()
    for nbytes in (0, 1, 7, 8, 1024, 1030, 4096, 8192 * 3, 8192 * 3 + 100):
        for ba in (8, 36, 256, 1024, 1020, 8192):
            for ch in (1, 2, 8):
                for dia in (0, 1):
                    if dia == 0 and ch > 2:
                        continue
                    assert lib.aukit_ima_adpcm_wav_frames(nbytes, ba, ch, dia) == Ol.auko_wav_ima_len(nbytes, ba, ch, dia), (nbytes, ba, ch, dia)
                assert lib.aukit_msadpcm_frames(nbytes, ba, ch) == Ol.auko_msadpcm_len(nbytes, ba, ch)


def test_argument_validation_strings(ak):
    # raised on the host, before any device work (so they can be checked without a GPU)
    with pytest.raises(ak.AukitError, match=r"bad argument #2 \(invalid bit depth\)"):
        ak.pcm(b"\0\0", 12)
    with pytest.raises(ak.AukitError, match=r"bad argument #3 \(invalid data type\)"):
        ak.pcm(b"\0\0", 16, "double")
    with pytest.raises(ak.AukitError, match=r"bad argument #2 \(expected number or nil, got string\)|bad argument #2 \(expected nil or number, got string\)"):
        ak.pcm(b"\0\0", "16")
    with pytest.raises(ak.AukitError, match=r"bad argument #1 \(expected Audio"):
        ak.effects.amplify("nope", 2)


def test_numpy_window_reference(O):
    """The numpy restatement used for sharded / huge-index GPU tests agrees with the C oracle."""
    from util import ref_resample_window
    rng = np.random.default_rng(3)
    x = rng.uniform(-1.2, 1.2, (2, 5003))
    for src, dst in ((44100, 48000), (96000, 44100), (8000, 48000), (48000, 48000)):
        for interp in ("none", "linear", "cubic"):
            ref = O.resample(x, src, dst, interp)
            got = ref_resample_window(x, 0, x.shape[1], src, dst, 0, ref.shape[1], interp)
            assert np.max(np.abs(got - ref)) <= 1e-15       # pow(fx, 3) vs fx**3: last-ulp differences only
            if interp == "none":
                assert np.array_equal(got, ref)


def test_static_run_kernel_half_period_script():
    """The straight-line K10 kernel (csrc/pipeline_run.cu, half_geom) bakes, per half period, how many outputs every
    input-frame step produces.  Restate its constexpr arithmetic and check it against the reference's own positions
    (A:666: x = (i-1)/ratio + 1, floor taken per output): every output of a period is produced exactly once, by the
    step whose frame is its floor position, at most CMIN + 1 per step, and the taps a half touches stay inside the
    32-period tile plus its -1/+2 halo."""
    for L, M in ((160, 147), (320, 147), (640, 147)):
        LH = L // 2
        produced = []
        for cls in (0, 1):
            eb = cls * LH
            fbase = eb * M // L
            nsteps = (eb + LH - 1) * M // L - fbase + 1
            for s in range(nsteps):
                f = fbase + s
                lo, hi = max(-(-f * L // M), eb), min(-(-(f + 1) * L // M), eb + LH)
                cnt = max(hi - lo, 0)
                assert cnt <= L // M + 1
                produced += [(e, f) for e in range(lo, lo + cnt)]
            # rows read: floor position - 1 .. + 2 relative to the period start -> within [-1, M + 1]
            assert fbase - 1 >= -1 and fbase + nsteps - 1 + 2 <= M + 1
        assert [e for e, _ in produced] == list(range(L))
        assert all(f == e * M // L for e, f in produced)


def test_fma_sample_scaling_equals_the_double_division(O):
    """csrc/pipeline_tile.cu::conv_sample converts s >= 0 as fma(lo, RN(1/(2^(b-1) - 1)), lo), lo = s * 2^-(b-1); that must
    be the reference's s / (2^(b-1) - 1) (A:1133) narrowed to float, for every 24-bit (and 16-bit) sample."""
    import ctypes as C
    f = O.lib().auko_selftest_fma_scale
    f.restype, f.argtypes = C.c_long, [C.c_int]
    assert f(24) == 0 and f(16) == 0
