"""Known-answer anchors for the oracle, hand-derived from the cited
reference lines -- the reference itself ships no tests or fixtures."""
import struct

import numpy as np
import pytest


def test_pcm_s16_le(O):                                   # A:1133
    x = O.pcm(bytes([0x00, 0x80, 0xFF, 0x7F, 0x01, 0x00, 0xFF, 0xFF]), 16, "signed")[0]
    assert x.tolist() == [-1.0, 1.0, 1 / 32767, -1 / 32768]


def test_pcm_u8_and_u16_quirk(O):                         # A:1152: literal 128 at every depth
    assert O.pcm(bytes([0, 128, 255]), 8, "unsigned")[0].tolist() == [-1.0, 0.0, 1.0]
    u16 = np.array([0, 127, 128, 65535], dtype="<u2")
    assert O.pcm(u16, 16, "unsigned")[0].tolist() == [-0.00390625, -3.0517578125e-05, 0.0, 1.9961241492965485]


def test_pcm_s24_be(O):                                   # A:1068
    assert O.pcm(bytes([0x80, 0, 0, 0x7F, 0xFF, 0xFF]), 24, "signed", 1, True, True)[0].tolist() == [-1.0, 1.0]


def test_pcm_float_passthrough_and_layouts(O):
    f = np.array([0.5, -2.0, np.inf, 1e-30], dtype="<f4")
    assert O.pcm(f, 32, "float")[0].tolist() == [0.5, -2.0, np.inf, float(np.float32(1e-30))]
    s = np.arange(12, dtype="<i2")
    il = O.pcm(s, 16, "signed", 3, True)
    pl = O.pcm(s, 16, "signed", 3, False)
    assert (il * 32767).round().astype(int).tolist() == [[0, 3, 6, 9], [1, 4, 7, 10], [2, 5, 8, 11]]
    assert (pl * 32767).round().astype(int).tolist() == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]]


def test_pcm_errors(O):                                   # A:1058-1064
    with pytest.raises(O.OracleError, match=r"bad argument #2 \(invalid bit depth\)"):
        O.pcm(b"\0", 12)
    with pytest.raises(O.OracleError, match=r"float audio must have 32-bit depth"):
        O.pcm(b"\0\0", 16, "float")
    with pytest.raises(O.OracleError, match=r"uneven amount of data per channel"):
        O.pcm(b"\0\0\0", 16, "signed", 1)
    with pytest.raises(O.OracleError, match=r"uneven amount of data per channel"):
        O.pcm(b"\0" * 6, 16, "signed", 2)
    assert O.pcm(b"", 16).shape == (1, 0)


def test_g711_anchors(O):                                 # A:1374-1379
    mu = O.g711(bytes([0x00, 0x7F, 0x80, 0xFF]), True)[0]
    assert mu.tolist() == [-0.9803466796875, -0.0, 0.9803466796875, 0.0]
    assert np.signbit(mu[1]) and not np.signbit(mu[3])
    al = O.g711(bytes([0x55, 0xD5, 0x2A, 0xAA, 0x00, 0x80]), False)[0]
    assert al.tolist() == [-0.000244140625, 0.000244140625, -0.984375, 0.984375, -0.16796875, 0.16796875]


def test_g711_ragged_channels(O):                         # A:1379 byte k -> channel k % C
    out = O.g711(bytes(range(7)), True, 3)
    assert [len(c) for c in out] == [3, 2, 2]


def test_g711_exact_in_f32(O):
    for ulaw in (True, False):
        v = O.g711(bytes(range(256)), ulaw)[0]
        assert np.all(v.astype(np.float32).astype(np.float64) == v)
        assert np.all(v * 8192 == np.round(v * 8192))


def test_ima_step_anchors(O):                             # A:1250-1255
    assert O.ima_step(7, 0, 0) == (12 / 32767, 12, 8)
    assert O.ima_step(0xF, 0, 0) == (-12 / 32768, -12, 8)
    v, p, i = O.ima_step(1, 0, 0)                         # dialect: (7>>2)+(7>>3) = 1, not 2
    assert p == 1 and i == 0
    assert O.ima_step(7, 32760, 88)[1] == 32767           # clamp high
    assert O.ima_step(15, -32760, 88)[1] == -32768        # clamp low


def test_ima_wav_mono_mask_bug_and_header_not_emitted(O):  # A:1544, A:1546
    blk = struct.pack("<hBB", 100, 0x25, 0) + bytes([0x00] * 4)
    out = O.wav_ima(blk, 8, 1)
    assert out.shape == (1, 8)
    # index 0x25 & 0x0F = 5 -> step 12; nibble 0: diff = 12>>3 = 1 -> 101 (header value 100 is not output)
    assert round(out[0, 0] * 32767) == 101
    gen = O.wav_ima(blk, 8, 1, O.GENERAL)                  # no mask: index 37 -> step 253, diff 31
    assert round(gen[0, 0] * 32767) == 131


def test_ima_wav_stereo_layout(O):                        # A:1513-1541: 4 bytes L then 4 bytes R, low nibble first
    blk = struct.pack("<hBBhBB", 0, 0, 0, 0, 0, 0) + bytes([0x17, 0, 0, 0]) + bytes([0x8F, 0, 0, 0])
    out = O.wav_ima(blk, 16, 2)
    assert out.shape == (2, 8)
    l0, _, i = O.ima_step(7, 0, 0)
    l1 = O.ima_step(1, 12, i)[0]
    assert out[0, 0] == l0 and out[0, 1] == l1
    assert out[1, 0] == O.ima_step(0xF, 0, 0)[0]
    with pytest.raises(O.OracleError, match="outside of range"):
        O.wav_ima(struct.pack("<hBBhBB", 0, 89, 0, 0, 0, 0) + bytes(8), 16, 2)
    with pytest.raises(O.OracleError, match="table too short"):
        O.wav_ima(bytes(64), 32, 3)


def test_ms_anchors(O):                                   # A:1321-1324
    blk = struct.pack("<BhhhB", 1, 16, 100, 50, 0x30)
    assert (O.msadpcm(blk, 8, 1)[0] * 32767).round().tolist() == [50, 100, 198, 296]
    blk = struct.pack("<BhhhB", 3, 16, -1, 0, 0x00)       # floor(-192/256) = -1 (C truncation gives 0)
    assert round(O.msadpcm(blk, 8, 1)[0, 2] * 32768) == -1


def test_ms_mono_reuses_first_header(O):                  # A:1331 (bug): every block starts from block 1's header
    b1 = struct.pack("<BhhhB", 0, 16, 1000, 900, 0x11)
    b2 = struct.pack("<BhhhB", 1, 99, -5, -7, 0x11)
    out = O.msadpcm(b1 + b2, 8, 1)[0]
    assert out[:2].tolist() == out[4:6].tolist()
    gen = O.msadpcm(b1 + b2, 8, 1, dialect=O.GENERAL)[0]
    assert round(gen[4] * 32768) == -7


def test_ms_errors(O):
    with pytest.raises(O.OracleError, match="Unsupported number of channels: 3"):
        O.msadpcm(bytes(64), 32, 3)
    with pytest.raises(O.OracleError):
        O.msadpcm(struct.pack("<BhhhB", 9, 16, 0, 0, 0), 8, 1)   # predictor index 9 has no coefficients


def test_interpolation_anchors(O):                        # A:257-266 via resample at ratio 2 / 4
    d = np.array([0, 1, 0, -1.0])
    lin = O.resample(d, 1, 2, "linear")[0]                 # x = 1, 1.5, ... 4.5
    assert lin.tolist() == [0, 0.5, 1, 0.5, 0, -0.5, -1, -1]
    cub = O.resample(d, 1, 2, "cubic")[0]
    assert cub.tolist() == [0, 0.5625, 1, 0.625, 0, -0.5625, -1, -1.0]   # x=4.5 -> -1.0625 clamped (A:668)
    non = O.resample(d, 1, 2, "none")[0]
    assert non.tolist() == [0, 0, 1, 1, 0, 0, -1, -1]
    assert O.resample(np.array([0, 1, 0, -1.0]), 1, 4, "linear")[0, 13] == -1.0   # x = 4.25, d[5] nil -> d[4]


def test_exact_hits_are_not_clamped(O):                   # A:667
    d = np.array([2.0, -3.0, 0.5])
    out = O.resample(d, 1, 2, "linear")[0]
    assert out.tolist() == [2.0, -0.5, -3.0, -1.0, 0.5, 0.5]


def test_resample_length_and_index_quirk(O):              # A:658-667, SURVEY finding 5
    assert O.resample_len(441000, 44100, 48000) == 480000
    hits = below = 0
    for i in range(1, 480001):
        x = O.resample_pos(i, 44100, 48000)
        exact = ((i - 1) * 147) % 160 == 0
        if x % 1 == 0:
            hits += 1
            assert exact
        elif exact:
            below += 1
            assert int(x) == (i - 1) * 147 // 160          # one below the rational floor + 1
    assert (hits, below) == (828, 2172)


def test_mono_amplify_normalize(O):
    x = np.array([[0.1, 0.2, 0.3], [0.3, -0.2, 0.9]])
    assert O.mono(x)[0].tolist() == [(0 + 0.1 + 0.3) / 2, (0 + 0.2 - 0.2) / 2, (0 + 0.3 + 0.9) / 2]
    assert O.amplify(x, 1)[1].tolist() == x[1].tolist()
    assert O.amplify(x, 2)[1].tolist() == [0.6, -0.4, 1.0]
    n = O.normalize(x, 0.8)
    assert n[1, 2] == 0.9 * (0.8 / 0.9) and n[0, 0] == 0.1 * (0.8 / 0.9)
    ind = O.normalize(x, 1.0, True)
    assert ind[0, 2] == 0.3 * (1.0 / 0.3)
    assert np.isnan(O.normalize(np.zeros((1, 4)))).all()   # 0 * inf
    assert O.normalize(np.array([[np.nan, 0.5]]))[0, 1] == 1.0   # math.max ignores NaN


def test_encode_pcm_formula(O):                           # A:874
    assert O.encode_pcm(-1.0, 8) == -128 and O.encode_pcm(1.0, 8) == 127
    assert O.encode_pcm(0.5, 16) == 0.5 * 32767 and O.encode_pcm(0.0, 8, "unsigned") == 128
