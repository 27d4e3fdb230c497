"""Parity of the integer decode stages against the oracle: BIT-EXACT (f32 == (float)ref_double)."""
import struct

import numpy as np
import pytest

from util import f32_equal_bits, ima_blocks, ms_blocks, riff, fmt_chunk, wav_pcm

pytestmark = pytest.mark.gpu


def _check(got, ref):
    ref32 = np.asarray(ref, dtype=np.float64).astype(np.float32)
    assert got.shape == ref32.shape, (got.shape, ref32.shape)
    if not f32_equal_bits(got, ref32):
        bad = np.argwhere(got.view(np.uint32) != ref32.view(np.uint32))
        raise AssertionError("first mismatch at %s: got %r want %r (%d bad)" % (bad[0], got[tuple(bad[0])], ref32[tuple(bad[0])], len(bad)))


@pytest.mark.parametrize("bits,dtype", [(8, "signed"), (8, "unsigned"), (16, "signed"), (16, "unsigned"),
                                        (24, "signed"), (24, "unsigned"), (32, "signed"), (32, "unsigned"), (32, "float")])
@pytest.mark.parametrize("be", [False, True])
@pytest.mark.parametrize("channels,interleaved", [(1, True), (2, True), (2, False), (3, True), (8, True), (5, False)])
def test_pcm_all_formats(ak, O, bits, dtype, be, channels, interleaved):
    rng = np.random.default_rng(bits * 100 + channels)
    frames = 4099                                    # odd: exercises the vector path and the scalar tail
    raw = rng.integers(0, 256, frames * channels * bits // 8, dtype=np.uint8)
    if dtype == "float":
        f = (rng.standard_normal(frames * channels) * 0.5).astype(">f4" if be else "<f4")
        f[:4] = [np.inf, -np.inf, 2.5, -0.0]
        raw = f.view(np.uint8)
    a = ak.pcm(raw, bits, dtype, channels, 44100, interleaved, be)
    assert a.channels() == channels and a.frames == frames and a.sampleRate == 44100
    assert a.info == {"bitDepth": bits, "dataType": dtype}
    _check(a.numpy(), O.pcm(raw, bits, dtype, channels, interleaved, be))


def test_pcm_exhaustive_s16_u8_and_s24_scaling(ak, O):
    s16 = np.arange(-32768, 32768, dtype="<i2")
    _check(ak.pcm(s16, 16, "signed").numpy(), O.pcm(s16, 16, "signed"))
    u8 = np.arange(256, dtype=np.uint8)
    _check(ak.pcm(u8, 8, "unsigned").numpy(), O.pcm(u8, 8, "unsigned"))
    _check(ak.pcm(u8, 8, "signed").numpy(), O.pcm(u8, 8, "signed"))
    u16 = np.arange(65536, dtype="<u2")
    _check(ak.pcm(u16, 16, "unsigned").numpy(), O.pcm(u16, 16, "unsigned"))
    v = np.arange(0, 1 << 24, dtype="<u4")
    s24 = np.ascontiguousarray(v.view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1)   # all 2^24 values, LE
    _check(ak.pcm(s24, 24, "signed").numpy(), O.pcm(s24, 24, "signed"))
    _check(ak.pcm(s24, 24, "unsigned").numpy(), O.pcm(s24, 24, "unsigned"))


def test_pcm_empty_and_errors(ak):
    e = ak.pcm(b"", 16, "signed", 2)
    assert e.channels() == 2 and e.frames == 0 and e.numpy().shape == (2, 0)
    with pytest.raises(ak.AukitError, match="uneven amount of data per channel"):
        ak.pcm(b"\0" * 6, 16, "signed", 2)
    with pytest.raises(ak.AukitError, match="float audio must have 32-bit depth"):
        ak.pcm(b"\0" * 4, 16, "float")


@pytest.mark.parametrize("ulaw", [True, False])
@pytest.mark.parametrize("channels", [1, 2, 3])
def test_g711(ak, O, ulaw, channels):
    rng = np.random.default_rng(11)
    raw = np.concatenate([np.arange(256, dtype=np.uint8), rng.integers(0, 256, 10007, dtype=np.uint8)])   # ragged for C=2,3
    a = ak.g711(raw, ulaw, channels)
    ref = O.g711(raw, ulaw, channels)
    assert a.sampleRate == 8000 and a.metadata == {"bitDepth": 14 if ulaw else 13, "dataType": "signed"}
    for c in range(channels):
        got = a.data[c]
        assert got.shape == ref[c].shape
        assert f32_equal_bits(got, ref[c].astype(np.float32))       # includes the -0.0 of mu-law 0x7F
        assert np.array_equal(np.signbit(got), np.signbit(ref[c]))


@pytest.mark.parametrize("channels,block_align,dialect", [(1, 256, 0), (1, 1024, 0), (2, 256, 0), (2, 1024, 0),
                                                          (1, 256, 1), (2, 512, 1), (8, 8192, 1), (6, 1536, 1), (1, 37, 0)])
def test_ima_adpcm_wav(ak, O, channels, block_align, dialect):
    nblocks = 67
    raw = ima_blocks(nblocks, block_align, channels, seed=channels + block_align)
    if dialect == 0 and channels == 1:
        raw = raw[: len(raw) - 11]                   # literal mono accepts a short last block (A:1546)
    blob = riff([(b"fmt ", fmt_chunk(0x11, channels, 22050, block_align, 4)), (b"data", raw.tobytes())])
    a = ak.wav(blob, dialect=dialect)
    assert a.sampleRate == 22050 and a.info == {"dataType": "adpcm", "bitDepth": 4}
    _check(a.numpy(), O.wav_ima(raw, block_align, channels, dialect))


def test_ima_errors(ak, O):
    raw = ima_blocks(4, 64, 2)
    raw[64 + 2] = 200                                # step index > 88 in block 1 (A:1213)
    with pytest.raises(ak.AukitError, match="outside of range"):
        ak._aukit.Audio(ak.context(), _ima_handle(ak, raw, 64, 2)).numpy()
    with pytest.raises(ak.AukitError, match="table too short"):
        _ima_handle(ak, ima_blocks(2, 96, 3), 96, 3)
    with pytest.raises(ak.AukitError, match="band"):
        _ima_handle(ak, raw[:100], 64, 2)            # stereo: a short last block reads nil bytes


def _ima_handle(ak, raw, block_align, channels, dialect=0):
    import ctypes as C
    ctx = ak.context()
    out = C.c_void_p()
    raw = np.ascontiguousarray(raw)
    ak._lib.check(ctx.lib.aukit_cuda_ima_adpcm_wav(ctx.handle, C.c_void_p(raw.ctypes.data), raw.size, block_align, channels,
                                                   22050.0, dialect, C.byref(out)))
    return out


def test_adpcm_headerless(ak, O):
    rng = np.random.default_rng(12)
    raw = rng.integers(0, 256, 3001, dtype=np.uint8)
    for ch, top, il, pred, idx in [(1, True, True, None, None), (2, False, True, [100, -200], [5, 60]),
                                   (3, True, False, [0, 1, 2], [88, 0, 44]), (1, False, False, 77, 3)]:
        a = ak.adpcm(raw, ch, 48000, top, il, pred, idx)
        p = [pred] if isinstance(pred, int) else pred
        i = [idx] if isinstance(idx, int) else idx
        _check(a.numpy(), O.adpcm(raw, ch, top, il, p, i))
    with pytest.raises(ak.AukitError, match="table too short"):
        ak.adpcm(raw, 2, 48000, True, True, 5, None)


@pytest.mark.parametrize("channels,block_align,dialect,tame", [(1, 256, 0, True), (2, 256, 0, True), (2, 1024, 0, False),
                                                               (1, 512, 1, True), (8, 8192, 1, True), (8, 2048, 1, False),
                                                               (3, 300, 1, True), (1, 1024, 0, False)])
def test_msadpcm(ak, O, channels, block_align, dialect, tame):
    raw = ms_blocks(53, block_align, channels, seed=block_align + channels, tame=tame)
    a = ak.msadpcm(raw, block_align, channels, 44100, None, dialect)
    ref = O.msadpcm(raw, block_align, channels, None, dialect)
    got = a.numpy()
    # untamed random nibbles drive delta beyond 2^31 (the reference carries it as a double, can reach inf/NaN):
    # the fp64 continuation must still match bit for bit
    _check(got, ref)


@pytest.mark.parametrize("channels,block_align,nblocks,tame", [
    (1, 256, 37, True), (1, 258, 5, False), (2, 256, 33, True), (2, 1028, 7, False), (4, 512, 9, True), (4, 520, 3, False),
    (8, 2048, 5, True), (8, 2064, 4, False), (3, 300, 11, True), (5, 1000, 3, False), (6, 516, 7, True), (2, 254, 9, True),
    (8, 8192, 9, True), (8, 8192, 6, False), (8, 112, 21, True), (8, 80, 3, True), (8, 1040, 14, False), (8, 8240, 5, True)])
def test_msadpcm_tiled_and_chain_kernels_agree(ak, O, monkeypatch, channels, block_align, nblocks, tame):
    """The staged 8-channel kernel (records through shared memory, speculated straight-line periods, output segments
    aligned in absolute address: rows that start at every 16-byte phase of a 256-byte segment, blocks shorter than one
    segment), the warp-tiled kernel (aligned record loads for 1/2/4/8 channels, byte reads otherwise), the
    chain-per-lane kernel and the oracle agree bit for bit; chains not a multiple of 32, partial last
    flush, deltas that leave the 32-bit fast path mid-block (the speculated period is redone)."""
    raw = ms_blocks(nblocks, block_align, channels, seed=7 * block_align + channels, tame=tame)
    blocks = raw.reshape(nblocks, block_align)
    blocks[::3, channels:3 * channels] = np.array([0xFF, 0x7F] * channels, dtype=np.uint8)      # delta 32767 in the header
    blocks[1::3, 7 * channels:7 * channels + 8] = 0x88                                           # nibble -8: delta triples
    ref = O.msadpcm(raw, block_align, channels, None, 1)
    got = ak.msadpcm(raw, block_align, channels, 44100, None, 1).numpy()
    _check(got, ref)
    monkeypatch.setenv("AUKIT_DISABLE_STAGED_ADPCM", "1")
    _check(ak.msadpcm(raw, block_align, channels, 44100, None, 1).numpy(), ref)
    monkeypatch.setenv("AUKIT_DISABLE_TILED_ADPCM", "1")
    _check(ak.msadpcm(raw, block_align, channels, 44100, None, 1).numpy(), ref)


@pytest.mark.parametrize("channels,block_align,nblocks", [(1, 256, 37), (1, 36, 70), (2, 264, 33), (2, 1024, 3), (3, 600, 9),
                                                          (8, 2048, 5), (8, 96, 13), (5, 40, 7), (8, 8192, 7), (8, 8224, 5), (8, 64, 40),
                                                          (6, 1560, 9)])
def test_ima_tiled_and_chain_kernels_agree(ak, O, monkeypatch, channels, block_align, nblocks):
    """Transition-table + transposed-store kernel vs the chain-per-lane kernel vs the oracle: odd and
    even group counts, a single group, chains not a multiple of 32, every step index reached."""
    raw = ima_blocks(nblocks, block_align, channels, seed=block_align + channels)
    rng = np.random.default_rng(block_align)
    blocks = raw.reshape(nblocks, block_align)
    blocks[:, 4 * channels:] = rng.integers(0, 256, blocks[:, 4 * channels:].shape, dtype=np.uint8)   # wild nibbles
    ref = O.wav_ima(raw, block_align, channels, 1)
    got = ak._aukit.Audio(ak.context(), _ima_handle(ak, raw, block_align, channels, 1)).numpy()
    _check(got, ref)
    monkeypatch.setenv("AUKIT_DISABLE_TILED_ADPCM", "1")
    _check(ak._aukit.Audio(ak.context(), _ima_handle(ak, raw, block_align, channels, 1)).numpy(), ref)


def test_msadpcm_custom_coefficients_and_wav(ak, O):
    coefs = [[256, 512, 0, 192, 240, 460, 392, -300], [0, -256, 0, 64, 0, -208, -232, 77]]
    raw = ms_blocks(9, 256, 2, seed=3)
    raw.reshape(9, 256)[:, :2] = 7                   # use the extra coefficient pair
    extra = struct.pack("<HHH", 4 + 4 * 8, 500, 8) + b"".join(struct.pack("<hh", a, b) for a, b in zip(*coefs))
    blob = riff([(b"fmt ", fmt_chunk(2, 2, 11025, 256, 4, extra)), (b"data", raw.tobytes())])
    a = ak.wav(blob)
    assert a.info == {"dataType": "msadpcm", "bitDepth": 4}
    _check(a.numpy(), O.msadpcm(raw, 256, 2, coefs))
    with pytest.raises(ak.AukitError, match="Unsupported number of channels: 3"):
        ak.msadpcm(ms_blocks(2, 64, 3), 64, 3)
    with pytest.raises(ak.AukitError, match="nil"):
        bad = ms_blocks(2, 64, 2)
        bad[0] = 9
        ak.msadpcm(bad, 64, 2).numpy()


def test_wav_pcm_and_g711_dispatch(ak, O):
    rng = np.random.default_rng(13)
    for bits, fmt, ch in ((8, 1, 2), (16, 1, 2), (24, 1, 1), (32, 1, 2), (32, 3, 2), (8, 6, 2), (8, 7, 1)):
        payload = rng.integers(0, 256, 4800 * ch * bits // 8, dtype=np.uint8)
        if fmt == 3:
            payload = rng.standard_normal(4800 * ch).astype("<f4").view(np.uint8)
        blob = wav_pcm(payload.tobytes(), ch, 32000, bits, fmt)
        a = ak.wav(blob)
        ref, info = O.wav(blob)
        assert a.info["dataType"] == info["dataType"] and a.sampleRate == 32000 and a.channels() == ch
        _check(a.numpy(), np.stack(ref) if isinstance(ref, list) else ref)
    h = ak.wav(wav_pcm(bytes(400), 2, 32000, 16), head=True)
    assert h.channels() == 2 and h.frames == 0
