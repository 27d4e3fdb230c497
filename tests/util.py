"""Synthetic inputs shared by the tests (seeded; nothing is read from /root/reference)."""
import struct

import numpy as np

TOL = 2.0 ** -20   # float-stage tolerance stated by BASELINE.json:north_star


def riff(chunks):
    body = b"WAVE" + b"".join(cid + struct.pack("<I", len(payload)) + payload for cid, payload in chunks)
    return b"RIFF" + struct.pack("<I", len(body)) + body


def fmt_chunk(fmt, channels, rate, block_align, bits, extra=b""):
    return struct.pack("<HHIIHH", fmt, channels, rate, rate * block_align, block_align, bits) + extra


def wav_pcm(payload, channels=2, rate=44100, bits=16, fmt=1, extra_chunks=()):
    ba = channels * bits // 8
    return riff([(b"fmt ", fmt_chunk(fmt, channels, rate, ba, bits))] + list(extra_chunks) + [(b"data", payload)])


def tone_s16(n, channels=2, rate=44100, seed=1, amp=0.5):
    """BASELINE config 1's signal: 440/660 Hz tones at half scale plus +-256 integer noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / rate
    out = np.empty((n, channels), dtype=np.int16)
    for c in range(channels):
        s = np.round(amp * 32767 * np.sin(2 * np.pi * (440 + 220 * c) * t)) + rng.integers(-256, 257, n)
        out[:, c] = np.clip(s, -32768, 32767).astype(np.int16)
    return out


def ima_blocks(nblocks, block_align, channels, seed=4, stereo_literal=False):
    """Random IMA ADPCM blocks with valid headers (index 0..88), general N-channel layout."""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, size=(nblocks, block_align), dtype=np.uint8)
    for c in range(channels):
        data[:, 4 * c + 2] = rng.integers(0, 89, nblocks)
    return data.reshape(-1)


def ms_blocks(nblocks, block_align, channels, seed=5, tame=True):
    """Random MS-ADPCM blocks with valid headers (predictor 0..6, delta 16..2047)."""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, size=(nblocks, block_align), dtype=np.uint8)
    C = channels
    data[:, :C] = rng.integers(0, 7, (nblocks, C))
    delta = rng.integers(16, 2048, (nblocks, C)).astype("<i2")
    data[:, C:3 * C] = delta.view(np.uint8).reshape(nblocks, 2 * C)
    if tame:
        # nibbles drawn from {-2..2} mostly keep delta small (encoder-like streams)
        nib = rng.choice([0, 1, 2, 15, 14, 3, 13], p=[.3, .2, .1, .2, .1, .05, .05], size=(nblocks, (block_align - 7 * C) * 2))
        data[:, 7 * C:] = (nib[:, 0::2] << 4 | nib[:, 1::2]).astype(np.uint8)
    return data.reshape(-1)


def f32_equal_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


def ref_resample_window(window, in_first, n_total, src, dst, out_first, n_out, interp):
    """numpy float64 restatement of Audio:resample (A:653-673) for global output frames
    [out_first, out_first + n_out) when only input frames [in_first, in_first + window.shape[1])
    are held.  Pinned to the C oracle by tests/test_host_logic.py::test_numpy_window_reference."""
    w = np.asarray(window, dtype=np.float64)
    ratio = np.float64(dst) / np.float64(src)                      # A:658
    i0 = np.arange(out_first, out_first + n_out, dtype=np.float64)  # i - 1
    x = i0 / ratio + 1.0                                           # A:666
    fl = np.floor(x)
    hit = x == fl
    fx = x - fl
    f = fl.astype(np.int64)                                        # 1-based index of p1

    def tap(k):                                                    # data[f + k] with nil -> clamped index
        idx = np.clip(f + k, 1, n_total) - 1 - in_first
        return w[:, idx]

    p1 = tap(0)
    if interp == "none":
        v = p1
    elif interp == "linear":
        v = p1 + (tap(1) - p1) * fx
    else:
        p0, p2, p3 = tap(-1), tap(1), tap(2)
        v = (-0.5 * p0 + 1.5 * p1 - 1.5 * p2 + 0.5 * p3) * fx ** 3 + (p0 - 2.5 * p1 + 2 * p2 - 0.5 * p3) * fx ** 2 \
            + (-0.5 * p0 + 0.5 * p2) * fx + p1
    return np.where(hit, p1, np.clip(v, -1, 1))
