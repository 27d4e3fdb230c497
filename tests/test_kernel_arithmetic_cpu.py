"""CPU restatements of the integer / float tricks the round-2 ADPCM kernels rely on (csrc/adpcm.cu), checked exhaustively or
on large random samples against the reference's own arithmetic (A:1321-1324, A:1133) -- so that the exactness arguments in
the kernel comments do not rest on the GPU tests alone."""
from fractions import Fraction

import numpy as np


def _f32(fr: Fraction) -> np.float32:
    """Correctly rounded float32 of an exact rational (what one FMA / one division returns)."""
    c = np.float32(float(fr))                      # float(Fraction) is correctly rounded to double; fix a double rounding below
    cands = {c, np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))}
    best = min(cands, key=lambda v: (abs(Fraction(float(v)) - fr), int(np.float32(v).view(np.uint32)) & 1))
    return np.float32(best)


def test_ms_output_conversion_from_offset_binary_is_the_reference_division():
    """u = p + 32768; x = float bits (u + 0x4B000000) = 2^23 + u; lo = fma(x, 2^-15, -257); out = fma(sat(lo), 1/32767', lo)
    must equal (float)(p / (p < 0 and 32768 or 32767)) for all 65536 values (A:1133 / s16_to_float)."""
    K = np.float32((1.0 / (32767.0 * 32768.0)) * 32768.0)
    for p in range(-32768, 32768):
        u = p + 32768
        x = np.array([u + 0x4B000000], dtype=np.uint32).view(np.float32)[0]
        assert Fraction(float(x)) == Fraction(2 ** 23 + u)
        lo = _f32(Fraction(float(x)) * Fraction(1, 32768) - 257)
        assert Fraction(float(lo)) == Fraction(p, 32768)                      # exact
        sat = min(max(lo, np.float32(0.0)), np.float32(1.0))
        out = _f32(Fraction(float(sat)) * Fraction(float(K)) + Fraction(float(lo)))
        want = np.float32(p / (32768.0 if p < 0 else 32767.0))
        assert out.view(np.uint32) == want.view(np.uint32), p


def test_ms_fast_step_in_offset_binary_with_wrapping_int32():
    """The staged kernel's step -- t = u1*c1 + (u2*c2 + kc) in wrapping 32-bit arithmetic, u = relu(min((t >> 8) + nib*delta,
    65535)), delta = max((adapt*delta) >> 8, 16) -- against the reference's floor / clamp on (s1, s2) = (u1, u2) - 32768."""
    rng = np.random.default_rng(7)
    adapt = [230, 230, 230, 230, 307, 409, 512, 614, 768, 614, 512, 409, 307, 230, 230, 230]
    coefs = [(256, 0), (512, -256), (0, 0), (192, 64), (240, 0), (460, -208), (392, -232), (8000, -8384), (-16384, 0)]

    def wrap(v):
        return ((v + 2 ** 31) % 2 ** 32) - 2 ** 31

    for c1, c2 in coefs:
        assert abs(c1) + abs(c2) <= 16384                                      # the kernel's `narrow` condition
        kc = wrap((1 << 23) - 32768 * (c1 + c2))
        for _ in range(4000):
            s1, s2 = int(rng.integers(-32768, 32768)), int(rng.integers(-32768, 32768))
            delta = int(rng.integers(-65534, 16384))                           # the range the kernel calls `quick`
            for _k in range(4):                                                # a quad: delta may grow to 81 x 2^14 inside it
                nib = int(rng.integers(-8, 8))
                u1, u2 = s1 + 32768, s2 + 32768
                t = wrap(wrap(u1 * c1) + wrap(wrap(u2 * c2) + kc))
                u = max(min((t >> 8) + nib * delta, 65535), 0)
                ref = max(min((s1 * c1 + s2 * c2) // 256 + nib * delta, 32767), -32768)   # A:1321-1322
                assert u - 32768 == ref, (c1, c2, s1, s2, nib, delta)
                assert abs(nib * delta) < 2 ** 31 and abs(adapt[nib & 15] * delta) < 2 ** 31
                nd = max((adapt[nib & 15] * delta) >> 8, 16)                   # A:1324 (floor division, then math.max)
                assert nd == max((adapt[nib & 15] * delta) // 256, 16)
                s2, s1, delta = s1, ref, nd


def test_aligned_segment_schedule_visits_every_quad_once():
    """Iteration i of a lane produces the quads of ITS row that fall into the row's i-th 256-byte aligned segment: local
    quads j = 16 i - Q0 + q.  Every quad exactly once, interior iterations full for every phase, flush addresses aligned."""
    for nquads in (1, 3, 16, 17, 61, 509, 512):
        M = nquads // 16
        niter = (nquads + 15 + 15) // 16
        for Q0 in range(16):
            row_addr = 4096 * 7 + 16 * Q0                                      # any address with this phase
            seen = []
            for i in range(niter):
                j0 = 16 * i - Q0
                if 1 <= i < M:                                                 # the straight-line iterations
                    assert j0 >= 1 and j0 + 15 < nquads                        # never the header quad, never past the block
                    assert (row_addr + 16 * j0) % 256 == 0                     # whole aligned segment
                    seen += list(range(j0, j0 + 16))
                else:
                    qa, qb = max(0, -j0), min(16, nquads - j0)
                    seen += list(range(j0 + qa, j0 + max(qa, qb)))
            assert seen == list(range(nquads)), (nquads, Q0)
            # records a fast iteration reads: 16 consecutive ones starting at (16 j0) mod 1024 -- inside ring + mirror
            for i in range(1, M):
                start = (16 * (16 * i - Q0)) % 1024
                assert start + 16 * 16 <= 1024 + 256


def test_ima_aligned_word_schedule_visits_every_word_once():
    """ima_wav_aligned_kernel: iteration i decodes words 8 i - G0 + s (s = 0..7) into quad slots 2 s, 2 s + 1 of the row's
    i-th 256-byte aligned segment; the prefetch of a fast iteration never reads past the block."""
    for groups in (1, 2, 7, 8, 9, 31, 255, 256):
        M = groups // 8
        niter = (groups + 7 + 7) // 8
        for G0 in range(8):
            row_addr = 8192 * 3 + 32 * G0
            seen = []
            for i in range(niter):
                w0 = 8 * i - G0
                if 1 <= i < M:
                    assert w0 >= 1 and w0 + 7 < groups
                    assert (row_addr + 32 * w0) % 256 == 0
                    if i + 1 < M:                                              # words prefetched for the next fast iteration
                        assert w0 + 8 + 7 < groups
                    seen += list(range(w0, w0 + 8))
                else:
                    sa, sb = max(0, -w0), min(8, groups - w0)
                    seen += list(range(w0 + sa, w0 + max(sa, sb)))
            assert seen == list(range(groups)), (groups, G0)
            if M > 1:                                                          # words loaded before the loop for iteration 1
                assert 8 - G0 + 7 < groups


def test_blocked_lowpass_chunk_start_state_is_exact_to_fp64():
    """csrc/lowpass.cu, BLOCKED variant: where (1-a)^8192 < 8.3e-25 (2^-80) the state entering a chunk is taken to be the
    zero-start end state of the ONE tile before it.  Restated serially in fp64: against the full recurrence (A:3592-3595,
    A:3613-3615) the chunk-start state differs by at most (1-a)^8192 * |state|, far below an fp64 ulp of a sample-sized
    value, at the lowest cut-off `lp_run` lets through (and its threshold is where this test says it is)."""
    import math
    T = 8192
    rng = np.random.default_rng(11)
    x = (rng.uniform(-1, 1, 3 * T) + 0.3).astype(np.float32).astype(np.float64)

    def low(a, xs, y):
        for v in xs:
            y = y + a * (v - y)
        return y

    def high(a, xs, y, xp):
        for v in xs:
            y = a * ((y + v) - xp)
            xp = v
        return y

    for rate in (48000.0, 44100.0, 8000.0):
        # the lowest low-pass cut-off that passes the host test: (1 - a)^T < 8.3e-25 with a = 1 - exp(-2 pi f / rate)
        f_min = -math.log(8.3e-25) / T * rate / (2 * math.pi)
        assert 40.0 < f_min * 48000.0 / rate < 53.0
        a = 1.0 - math.exp(-(f_min * 1.001 / rate) * 2 * math.pi)
        assert (1.0 - a) ** T < 8.3e-25
        full = low(a, x[: 2 * T], x[0])                       # true state after two tiles (d[1] untouched: state = d[1])
        trunc = low(a, x[T: 2 * T], 0.0)                      # what lp_chunk_states computes for a chunk starting at tile 2
        assert abs(full - trunc) <= 8.3e-25 * 1.5 and abs(full) > 1e-3
        # high-pass: ratio a = 1 / (2 pi f / rate + 1); the same bound with a^T
        ah = 8.3e-25 ** (1.0 / T) * 0.9999
        assert ah ** T < 8.3e-25
        full = high(ah, x[1: 2 * T], x[0], x[0])
        trunc = high(ah, x[T: 2 * T], 0.0, x[T - 1])
        assert abs(full - trunc) <= 8.3e-25 * 4.0
    # lower cut-offs: the pre-pass runs from zero over `warm` tiles, the fewest with ratio^(T warm) < 8.3e-25 (lp_run)
    for f, warm in ((30.0, 2), (12.0, 5)):
        a = 1.0 - math.exp(-(f / 48000.0) * 2 * math.pi)
        assert (1.0 - a) ** (T * warm) < 8.3e-25 <= (1.0 - a) ** (T * (warm - 1))
        xs = (rng.uniform(-1, 1, (warm + 1) * T) + 0.3).astype(np.float32).astype(np.float64)
        full = low(a, xs, xs[0])
        trunc = low(a, xs[T:], 0.0)
        assert abs(full - trunc) <= 8.3e-25 * 1.5 and abs(full) > 1e-3
