#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.npz by executing the UNMODIFIED reference
/root/reference/aukit.lua (inside oracle/luavm, a Lua 5.2 interpreter written for this repo --
no Lua binary exists in the image) on seeded inputs.

    python tests/golden/generate.py          # only works where /root/reference exists

The .npz holds, per case, the input bytes / arrays, the declarative description of the call
(`op` + arguments, JSON) and the reference's outputs as float64 (or its error message).
tests/test_golden_reference.py replays every case against the C oracle (bit-exact, CPU) and
against the CUDA path (GPU).  Nothing reads /root/reference at test time.
"""
from __future__ import annotations

import json
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.luavm.aukit_ref import Reference  # noqa: E402
from oracle.luavm.lua import LuaError, LuaTable, call, index, to_lua  # noqa: E402
from util import fmt_chunk, ima_blocks, ms_blocks, riff, tone_s16, wav_pcm  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz")


def lua_table(values):
    t = LuaTable()
    t.arr = [float(v) for v in values]
    return t


def audio_from_numpy(R, x, rate):
    """Builds an aukit.Audio from doubles by calling aukit.new and filling data (like a loader would)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    a = R.call("new", 0, x.shape[0], rate)[0]
    data = a.get(b"data")
    for c in range(x.shape[0]):
        data.arr[c].arr = [float(v) for v in x[c]]
    return a


def run_case(R, case, blobs):
    op = case["op"]
    A = case.get("args", {})
    if op == "pcm":
        a = R.call("pcm", blobs["in"], A["bitDepth"], A["dataType"], A["channels"], A["sampleRate"], A["interleaved"], A["bigEndian"])[0]
    elif op == "g711":
        a = R.call("g711", blobs["in"], A["ulaw"], A["channels"], A.get("sampleRate"))[0]
    elif op == "adpcm":
        args = [blobs["in"], A["channels"], A["sampleRate"], A["topFirst"], A["interleaved"]]
        args.append(lua_table(A["predictor"]) if A.get("predictor") is not None else None)
        args.append(lua_table(A["step_index"]) if A.get("step_index") is not None else None)
        a = call(R.fn("adpcm"), [to_lua(v) if not isinstance(v, LuaTable) else v for v in args])[0]
    elif op == "msadpcm":
        co = A.get("coefficients")
        cot = None
        if co is not None:
            cot = LuaTable()
            cot.arr = [lua_table(co[0]), lua_table(co[1])]
        a = call(R.fn("msadpcm"), [blobs["in"], float(A["blockAlign"]), float(A["channels"]), float(A["sampleRate"]), cot])[0]
    elif op == "wav":
        a = R.call("wav", blobs["in"], A.get("head", False))[0]
    elif op == "au":
        a = R.call("au", blobs["in"])[0]
    elif op == "aiff":
        a = R.call("aiff", blobs["in"], A.get("head", False))[0]
    elif op in ("invert", "fade", "delay", "center"):
        a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        extra = {"invert": [], "center": [],
                 "fade": [A.get("startTime"), A.get("startAmplitude"), A.get("endTime"), A.get("endAmplitude")],
                 "delay": [A.get("delay"), A.get("multiplier")]}[op]
        r = call(R.fn("effects", op), [a] + [to_lua(v) for v in extra])[0]
        assert r is a
    elif op == "wav_out":
        a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        if A.get("metadata"):
            for k, v in A["metadata"].items():
                a.get(b"metadata").set(k.encode(), to_lua(v))
        s = R.method(a, "wav", A.get("bitDepth"))[0]
        out = R.call("new", 0, 1, A["sampleRate"])[0]          # carrier: the file's bytes as numbers
        out.get(b"data").arr[0].arr = [float(b) for b in s]
        return out
    elif op == "pcm_out":
        a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        t = R.method(a, "pcm", A.get("bitDepth"), A.get("dataType"), A.get("interleaved"))[0]
        out = R.call("new", 0, 1, A["sampleRate"])[0]          # carrier: one "channel" holding the flat result
        out.get(b"data").arr[0].arr = list(t.arr)
        return out
    elif op == "chain3":                      # BASELINE config 3 per clip: aukit.pcm -> Audio:resample -> effects.amplify
        a = R.call("pcm", blobs["in"], A["bitDepth"], A["dataType"], A["channels"], A["sampleRate"], True, A["bigEndian"])[0]
        a = R.method(a, "resample", A["targetRate"], A.get("interpolation"))[0]
        r = R.call(("effects", "amplify"), a, A["multiplier"])[0]
        assert r is a
    elif op == "chain5":                      # BASELINE config 5: aukit.pcm (f32, 8 ch) -> Audio:resample -> effects.normalize
        a = R.call("pcm", blobs["in"], 32, "float", A["channels"], A["sampleRate"], True, False)[0]
        a = R.method(a, "resample", A["targetRate"], A.get("interpolation"))[0]
        r = call(R.fn("effects", "normalize"), [a, to_lua(A.get("peak")), to_lua(A.get("independent"))])[0]
        assert r is a
    elif op == "stream_adpcm":                # aukit.stream.adpcm (A:2798-2815): the N-channel IMA block layout's authority
        it = R.call(("stream", "adpcm"), blobs["in"], A["blockAlign"], A["channels"], A["sampleRate"], False)[0]
        chans = None
        while True:
            r = call(it, [])
            if not r or r[0] is None:
                break
            if chans is None:
                chans = [[] for _ in r[0].arr]
            for c, t in enumerate(r[0].arr):
                chans[c].extend(t.arr)
        out = R.call("new", 0, len(chans), 48000)[0]
        for c, v in enumerate(chans):
            out.get(b"data").arr[c].arr = [float(x) for x in v]
        return out
    elif op == "stream_out":                  # Audio:stream (A:921-937): chunks of encodePCM values + positions
        a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        r = R.method(a, "stream", A.get("chunkSize"), A.get("bitDepth"), A.get("dataType"))
        it, total = r[0], r[1]
        chans, poss, sizes = None, [], []
        while True:
            r = call(it, [])
            if not r or r[0] is None:
                break
            if chans is None:
                chans = [[] for _ in r[0].arr]
            for c, t in enumerate(r[0].arr):
                chans[c].extend(t.arr)
            poss.append(float(r[1]))
            sizes.append(float(len(r[0].arr[0].arr)))
        chans = chans or [[]]
        out = R.call("new", 0, len(chans) + 2, A["sampleRate"])[0]      # carrier: channels, then positions, then [total, sizes...]
        for c, v in enumerate(chans):
            out.get(b"data").arr[c].arr = [float(x) for x in v]
        out.get(b"data").arr[len(chans)].arr = poss
        out.get(b"data").arr[len(chans) + 1].arr = [float(total)] + sizes
        return out
    elif op in ("lowpass", "highpass"):
        a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        r = R.call(("effects", op), a, A["frequency"])[0]
        assert r is a                           # in place (A:3597)
    elif op in ("resample", "mono", "amplify", "normalize", "chain"):
        if op == "chain":
            a = R.call("wav", blobs["in"])[0]
        else:
            a = audio_from_numpy(R, blobs["x"], A["sampleRate"])
        if op in ("resample", "chain"):
            a = R.method(a, "resample", A["targetRate"], A.get("interpolation"))[0]
        if op in ("mono", "chain"):
            a = R.method(a, "mono")[0]
        if op == "amplify":
            r = R.call(("effects", "amplify"), a, A["multiplier"])[0]
            assert r is a                       # mutates and returns the argument (A:3368)
        if op in ("normalize", "chain"):
            r = call(R.fn("effects", "normalize"), [a, to_lua(A.get("peak")), to_lua(A.get("independent"))])[0]
            assert r is a
    else:
        raise ValueError(op)
    return a


def main():
    t0 = time.time()
    R = Reference()
    rng = np.random.default_rng(20260101)
    cases, store = [], {}

    def add(name, op, args=None, **blobs):
        cases.append({"name": name, "op": op, "args": args or {}})
        i = len(cases) - 1
        for k, v in blobs.items():
            store["c%d/%s" % (i, k)] = np.frombuffer(v, dtype=np.uint8).copy() if isinstance(v, (bytes, bytearray)) else np.asarray(v)

    # ---- aukit.pcm: every depth / type / endianness / layout
    for bits, dt in ((8, "signed"), (8, "unsigned"), (16, "signed"), (16, "unsigned"), (24, "signed"), (24, "unsigned"),
                     (32, "signed"), (32, "unsigned"), (32, "float")):
        for be in (False, True):
            for ch, il in ((1, True), (2, True), (3, False)):
                n = 96
                raw = rng.integers(0, 256, n * ch * bits // 8, dtype=np.uint8)
                raw[: bits // 8] = 0
                raw[bits // 8: 2 * (bits // 8)] = 255
                if dt == "float":
                    f = (rng.standard_normal(n * ch) * 0.7).astype(">f4" if be else "<f4")
                    f[:3] = [2.5, -0.0, 1e-30]
                    raw = f.view(np.uint8)
                add("pcm_%d%s_%s_c%d_%s" % (bits, dt[0], "be" if be else "le", ch, "il" if il else "pl"), "pcm",
                    dict(bitDepth=bits, dataType=dt, channels=ch, sampleRate=44100, interleaved=il, bigEndian=be), **{"in": raw.tobytes()})
    add("pcm_err_depth", "pcm", dict(bitDepth=12, dataType="signed", channels=1, sampleRate=48000, interleaved=True, bigEndian=False), **{"in": b"\0\0"})
    add("pcm_err_uneven", "pcm", dict(bitDepth=16, dataType="signed", channels=2, sampleRate=48000, interleaved=True, bigEndian=False), **{"in": b"\0" * 6})
    add("pcm_err_float16", "pcm", dict(bitDepth=16, dataType="float", channels=1, sampleRate=48000, interleaved=True, bigEndian=False), **{"in": b"\0" * 4})
    # exhaustive 16-bit signed scaling (A:1133)
    add("pcm_s16_all", "pcm", dict(bitDepth=16, dataType="signed", channels=1, sampleRate=48000, interleaved=True, bigEndian=False),
        **{"in": np.arange(-32768, 32768, 17, dtype="<i2").tobytes()})

    # ---- aukit.g711
    allb = bytes(range(256))
    for ulaw in (True, False):
        add("g711_%s_all" % ("u" if ulaw else "a"), "g711", dict(ulaw=ulaw, channels=1, sampleRate=None), **{"in": allb})
        add("g711_%s_ragged3" % ("u" if ulaw else "a"), "g711", dict(ulaw=ulaw, channels=3, sampleRate=8000), **{"in": allb[:200]})

    # ---- aukit.adpcm (headerless nibble strings)
    raw = rng.integers(0, 256, 150, dtype=np.uint8).tobytes()
    for ch, top, il, pr, si in ((1, True, True, None, None), (2, False, True, [100, -200], [5, 60]), (3, True, False, [0, 1, 2], [88, 0, 44])):
        add("adpcm_c%d_%s_%s" % (ch, "top" if top else "low", "il" if il else "pl"), "adpcm",
            dict(channels=ch, sampleRate=48000, topFirst=top, interleaved=il, predictor=pr, step_index=si), **{"in": raw})

    # ---- aukit.wav: IMA ADPCM (A:1509-1548), literal mono / stereo incl. the quirks
    for ch, ba, nb, cut in ((1, 64, 5, 7), (1, 37, 3, 0), (2, 72, 4, 0), (2, 256, 2, 0)):
        blocks = ima_blocks(nb, ba, ch, seed=ba + ch)
        if ch == 1:
            blocks[2] = 0x5B                            # header index 91 -> & 0x0F = 11 (A:1544)
        payload = blocks.tobytes()[: nb * ba - cut]
        add("wav_ima_c%d_ba%d" % (ch, ba), "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(0x11, ch, 22050, ba, 4)), (b"data", payload)])})
    bad = ima_blocks(2, 72, 2, seed=9)
    bad[72 + 6] = 120
    add("wav_ima_err_index", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(0x11, 2, 22050, 72, 4)), (b"data", bad.tobytes())])})
    add("wav_ima_err_3ch", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(0x11, 3, 22050, 96, 4)), (b"data", ima_blocks(2, 96, 3).tobytes())])})

    # ---- aukit.msadpcm (A:1283-1353) mono (first-header bug) / stereo, tame + wild nibbles, custom coefficients
    for ch, ba, nb, tame in ((1, 64, 4, True), (2, 64, 4, True), (2, 128, 3, False), (1, 256, 2, False)):
        add("msadpcm_c%d_ba%d_%s" % (ch, ba, "tame" if tame else "wild"), "msadpcm",
            dict(blockAlign=ba, channels=ch, sampleRate=44100, coefficients=None), **{"in": ms_blocks(nb, ba, ch, seed=ba + ch, tame=tame).tobytes()})
    coefs = [[256, 512, 0, 192, 240, 460, 392, -300], [0, -256, 0, 64, 0, -208, -232, 77]]
    blk = ms_blocks(3, 64, 2, seed=3)
    blk.reshape(3, 64)[:, :2] = 7
    extra = struct.pack("<HHH", 4 + 4 * 8, 100, 8) + b"".join(struct.pack("<hh", a, b) for a, b in zip(*coefs))
    add("wav_msadpcm_coefs", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(2, 2, 11025, 64, 4, extra)), (b"data", blk.tobytes())])})
    add("msadpcm_err_3ch", "msadpcm", dict(blockAlign=64, channels=3, sampleRate=44100, coefficients=None), **{"in": ms_blocks(1, 64, 3).tobytes()})

    # ---- aukit.wav containers: PCM / float / G.711 / extensible / INFO tags / errors
    for bits, fmt, ch in ((8, 1, 2), (16, 1, 2), (24, 1, 1), (32, 1, 2), (32, 3, 2), (8, 6, 2), (8, 7, 1)):
        payload = rng.integers(0, 256, 60 * ch * bits // 8, dtype=np.uint8)
        if fmt == 3:
            payload = rng.standard_normal(60 * ch).astype("<f4").view(np.uint8)
        add("wav_fmt%d_%dbit_c%d" % (fmt, bits, ch), "wav", {}, **{"in": wav_pcm(payload.tobytes(), ch, 32000, bits, fmt)})
    guid_tail = bytes.fromhex("000010008000" "00aa00389b71")
    ext = struct.pack("<HHI", 22, 16, 3) + bytes([1, 0, 0, 0]) + guid_tail
    add("wav_extensible_pcm", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(0xFFFE, 2, 48000, 4, 16, ext)),
                                                         (b"data", rng.integers(0, 256, 80, dtype=np.uint8).tobytes())])})
    tags = b"INFO" + b"INAM" + struct.pack("<I", 5) + b"Song\0" + b"\0" + b"ITRK" + struct.pack("<I", 2) + b"7\0" + b"IART" + struct.pack("<I", 2) + b"12"
    add("wav_info_tags", "wav", {}, **{"in": riff([(b"LIST", tags), (b"fmt ", fmt_chunk(1, 1, 8000, 2, 16)), (b"data", bytes(range(16)))])})
    add("wav_two_data_chunks", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(1, 1, 8000, 1, 8)), (b"data", b"\1\2\3"), (b"data", b"\4\5")])})
    add("wav_head_only", "wav", dict(head=True), **{"in": wav_pcm(bytes(40), 2, 32000, 16)})
    add("wav_err_riff", "wav", {}, **{"in": b"RIFX" + bytes(40)})
    add("wav_err_unsupported", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(85, 2, 44100, 4, 16)), (b"data", bytes(4))])})
    add("wav_err_nodata", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(1, 2, 44100, 4, 16))])})
    add("wav_err_truncated", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(1, 2, 44100, 4, 16)), (b"data", bytes(8))])[:-3]})

    # ---- Audio:resample (A:653-673): every mode x rate pair; a signal that exceeds [-1, 1] (exact hits unclamped)
    x = rng.uniform(-1, 1, (2, 400))
    x[:, :200] *= 0.3
    wild = rng.standard_normal((1, 300)) * 1.2
    for src, dst in ((44100, 48000), (48000, 44100), (22050, 48000), (96000, 48000), (8000, 48000), (44100, 44100), (96000, 44100), (3, 7)):
        for interp in ("none", "linear", "cubic"):
            add("resample_%d_%d_%s" % (src, dst, interp), "resample", dict(sampleRate=src, targetRate=dst, interpolation=interp), x=x)
            add("resample_wild_%d_%d_%s" % (src, dst, interp), "resample", dict(sampleRate=src, targetRate=dst, interpolation=interp), x=wild)
    add("resample_default_interp", "resample", dict(sampleRate=44100, targetRate=48000, interpolation=None), x=x[:1])
    add("resample_err_interp", "resample", dict(sampleRate=44100, targetRate=48000, interpolation="bogus"), x=x[:1])
    add("resample_one_frame", "resample", dict(sampleRate=8000, targetRate=48000, interpolation="cubic"), x=np.array([[0.5]]))
    # index quirk (SURVEY finding 5): a ramp through 'none' shows which sample floor(x) selects
    ramp = (np.arange(14700) % 4093) / 4096.0
    add("resample_index_quirk_none", "resample", dict(sampleRate=44100, targetRate=48000, interpolation="none"), x=ramp[None, :])

    # ---- Audio:mono, effects.amplify, effects.normalize
    y = rng.uniform(-0.9, 0.9, (3, 257))
    add("mono_3ch", "mono", dict(sampleRate=48000), x=y)
    add("mono_2ch", "mono", dict(sampleRate=48000), x=y[:2])
    add("amplify_1", "amplify", dict(sampleRate=48000, multiplier=1), x=y)
    add("amplify_2p5", "amplify", dict(sampleRate=48000, multiplier=2.5), x=y)
    add("normalize_default", "normalize", dict(sampleRate=48000, peak=None, independent=None), x=y * 0.5)
    add("normalize_0p8", "normalize", dict(sampleRate=48000, peak=0.8, independent=False), x=y * 0.5)
    add("normalize_independent", "normalize", dict(sampleRate=48000, peak=0.8, independent=True), x=y * np.array([[0.2], [0.5], [1.0]]))
    add("normalize_silence", "normalize", dict(sampleRate=48000, peak=0.8, independent=False), x=np.zeros((1, 16)))
    add("normalize_nan", "normalize", dict(sampleRate=48000, peak=1.0, independent=False), x=np.array([[np.nan, 0.5, -0.25]]))

    # ---- the auplay chain on a miniature of BASELINE config 1 (0.2 s): wav -> resample -> mono -> normalize(0.8)
    pcm = tone_s16(8820, 2, 44100, seed=1)
    for interp in ("linear", "cubic", "none"):
        add("chain_c1_mini_%s" % interp, "chain", dict(targetRate=48000, interpolation=interp, peak=0.8, independent=None),
            **{"in": wav_pcm(pcm.tobytes(), 2, 44100, 16)})

    # ---- effects.lowpass (appended after the round-1 cases so that their vectors keep their seeds)
    rng2 = np.random.default_rng(20260102)
    z = rng2.uniform(-1, 1, (2, 6000))
    add("lowpass_auplay_nyquist", "lowpass", dict(sampleRate=48000, frequency=24000.0), x=z)        # auplay.lua:30: sampleRate / 2
    add("lowpass_1khz", "lowpass", dict(sampleRate=48000, frequency=1000.0), x=z)
    add("lowpass_20hz_long_memory", "lowpass", dict(sampleRate=48000, frequency=20.0), x=z[:1, :5000] + 0.25)
    add("lowpass_single_sample", "lowpass", dict(sampleRate=8000, frequency=100.0), x=np.array([[0.5]]))
    add("lowpass_two_samples", "lowpass", dict(sampleRate=8000, frequency=100.0), x=np.array([[0.5, -0.5], [1.0, 0.0]]))
    add("lowpass_zero_hz", "lowpass", dict(sampleRate=44100, frequency=0.0), x=z[:, :64])           # a = 0: every sample becomes d[1]

    # ---- Audio:pcm / encodePCM: un-rounded values; inputs are f32-representable
    w = np.concatenate([np.array([-1.0, -0.5, 0.0, 0.5, 1.0, 0.999969482421875, -3.0517578125e-05, 1.5, -1.25]),
                        rng2.uniform(-1, 1, 291)]).astype(np.float32).astype(np.float64)
    w2 = np.stack([w, w[::-1]])
    for bits, dt in ((8, "signed"), (8, "unsigned"), (16, "signed"), (16, "unsigned"), (24, "signed"), (32, "signed"),
                     (32, "unsigned"), (32, "float")):
        add("pcmout_%d_%s_il" % (bits, dt), "pcm_out", dict(sampleRate=48000, bitDepth=bits, dataType=dt, interleaved=True), x=w2)
    add("pcmout_defaults", "pcm_out", dict(sampleRate=48000, bitDepth=None, dataType=None, interleaved=None), x=w2)
    add("pcmout_16_planar", "pcm_out", dict(sampleRate=48000, bitDepth=16, dataType="signed", interleaved=False), x=w2)
    add("pcmout_mono_24u", "pcm_out", dict(sampleRate=48000, bitDepth=24, dataType="unsigned", interleaved=True), x=w2[:1])
    add("pcmout_bad_depth", "pcm_out", dict(sampleRate=48000, bitDepth=12, dataType="signed", interleaved=True), x=w2)
    add("pcmout_bad_type", "pcm_out", dict(sampleRate=48000, bitDepth=16, dataType="int", interleaved=True), x=w2)
    add("pcmout_float16", "pcm_out", dict(sampleRate=48000, bitDepth=16, dataType="float", interleaved=True), x=w2)

    # ---- aukit.au / aukit.aiff containers
    def au_file(enc, ch, rate, payload, size=None, offset=24, extra=b""):
        return b".snd" + struct.pack(">IIIII", offset, len(payload) if size is None else size, enc, rate, ch) + extra + payload

    pay = rng2.integers(0, 256, 960, dtype=np.uint8).tobytes()
    fpay = rng2.standard_normal(240).astype(">f4").tobytes()
    for enc, name in ((1, "ulaw"), (2, "s8"), (3, "s16"), (5, "s32"), (27, "alaw")):
        add("au_%s_stereo" % name, "au", **{"in": au_file(enc, 2, 22050, pay)})
    add("au_s24_mono", "au", **{"in": au_file(4, 1, 8000, pay + b"\0\1\2")})
    add("au_f32_mono_unknown_size", "au", **{"in": au_file(6, 1, 48000, fpay + b"\0\0\0", size=0xFFFFFFFF)})
    add("au_s16_annotation", "au", **{"in": au_file(3, 2, 44100, pay, offset=32, extra=b"hello!!\0")})
    add("au_bad_magic", "au", **{"in": b".sndX"[1:] + bytes(24)})
    add("au_short", "au", **{"in": b".snd" + bytes(10)})
    add("au_bad_encoding", "au", **{"in": au_file(23, 1, 8000, pay)})

    def ext80(rate):
        m, e = np.frexp(float(rate))                     # rate = m * 2^e, 0.5 <= m < 1
        return struct.pack(">HQ", 0x3FFE + int(e), int(m * 2 ** 64))

    def chunk(tag, body):
        return tag + struct.pack(">I", len(body)) + body

    def aiff_file(ch, bits, rate, payload, comp=None, texts=(), frames=None, ssnd_off=0):
        frames = len(payload) // (ch * (bits // 8)) if frames is None else frames
        comm = struct.pack(">hIh", ch, frames, bits) + ext80(rate)
        if comp is not None:
            cname = b"not compressed"
            comm += comp + bytes([len(cname)]) + cname + (b"\0" if len(cname) % 2 == 0 else b"")
        body = (b"AIFC" if comp is not None else b"AIFF") + chunk(b"COMM", comm)
        for tag, text in texts:
            body += chunk(tag, text)
        body += chunk(b"SSND", struct.pack(">II", ssnd_off, 0) + bytes(ssnd_off) + payload)
        return b"FORM" + struct.pack(">I", len(body)) + body

    add("aiff_s16_stereo", "aiff", dict(head=False), **{"in": aiff_file(2, 16, 44100, pay)})
    add("aiff_s8_mono_meta", "aiff", dict(head=False), **{"in": aiff_file(1, 8, 22050, pay, texts=((b"NAME", b"Song"), (b"AUTH", b"Me"),
                                                                                                     (b"(c) ", b"2026"), (b"ANNO", b"note")))})
    add("aiff_s24_3ch", "aiff", dict(head=False), **{"in": aiff_file(3, 24, 48000, pay[:954])})
    add("aiff_s32_offset", "aiff", dict(head=False), **{"in": aiff_file(2, 32, 96000, pay, ssnd_off=4)})
    add("aifc_none", "aiff", dict(head=False), **{"in": aiff_file(2, 16, 32000, pay, comp=b"NONE")})
    add("aifc_sowt", "aiff", dict(head=False), **{"in": aiff_file(2, 16, 44100, pay, comp=b"sowt")})
    add("aifc_fl32", "aiff", dict(head=False), **{"in": aiff_file(1, 32, 48000, fpay, comp=b"fl32")})
    add("aifc_ulaw", "aiff", dict(head=False), **{"in": aiff_file(2, 8, 8000, pay, comp=b"ulaw")})
    add("aifc_alaw_upper", "aiff", dict(head=False), **{"in": aiff_file(1, 8, 8000, pay, comp=b"ALAW")})
    add("aiff_head_only", "aiff", dict(head=True), **{"in": aiff_file(2, 16, 44100, pay, texts=((b"NAME", b"T"),))})
    add("aiff_odd_rate", "aiff", dict(head=False), **{"in": aiff_file(1, 16, 11025.5, pay)})
    add("aiff_short_payload", "aiff", dict(head=False), **{"in": aiff_file(2, 16, 44100, pay, frames=300)})
    add("aifc_bad_compression", "aiff", dict(head=False), **{"in": aiff_file(2, 16, 44100, pay, comp=b"ima4")})
    add("aiff_not_aiff", "aiff", dict(head=False), **{"in": b"FORM" + bytes(4) + b"WAVE" + bytes(20)})
    add("aiff_no_ssnd", "aiff", dict(head=False), **{"in": b"FORM" + bytes(4) + b"AIFF" + chunk(b"COMM", struct.pack(">hIh", 1, 0, 8) + ext80(8000))})

    # ---- interpolate.sinc: window +-10, taps outside the signal skipped
    zs = rng2.uniform(-1, 1, (2, 700))
    for src, dst in ((44100, 48000), (48000, 44100), (8000, 48000), (96000, 48000), (48000, 48000)):
        add("resample_sinc_%d_%d" % (src, dst), "resample", dict(sampleRate=src, targetRate=dst, interpolation="sinc"), x=zs)
    add("resample_sinc_short", "resample", dict(sampleRate=22050, targetRate=48000, interpolation="sinc"), x=zs[:1, :7])
    add("resample_sinc_loud", "resample", dict(sampleRate=44100, targetRate=48000, interpolation="sinc"), x=zs[:1, :300] * 1.4)

    # ---- effects.invert / fade / delay / center
    ze = rng2.uniform(-1, 1, (2, 2500)) + 0.2
    add("invert_2ch", "invert", dict(sampleRate=1000), x=ze)
    add("fade_out", "fade", dict(sampleRate=1000, startTime=0.5, startAmplitude=1.0, endTime=2.0, endAmplitude=0.0), x=ze)
    add("fade_in_boost", "fade", dict(sampleRate=1000, startTime=0.001, startAmplitude=0.0, endTime=1.2345, endAmplitude=1.7), x=ze)
    add("fade_noop", "fade", dict(sampleRate=1000, startTime=0.5, startAmplitude=1, endTime=2.0, endAmplitude=1), x=ze)
    add("fade_fractional_start", "fade", dict(sampleRate=1000, startTime=0.0005, startAmplitude=0.5, endTime=1.0, endAmplitude=1.0), x=ze)
    add("fade_from_zero_time", "fade", dict(sampleRate=1000, startTime=0.0, startAmplitude=0.5, endTime=1.0, endAmplitude=1.0), x=ze)
    add("fade_past_end", "fade", dict(sampleRate=1000, startTime=1.0, startAmplitude=0.5, endTime=3.0, endAmplitude=1.0), x=ze)
    add("delay_default", "delay", dict(sampleRate=1000, delay=0.3, multiplier=None), x=ze)
    add("delay_loud", "delay", dict(sampleRate=1000, delay=0.0105, multiplier=1.5), x=ze)
    add("delay_longer_than_audio", "delay", dict(sampleRate=1000, delay=5.0, multiplier=0.5), x=ze)
    add("delay_negative", "delay", dict(sampleRate=1000, delay=-0.1, multiplier=0.5), x=ze)
    add("center_blocks", "center", dict(sampleRate=1000), x=ze)
    add("center_one_block", "center", dict(sampleRate=48000), x=ze[:, :300])
    add("center_fractional_rate", "center", dict(sampleRate=1000.5), x=ze)

    # ---- effects.highpass (same scan as lowpass)
    add("highpass_200hz", "highpass", dict(sampleRate=48000, frequency=200.0), x=ze + 0.3)
    add("highpass_10khz", "highpass", dict(sampleRate=44100, frequency=10000.0), x=ze)
    add("highpass_zero_hz", "highpass", dict(sampleRate=8000, frequency=0.0), x=ze[:1, :100])
    add("highpass_two_samples", "highpass", dict(sampleRate=8000, frequency=100.0), x=np.array([[0.5, -0.5]]))

    # ---- Audio:wav writer (8f rank 2, second half); luavm's string.pack floors non-integral numbers.  Sizes are
    # chosen around the writer's 32768-value chunks (A:981-985): below one chunk it raises, above it the tail is shifted
    # in range only: what string.pack does with a value that does not fit is the host's business (luavm wraps)
    wl = np.clip(np.tile(w, 130)[:, None].reshape(1, -1)[:, :39000], -1, 1).astype(np.float32).astype(np.float64)    # 39000 values mono
    wl2 = np.stack([wl[0, :35000], wl[0, 100:35100]])                                               # 70000 values stereo
    for bits in (8, 16, 24, 32):
        add("wavout_%d" % bits, "wav_out", dict(sampleRate=22050, bitDepth=bits), x=wl)
    add("wavout_default_stereo", "wav_out", dict(sampleRate=8000, bitDepth=None), x=wl2)
    add("wavout_exact_two_chunks", "wav_out", dict(sampleRate=8000, bitDepth=16), x=wl2[:, :32768])
    add("wavout_metadata", "wav_out", dict(sampleRate=44100, bitDepth=16, metadata={"title": "A song"}), x=wl)
    add("wavout_short_raises", "wav_out", dict(sampleRate=44100, bitDepth=16), x=w2)
    add("wavout_bad_depth", "wav_out", dict(sampleRate=44100, bitDepth=12), x=w2[:, :50])

    # ==== round 2 (appended: earlier vectors keep their seeds) ====
    rng3 = np.random.default_rng(20260103)
    # aukit.wav decodes each data chunk with the fmt state seen so far (A:1505-1555)
    d16 = rng3.integers(0, 256, 64, dtype=np.uint8).tobytes()
    add("wav_fmt_after_data", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(1, 2, 32000, 4, 16)), (b"data", d16), (b"fmt ", fmt_chunk(1, 1, 8000, 1, 8))])})
    add("wav_data_before_fmt", "wav", {}, **{"in": riff([(b"data", d16), (b"fmt ", fmt_chunk(1, 2, 32000, 4, 16))])})
    extra0 = struct.pack("<HHH", 4, 100, 0)
    add("wav_msadpcm_coefs_persist", "wav", {}, **{"in": riff([(b"fmt ", fmt_chunk(2, 2, 11025, 64, 4, extra)), (b"fmt ", fmt_chunk(2, 2, 11025, 64, 4, extra0)),
                                                                (b"data", blk.tobytes())])})
    # effects.normalize with a negative / large peak (the final clamp acts; A:3455)
    add("normalize_neg2", "normalize", dict(sampleRate=48000, peak=-2.0, independent=False), x=y * 0.5)
    add("normalize_1p5", "normalize", dict(sampleRate=48000, peak=1.5, independent=False), x=y * 0.5)
    # non-finite input through the one-pole filters: the state never recovers (A:3592-3595, A:3613-3615)
    zl = rng3.uniform(-1, 1, (2, 30000))
    zl[0, 12345] = np.nan
    zl[1, 20000] = np.inf
    add("lowpass_nonfinite", "lowpass", dict(sampleRate=48000, frequency=12000.0), x=zl)
    add("highpass_nonfinite", "highpass", dict(sampleRate=48000, frequency=200.0), x=zl)
    # BASELINE config 3 per clip: s24 big-endian stereo -> 48 kHz cubic -> amplify(0.5); full-scale noise, 0.05 s
    for src in (22050, 44100, 96000):
        n3 = src // 20
        raw3 = rng3.integers(0, 256, n3 * 2 * 3, dtype=np.uint8)
        add("chain3_s24be_%d" % src, "chain3", dict(bitDepth=24, dataType="signed", channels=2, sampleRate=src, bigEndian=True,
                                                    targetRate=48000, interpolation="cubic", multiplier=0.5), **{"in": raw3.tobytes()})
    add("chain3_s24be_44100_linear_boost", "chain3", dict(bitDepth=24, dataType="signed", channels=2, sampleRate=44100, bigEndian=True,
                                                          targetRate=48000, interpolation="linear", multiplier=1.7), **{"in": raw3[: 2205 * 6].tobytes()})
    # BASELINE config 5 / 5': f32 8-channel interleaved -> 48 kHz / 44.1 kHz cubic -> normalize
    f5 = (rng3.standard_normal(1200 * 8) * 0.25).astype("<f4")
    for dst in (48000, 44100):
        add("chain5_f32x8_96000_%d" % dst, "chain5", dict(channels=8, sampleRate=96000, targetRate=dst, interpolation="cubic", peak=None, independent=None),
            **{"in": f5.tobytes()})
    # aukit.stream.adpcm at 48 kHz (ratio 1: every position is an exact hit): pins the N-channel IMA block layout and step
    for ch, ba, nb in ((1, 36, 3), (2, 72, 3), (3, 108, 2), (8, 288, 2)):
        add("stream_adpcm_c%d" % ch, "stream_adpcm", dict(blockAlign=ba, channels=ch, sampleRate=48000), **{"in": ima_blocks(nb, ba, ch, seed=40 + ch).tobytes()})
    # Audio:stream (A:921-937)
    xs = rng3.uniform(-1, 1, (2, 1000)).astype(np.float32).astype(np.float64)
    add("audiostream_default_depth", "stream_out", dict(sampleRate=48000, chunkSize=300, bitDepth=None, dataType=None), x=xs)
    add("audiostream_s16", "stream_out", dict(sampleRate=44100, chunkSize=256, bitDepth=16, dataType="signed"), x=xs)
    add("audiostream_u8_exact_chunks", "stream_out", dict(sampleRate=8000, chunkSize=250, bitDepth=8, dataType="unsigned"), x=xs)
    add("audiostream_default_chunk", "stream_out", dict(sampleRate=48000, chunkSize=None, bitDepth=None, dataType=None), x=xs[:1])
    add("audiostream_bad_depth", "stream_out", dict(sampleRate=48000, chunkSize=100, bitDepth=12, dataType=None), x=xs[:1])

    # ---- run everything through the reference
    manifest = []
    for i, case in enumerate(cases):
        blobs = {}
        for k in ("in", "x"):
            key = "c%d/%s" % (i, k)
            if key in store:
                blobs[k] = store[key].tobytes() if k == "in" else store[key]
        entry = dict(case)
        t1 = time.time()
        try:
            a = run_case(R, case, blobs)
            chans = Reference.audio_data(a)
            fields = Reference.audio_fields(a)
            entry["channels"] = len(chans)
            entry["lengths"] = [int(c.size) for c in chans]
            entry["sampleRate"] = fields["sampleRate"]
            entry["metadata"] = {k: (v if not isinstance(v, float) or v == v else None) for k, v in fields["metadata"].items()}
            entry["info"] = fields["info"]
            for c, arr in enumerate(chans):
                store["c%d/out%d" % (i, c)] = arr
        except LuaError as ex:
            entry["error"] = str(ex)
        entry["seconds"] = round(time.time() - t1, 3)
        manifest.append(entry)
        print("%-40s %s" % (case["name"], entry.get("error") or entry.get("lengths")), flush=True)
    store["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **store)
    print("wrote %s: %d cases, %.1f KB, %.1f s" % (OUT, len(cases), os.path.getsize(OUT) / 1024, time.time() - t0))


if __name__ == "__main__":
    main()
