"""ctypes front end for the parity oracle (oracle/libaukit_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the aukit_b200 package.  Every function
mirrors one reference function of /root/reference/aukit.lua (see aukit_oracle.h for the
file:line citations) and returns float64 numpy arrays shaped [channels, frames].
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaukit_oracle.so")

SIGNED, UNSIGNED, FLOAT = 0, 1, 2
NONE, LINEAR, CUBIC = 0, 1, 2
LITERAL, GENERAL = 0, 1
DATATYPES = {"signed": SIGNED, "unsigned": UNSIGNED, "float": FLOAT}
INTERPS = {"none": NONE, "linear": LINEAR, "cubic": CUBIC, "sinc": 3}


class OracleError(RuntimeError):
    """The reference would have raised a Lua error here; .args[0] is its message."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "aukit_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"] if force else ["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.auko_last_error.restype = C.c_char_p
        sz, u8p, dp, ip = C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p
        L.auko_pcm.argtypes = [u8p, sz, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, dp, sz, C.POINTER(sz)]
        L.auko_g711.argtypes = [u8p, sz, C.c_int, C.c_int, dp, sz, C.c_void_p]
        L.auko_ima_step.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.auko_ima_step.restype = C.c_double
        L.auko_adpcm.argtypes = [u8p, sz, C.c_int, C.c_int, C.c_int, ip, ip, dp, sz, C.POINTER(sz)]
        L.auko_wav_ima.argtypes = [u8p, sz, C.c_int, C.c_int, C.c_int, dp, sz, C.POINTER(sz)]
        L.auko_wav_ima_len.argtypes = [sz, C.c_int, C.c_int, C.c_int]
        L.auko_wav_ima_len.restype = sz
        L.auko_msadpcm.argtypes = [u8p, sz, C.c_int, C.c_int, ip, ip, C.c_int, C.c_int, dp, sz, C.POINTER(sz)]
        L.auko_msadpcm_len.argtypes = [sz, C.c_int, C.c_int]
        L.auko_msadpcm_len.restype = sz
        L.auko_resample_len.argtypes = [sz, C.c_double, C.c_double]
        L.auko_resample_len.restype = sz
        L.auko_resample_pos.argtypes = [C.c_uint64, C.c_double, C.c_double]
        L.auko_resample_pos.restype = C.c_double
        L.auko_resample.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_double, C.c_int, dp, sz, C.POINTER(sz)]
        L.auko_mono.argtypes = [dp, sz, C.c_int, sz, dp]
        L.auko_amplify.argtypes = [dp, sz, C.c_int, sz, C.c_double]
        L.auko_normalize.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_int]
        L.auko_encode_pcm.argtypes = [C.c_double, C.c_int, C.c_int]
        L.auko_encode_pcm.restype = C.c_double
        L.auko_audio_pcm.argtypes = [dp, sz, C.c_int, sz, C.c_int, C.c_int, C.c_int, dp]
        L.auko_au_parse.argtypes = [u8p, sz, C.c_void_p]
        L.auko_aiff_parse.argtypes = [u8p, sz, C.c_void_p]
        L.auko_invert.argtypes = [dp, sz, C.c_int, sz]
        L.auko_fade.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        L.auko_delay.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_double, C.c_double]
        L.auko_center.argtypes = [dp, sz, C.c_int, sz, C.c_double]
        L.auko_highpass.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_double]
        L.auko_lowpass.argtypes = [dp, sz, C.c_int, sz, C.c_double, C.c_double]
        L.auko_wav_parse.argtypes = [u8p, sz, C.c_void_p]
        L.auko_chain_s16.argtypes = [u8p, sz, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.POINTER(sz)]
        L.auko_chain_s16.restype = C.c_void_p
        L.auko_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise OracleError(lib().auko_last_error().decode("latin-1"))


def _bytes(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data.view(np.uint8).reshape(-1))
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pcm(data, bitDepth=8, dataType="signed", channels=1, interleaved=True, bigEndian=False):
    b = _bytes(data)
    dt = DATATYPES.get(dataType, 99) if isinstance(dataType, str) else dataType
    stride = max(1, b.size // max(1, bitDepth // 8 if bitDepth >= 8 else 1) // max(1, channels))
    out = np.zeros((max(1, channels), stride), dtype=np.float64)
    n = C.c_size_t(0)
    _check(lib().auko_pcm(_ptr(b), b.size, bitDepth, dt, channels, int(interleaved), int(bigEndian),
                          _ptr(out), stride, C.byref(n)))
    return out[:channels, : n.value]


def g711(data, ulaw: bool, channels=1):
    """Returns a list of per-channel arrays (ragged when len(data) % channels != 0)."""
    b = _bytes(data)
    stride = max(1, -(-b.size // max(1, channels)))
    out = np.zeros((max(1, channels), stride), dtype=np.float64)
    lens = (C.c_size_t * max(1, channels))()
    _check(lib().auko_g711(_ptr(b), b.size, int(ulaw), channels, _ptr(out), stride, lens))
    return [out[c, : lens[c]].copy() for c in range(channels)]


def ima_step(nibble: int, pred: int, idx: int):
    p, i = C.c_int(pred), C.c_int(idx)
    v = lib().auko_ima_step(nibble, C.byref(p), C.byref(i))
    return v, p.value, i.value


def adpcm(data, channels=1, topFirst=True, interleaved=True, predictor=None, step_index=None):
    b = _bytes(data)
    stride = max(1, b.size * 2 // max(1, channels))
    out = np.zeros((max(1, channels), stride), dtype=np.float64)
    n = C.c_size_t(0)
    pa = np.asarray(predictor, dtype=np.int32) if predictor is not None else None
    ia = np.asarray(step_index, dtype=np.int32) if step_index is not None else None
    _check(lib().auko_adpcm(_ptr(b), b.size, channels, int(topFirst), int(interleaved),
                            _ptr(pa) if pa is not None else None, _ptr(ia) if ia is not None else None,
                            _ptr(out), stride, C.byref(n)))
    return out[:channels, : n.value]


def wav_ima(data, blockAlign, channels, dialect=LITERAL):
    b = _bytes(data)
    stride = max(8, lib().auko_wav_ima_len(b.size, blockAlign, channels, dialect))
    out = np.zeros((max(1, channels), stride), dtype=np.float64)
    n = C.c_size_t(0)
    _check(lib().auko_wav_ima(_ptr(b), b.size, blockAlign, channels, dialect, _ptr(out), stride, C.byref(n)))
    return out[:channels, : n.value]


def msadpcm(data, blockAlign, channels=1, coefficients=None, dialect=LITERAL):
    b = _bytes(data)
    stride = max(8, lib().auko_msadpcm_len(b.size, blockAlign, channels))
    out = np.zeros((max(1, channels), stride), dtype=np.float64)
    n = C.c_size_t(0)
    if coefficients is not None:
        c1 = np.asarray(coefficients[0], dtype=np.int32)
        c2 = np.asarray(coefficients[1], dtype=np.int32)
        args = (_ptr(c1), _ptr(c2), int(c1.size))
    else:
        args = (None, None, 0)
    _check(lib().auko_msadpcm(_ptr(b), b.size, blockAlign, channels, *args, dialect, _ptr(out), stride, C.byref(n)))
    return out[:channels, : n.value]


def resample_len(n_in, srcRate, dstRate) -> int:
    return int(lib().auko_resample_len(n_in, float(srcRate), float(dstRate)))


def resample_pos(i, srcRate, dstRate) -> float:
    return float(lib().auko_resample_pos(int(i), float(srcRate), float(dstRate)))


def resample(x: np.ndarray, srcRate, dstRate, interpolation="linear"):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    ch, n = x.shape
    mode = INTERPS.get(interpolation, 99) if isinstance(interpolation, str) else interpolation
    nl = resample_len(n, srcRate, dstRate)
    out = np.zeros((ch, max(1, nl)), dtype=np.float64)
    m = C.c_size_t(0)
    _check(lib().auko_resample(_ptr(x), n, ch, n, float(srcRate), float(dstRate), mode, _ptr(out),
                               max(1, nl), C.byref(m)))
    return out[:, : m.value]


def mono(x: np.ndarray):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    ch, n = x.shape
    out = np.zeros((1, n), dtype=np.float64)
    _check(lib().auko_mono(_ptr(x), n, ch, n, _ptr(out)))
    return out


def amplify(x: np.ndarray, multiplier):
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    ch, n = x.shape
    _check(lib().auko_amplify(_ptr(x), n, ch, n, float(multiplier)))
    return x


def normalize(x: np.ndarray, peak=1.0, independent=False):
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    ch, n = x.shape
    _check(lib().auko_normalize(_ptr(x), n, ch, n, float(peak), int(bool(independent))))
    return x


def lowpass(x: np.ndarray, frequency, sampleRate):
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    ch, n = x.shape
    _check(lib().auko_lowpass(_ptr(x), n, ch, n, float(frequency), float(sampleRate)))
    return x


def _inplace(fn, x, *args):
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    ch, n = x.shape
    _check(fn(_ptr(x), n, ch, n, *[float(a) for a in args]))
    return x


def invert(x):
    return _inplace(lib().auko_invert, x)


def fade(x, sampleRate, startTime, startAmplitude, endTime, endAmplitude):
    return _inplace(lib().auko_fade, x, sampleRate, startTime, startAmplitude, endTime, endAmplitude)


def delay(x, sampleRate, delay, multiplier=0.5):
    return _inplace(lib().auko_delay, x, sampleRate, delay, multiplier)


def center(x, sampleRate):
    return _inplace(lib().auko_center, x, sampleRate)


def highpass(x, frequency, sampleRate):
    return _inplace(lib().auko_highpass, x, frequency, sampleRate)


def audio_pcm(x: np.ndarray, bitDepth=8, dataType="signed", interleaved=True) -> np.ndarray:
    """Audio:pcm (A:901): flat float64 array of un-rounded encoded values."""
    x = np.array(np.atleast_2d(x), dtype=np.float64, order="C")
    ch, n = x.shape
    out = np.zeros(max(1, ch * n), dtype=np.float64)
    dt = DATATYPES.get(dataType, 99) if isinstance(dataType, str) else dataType
    _check(lib().auko_audio_pcm(_ptr(x), n, ch, n, int(bitDepth), dt, int(bool(interleaved)), _ptr(out)))
    return out[: ch * n]


def audio_stream(x: np.ndarray, sampleRate, chunkSize=None, bitDepth=None, dataType=None):
    """Audio:stream (A:921-937): list of (per-channel value arrays, position in seconds) and the total length.
    Each step is encodePCM with `multiple` set (A:881-891) over frames [pos, pos + chunkSize), pos 1-based."""
    chunkSize = 131072 if chunkSize is None else int(chunkSize)
    bitDepth = 8 if bitDepth is None else bitDepth
    dataType = "signed" if dataType is None else dataType
    if bitDepth not in (8, 16, 24, 32):
        raise OracleError("bad argument #2 (invalid bit depth)")
    if dataType not in DATATYPES:
        raise OracleError("bad argument #3 (invalid data type)")
    if dataType == "float" and bitDepth != 32:
        raise OracleError("bad argument #2 (float audio must have 32-bit depth)")
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    n = x.shape[1]
    vals = audio_pcm(x, bitDepth, dataType, False).reshape(x.shape[0], n)
    steps, pos = [], 1
    while pos <= n:                                     # A:878: pos > len ends the iteration
        steps.append(([vals[c, pos - 1: pos - 1 + chunkSize].copy() for c in range(x.shape[0])], pos / sampleRate))
        pos += chunkSize
    return steps, n / sampleRate


def stream_adpcm_48k(data, blockAlign, channels):
    """aukit.stream.adpcm (A:2738-2834) at sampleRate = 48000, mono = false: the ratio is 1, every position is an
    exact hit, so each output is clamp(floor(p / (p < 0 and 128 or 127)), -128, 127) of the decoded predictor p
    (A:2810, A:2826).  The block layout is the N-channel one of A:2798-2815, i.e. this oracle's GENERAL dialect; the
    reference stops 8 samples early in the LAST block (its bounds check `#data < n + i + channels*4`, A:2801)."""
    v = wav_ima(data, blockAlign, channels, GENERAL)
    p = np.where(v < 0, np.rint(v * 32768.0), np.rint(v * 32767.0))          # predictors back from p / 32768|32767
    q = np.floor(np.where(p < 0, p / 128.0, p / 127.0))
    return np.clip(q, -128, 127)[:, : v.shape[1] - 8]


WAV_METADATA = {"IPRD": "album", "INAM": "title", "IART": "artist", "IWRI": "author", "IMUS": "composer", "IPRO": "producer",
                "IPRT": "trackNumber", "ITRK": "trackNumber", "IFRM": "trackCount", "PRT1": "partNumber", "PRT2": "partCount",
                "TLEN": "length", "IRTD": "rating", "ICRD": "date", "ITCH": "encodedBy", "ISFT": "encoder", "ISRF": "media",
                "IGNR": "genre", "ICMT": "comment", "ICOP": "copyright", "ILNG": "language"}      # A:198-220


def wav_out(x: np.ndarray, sampleRate, bitDepth=None, metadata=None) -> bytes:
    """Audio:wav (A:942-1004) for the PCM depths.  string.pack of a non-integral number is the host Lua's business;
    this restatement floors (what oracle/luavm does) and wraps modulo 2^bits like a two's-complement store."""
    import struct
    bitDepth = 16 if bitDepth is None else bitDepth
    if bitDepth not in (8, 16, 24, 32):
        raise OracleError("bad argument #2 (invalid bit depth)")                               # A:975
    x = np.atleast_2d(x)
    vals = audio_pcm(x, bitDepth, "unsigned" if bitDepth == 8 else "signed", True)              # A:976
    q = np.floor(vals).astype(np.int64) & ((1 << bitDepth) - 1)
    B = bitDepth // 8
    body = np.zeros((q.size, B), dtype=np.uint8)
    for k in range(B):
        body[:, k] = (q >> (8 * k)) & 0xFF
    body = body.tobytes()
    # A:981-985: 32768 values per pack for i = 1, #data - csize, csize; then #data % csize values from index
    # floor(#data / csize) * csize (one value early; data[0] = nil when #data < csize)
    cs, nvals = 32768, int(q.size)
    k, rem = nvals // cs, nvals % cs
    if k == 0:
        raise OracleError("bad argument #2 to 'pack' (number expected, got nil)")
    nloop = len(range(1, nvals - cs + 1, cs))
    body = body[: nloop * cs * B] + body[(k * cs - 1) * B: (k * cs - 1 + rem) * B]
    nc = x.shape[0]
    fmt = struct.pack("<HHIIHH", 1, nc, int(np.floor(sampleRate)), int(np.floor(sampleRate * nc * bitDepth / 8)), nc * bitDepth // 8, bitDepth)
    head = b"RIFF" + struct.pack("<I", len(body) + 36) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt   # A:1001-1003
    if metadata:
        lst = b"INFO"
        for k, v in metadata.items():
            tag = next((t for t, w in WAV_METADATA.items() if w == k), None)
            if tag is None:
                continue
            vb = v if isinstance(v, bytes) else str(v).encode("latin-1")
            lst += tag.encode() + struct.pack("<I", len(vb)) + vb
            if len(lst) % 2:
                lst += b"\0"                                                                     # "Xh" in "!2<..."
        head += b"LIST" + struct.pack("<I", len(lst)) + lst
    return head + b"data" + struct.pack("<I", len(body)) + body


def encode_pcm(d: float, bitDepth=8, dataType="signed") -> float:
    return float(lib().auko_encode_pcm(float(d), bitDepth, DATATYPES[dataType]))


class _Tag(C.Structure):
    _fields_ = [("id", C.c_char * 5), ("off", C.c_size_t), ("len", C.c_size_t)]


class _WavInfo(C.Structure):
    _fields_ = [("format", C.c_int), ("channels", C.c_int), ("sampleRate", C.c_int),
                ("blockAlign", C.c_int), ("bitDepth", C.c_int), ("have_fmt", C.c_int),
                ("ncoef", C.c_int), ("coef1", C.c_int * 256), ("coef2", C.c_int * 256),
                ("data_off", C.c_size_t), ("data_size", C.c_size_t), ("have_data", C.c_int),
                ("ntags", C.c_int), ("tags", _Tag * 64)]


WAV_FORMATS = ["signed", "unsigned", "float", "alaw", "ulaw", "adpcm", "msadpcm", "dfpwm", None]


def wav_parse(data) -> dict:
    b = _bytes(data)
    info = _WavInfo()
    _check(lib().auko_wav_parse(_ptr(b), b.size, C.byref(info)))
    raw = b.tobytes()
    return {
        "dataType": WAV_FORMATS[info.format], "channels": info.channels, "sampleRate": info.sampleRate,
        "blockAlign": info.blockAlign, "bitDepth": info.bitDepth, "have_fmt": bool(info.have_fmt),
        "coefficients": ([list(info.coef1[: info.ncoef]), list(info.coef2[: info.ncoef])] if info.ncoef else None),
        "data_off": info.data_off, "data_size": info.data_size,
        "tags": [(info.tags[i].id.decode("latin-1"), raw[info.tags[i].off: info.tags[i].off + info.tags[i].len])
                 for i in range(info.ntags)],
    }


def wav(data, dialect=LITERAL):
    """aukit.wav (A:1456-1574): parse + dispatch. Returns (samples [C, N] or list, info dict)."""
    b = _bytes(data)
    info = wav_parse(b)
    payload = b[info["data_off"]: info["data_off"] + info["data_size"]]
    dt = info["dataType"]
    if dt == "adpcm":
        x = wav_ima(payload, info["blockAlign"], info["channels"], dialect)
    elif dt == "msadpcm":
        x = msadpcm(payload, info["blockAlign"], info["channels"], info["coefficients"], dialect)
    elif dt in ("alaw", "ulaw"):
        x = g711(payload, dt == "ulaw", info["channels"])
    elif dt == "dfpwm":
        raise OracleError("dfpwm is outside the hot path")
    elif dt is None:
        x = pcm(payload, 8, "signed", 1, True, False)
    else:
        x = pcm(payload, info["bitDepth"], dt, info["channels"], True, False)
    return x, info


class _Meta(C.Structure):
    _fields_ = [("key", C.c_char * 12), ("off", C.c_size_t), ("len", C.c_size_t)]


class _ContainerInfo(C.Structure):
    _fields_ = [("codec", C.c_int), ("bitDepth", C.c_int), ("dataType", C.c_int), ("bigEndian", C.c_int), ("ulaw", C.c_int),
                ("channels", C.c_int), ("sampleRate", C.c_double), ("data_off", C.c_size_t), ("data_len", C.c_size_t),
                ("nmeta", C.c_int), ("meta", _Meta * 16)]


def _container(parse, data, head=False):
    b = _bytes(data)
    ci = _ContainerInfo()
    _check(parse(_ptr(b), b.size, C.byref(ci)))
    raw = b.tobytes()
    info = {"codec": "g711" if ci.codec else "pcm", "bitDepth": ci.bitDepth, "dataType": ["signed", "unsigned", "float"][ci.dataType],
            "bigEndian": bool(ci.bigEndian), "ulaw": bool(ci.ulaw), "channels": ci.channels, "sampleRate": ci.sampleRate,
            "data_off": ci.data_off, "data_len": ci.data_len,
            "metadata": {ci.meta[i].key.decode(): raw[ci.meta[i].off: ci.meta[i].off + ci.meta[i].len] for i in range(ci.nmeta)}}
    if head:
        return [np.zeros(0)] * max(ci.channels, 0), info
    payload = b[ci.data_off: ci.data_off + ci.data_len]
    if ci.codec:
        x = g711(payload, bool(ci.ulaw), ci.channels)
    else:
        x = pcm(payload, ci.bitDepth, info["dataType"], ci.channels, True, bool(ci.bigEndian))
    return x, info


def au(data):
    """aukit.au (A:1634-1647): (samples, info)."""
    return _container(lib().auko_au_parse, data)


def aiff(data, head=False):
    """aukit.aiff (A:1580-1631): (samples, info)."""
    return _container(lib().auko_aiff_parse, data, head)


def chain_s16(data, channels, srcRate, dstRate, interpolation="cubic", peak=1.0) -> np.ndarray:
    """decode s16le -> resample -> mono -> normalize in one C call (CPU baseline timing)."""
    b = _bytes(data)
    n = C.c_size_t(0)
    p = lib().auko_chain_s16(_ptr(b), b.size, channels, float(srcRate), float(dstRate),
                             INTERPS[interpolation], float(peak), C.byref(n))
    if not p:
        raise OracleError(lib().auko_last_error().decode("latin-1") or "chain failed")
    try:
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n.value,)).copy()
    finally:
        lib().auko_free(p)
    return arr
