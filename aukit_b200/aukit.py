"""Host-side mirror of the reference's Lua API for the preload path, over libaukit_cuda.so.

The reference is Lua (aukit.lua); no Lua interpreter exists in this image, so the host layer
above the C ABI is Python with the reference's names, argument order, defaults and error
strings ("A:n" = /root/reference/aukit.lua line n):

    aukit.pcm(data, bitDepth=8, dataType="signed", channels=1, sampleRate=48000,
              interleaved=True, bigEndian=False)                         A:1049
    aukit.g711(data, ulaw, channels=1, sampleRate=8000)                  A:1361
    aukit.adpcm(data, channels=1, sampleRate=48000, topFirst=True, interleaved=True,
                predictor=None, step_index=None)                         A:1183
    aukit.msadpcm(data, blockAlign, channels=1, sampleRate=48000, coefficients=None)  A:1283
    aukit.wav(data, head=False)                                          A:1456
    Audio.resample(sampleRate, interpolation=None)   (new object)        A:653
    Audio.mono()                                     (new object)        A:677
    Audio.len(), Audio.channels()                                        A:638-646
    effects.amplify(audio, multiplier)               (in place, returns audio)  A:3356
    effects.normalize(audio, peakAmplitude=1, independent=None)  (in place)     A:3431
    aukit.defaultInterpolation = "linear"                                A:99

Indices are Python's (0-based); everything else follows the reference.  The Lua facade with
the identical surface is aukit_b200/lua/aukit.lua (bound through luaopen_aukit_cuda).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AukitError, ContainerInfo, PipelineDesc, WavInfo

_VERSION = "1.10.0-b200"
defaultInterpolation = "linear"                                           # A:99

_DATATYPES = {"signed": 0, "unsigned": 1, "float": 2}
_INTERPS = {"none": 0, "linear": 1, "cubic": 2, "sinc": 3}
DIALECT_LITERAL, DIALECT_GENERAL = 0, 1
_WAV_TYPES = ["signed", "unsigned", "float", "alaw", "ulaw", "adpcm", "msadpcm", "dfpwm", None]

# INFO tag names, A:198-220
_WAV_METADATA = {
    "IPRD": "album", "INAM": "title", "IART": "artist", "IWRI": "author", "IMUS": "composer",
    "IPRO": "producer", "IPRT": "trackNumber", "ITRK": "trackNumber", "IFRM": "trackCount",
    "PRT1": "partNumber", "PRT2": "partCount", "TLEN": "length", "IRTD": "rating", "ICRD": "date",
    "ITCH": "encodedBy", "ISFT": "encoder", "ISRF": "media", "IGNR": "genre", "ICMT": "comment",
    "ICOP": "copyright", "ILNG": "language",
}


# ------------------------------------------------------------------ context
class Context:
    """One aukit_ctx (device + stream).  Kernels are enqueued on its stream."""

    def __init__(self, device: int = -1, _borrowed=None):
        self.lib = _lib.load()
        self._owned = _borrowed is None
        if _borrowed is not None:                     # a context that belongs to a Group
            self.handle = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        _lib.check(self.lib.aukit_cuda_init(device, C.byref(h)))
        self.handle = h

    def make_current(self):
        """cudaSetDevice to this context's device (needed only when one process holds contexts on several devices)."""
        _lib.check(self.lib.aukit_cuda_make_current(self.handle))

    def set_stream(self, cuda_stream: Optional[int]):
        """None -> the context's own stream.  A raw cudaStream_t handle otherwise; torch reports the
        legacy default stream as handle 0, which is passed on as cudaStreamLegacy (1)."""
        if cuda_stream is None:
            h = 0
        else:
            h = int(cuda_stream) or 1
        _lib.check(self.lib.aukit_cuda_set_stream(self.handle, C.c_void_p(h)))

    def use_torch_stream(self):
        """Enqueue on torch's current stream (ordering with torch ops and NCCL collectives)."""
        import torch
        self.set_stream(torch.cuda.current_stream().cuda_stream)

    def synchronize(self):
        _lib.check(self.lib.aukit_cuda_synchronize(self.handle))

    @property
    def launches(self) -> int:
        return int(self.lib.aukit_cuda_launch_count(self.handle))

    def close(self):
        if self.handle:
            if self._owned:
                self.lib.aukit_cuda_shutdown(self.handle)
            self.handle = None


_tls = threading.local()


def context(device: Optional[int] = None) -> Context:
    """The calling thread's default context (created on first use; no CPU fallback)."""
    ctxs = getattr(_tls, "ctxs", None)
    if ctxs is None:
        ctxs = _tls.ctxs = {}
    key = -1 if device is None else int(device)
    if key not in ctxs:
        ctxs[key] = Context(key)
    return ctxs[key]


def _expect(index, value, *types):
    """cc.expect (A:84): type check with the reference's message."""
    names = {str: "string", bytes: "string", bytearray: "string", memoryview: "string", int: "number",
             float: "number", bool: "boolean", type(None): "nil", dict: "table", list: "table", tuple: "table"}
    for t in types:
        if t is float and isinstance(value, (int, float)) and not isinstance(value, bool):
            return value
        if t is bool and isinstance(value, bool):
            return value
        if t is not float and t is not bool and isinstance(value, t) and not (t is int and isinstance(value, bool)):
            return value
    want = sorted({names.get(t, t.__name__) for t in types})
    got = names.get(type(value), type(value).__name__)
    if isinstance(value, np.ndarray):
        got = "table"
    if len(want) > 1:
        exp = ", ".join(want[:-1]) + " or " + want[-1]
    else:
        exp = want[0]
    raise AukitError("bad argument #%d (expected %s, got %s)" % (index, exp, got))


def _as_bytes(data):
    if isinstance(data, (bytes, bytearray)):
        return bytes(data) if isinstance(data, bytearray) else data
    if isinstance(data, memoryview):
        return data.tobytes()
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    raise AukitError("bad argument #1 (expected string, got %s)" % type(data).__name__)


def _buf(b):
    """(ctypes pointer, nbytes, keepalive) for bytes or a uint8 ndarray."""
    if isinstance(b, np.ndarray):
        return C.c_void_p(b.ctypes.data), b.size, b
    return C.cast(C.c_char_p(b), C.c_void_p), len(b), b


# ------------------------------------------------------------------ Audio
class _ChannelData:
    """audio.data: sequence of per-channel float32 arrays, downloaded lazily."""

    def __init__(self, audio: "Audio"):
        self._a = audio

    def __len__(self):
        return self._a.channels()

    def __getitem__(self, c):
        n = len(self)
        if isinstance(c, slice):
            return [self[i] for i in range(*c.indices(n))]
        if c < 0:
            c += n
        if not 0 <= c < n:
            raise IndexError(c)
        return self._a._channel(c)

    def __iter__(self):
        return (self[c] for c in range(len(self)))


class Audio:
    """Device-resident aukit.Audio (A:116-123): planar float32 samples owned by the C library."""

    def __init__(self, ctx: Context, handle, metadata=None, info=None):
        self._ctx = ctx
        self._h = handle
        self.metadata = {} if metadata is None else metadata
        self.info = {} if info is None else info

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._ctx.handle:
            self._ctx.lib.aukit_cuda_audio_free(self._ctx.handle, h)

    # --- reference surface
    @property
    def sampleRate(self):
        r = self._ctx.lib.aukit_cuda_audio_sample_rate(self._h)
        return int(r) if r == int(r) else r

    @sampleRate.setter
    def sampleRate(self, rate):
        # a plain writable field in the reference (effects.speed assigns it, A:3383): later device calls see it
        _lib.check(self._ctx.lib.aukit_cuda_audio_set_sample_rate(self._h, float(rate)))

    @property
    def data(self):
        return _ChannelData(self)

    def len(self) -> float:                                               # A:638
        return self.frames / self._ctx.lib.aukit_cuda_audio_sample_rate(self._h)

    def channels(self) -> int:                                            # A:644
        return int(self._ctx.lib.aukit_cuda_audio_channels(self._h))

    def resample(self, sampleRate, interpolation=None) -> "Audio":        # A:653
        _expect(1, sampleRate, float)
        interpolation = _expect(2, interpolation, str, type(None)) or defaultInterpolation
        if interpolation not in _INTERPS:
            raise AukitError("bad argument #2 (invalid interpolation type)")
        out = C.c_void_p()
        _lib.check(self._ctx.lib.aukit_cuda_resample(self._ctx.handle, self._h, float(sampleRate),
                                                     _INTERPS[interpolation], C.byref(out)))
        return Audio(self._ctx, out, dict(self.metadata), dict(self.info))   # copy(), A:657

    def mono(self) -> "Audio":                                            # A:677
        out = C.c_void_p()
        _lib.check(self._ctx.lib.aukit_cuda_mono(self._ctx.handle, self._h, C.byref(out)))
        return Audio(self._ctx, out, dict(self.metadata), dict(self.info))

    def concat(self, *others: "Audio") -> "Audio":                        # A:696
        parts = [self]
        for i, o in enumerate(others):
            if not isinstance(o, Audio):
                raise AukitError("bad argument #%d (expected Audio, got %s)" % (i + 1, type(o).__name__))
            if o.sampleRate != self.sampleRate:
                o = o.resample(self.sampleRate)                           # A:702
            parts.append(o)
        arr = (C.c_void_p * len(parts))(*[p._h for p in parts])
        out = C.c_void_p()
        _lib.check(self._ctx.lib.aukit_cuda_concat(self._ctx.handle, arr, len(parts), C.byref(out)))
        return Audio(self._ctx, out, dict(self.metadata), dict(self.info))

    # --- Python-side conveniences
    @property
    def frames(self) -> int:
        return int(self._ctx.lib.aukit_cuda_audio_frames(self._h))

    @property
    def stride(self) -> int:
        return int(self._ctx.lib.aukit_cuda_audio_stride(self._h))

    @property
    def data_ptr(self) -> int:
        return int(self._ctx.lib.aukit_cuda_audio_data(self._h) or 0)

    def _channel(self, c: int) -> np.ndarray:
        n = int(self._ctx.lib.aukit_cuda_audio_channel_frames(self._h, c))
        out = np.empty(n, dtype=np.float32)
        _lib.check(self._ctx.lib.aukit_cuda_audio_download(self._ctx.handle, self._h, c, 0, n,
                                                           C.c_void_p(out.ctypes.data)))
        return out

    def pcm(self, bitDepth=None, dataType=None, interleaved=None) -> np.ndarray:          # A:901
        """Audio:pcm: the samples as un-rounded PCM values (float64, flat: interleaved or channel-major)."""
        bitDepth = _expect(1, bitDepth, float, type(None)) or 8
        dataType = _expect(2, dataType, str, type(None)) or "signed"
        _expect(3, interleaved, bool, type(None))
        if interleaved is None:
            interleaved = True
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        if dataType not in _DATATYPES:
            raise AukitError("bad argument #3 (invalid data type)")
        if dataType == "float" and bitDepth != 32:
            raise AukitError("bad argument #2 (float audio must have 32-bit depth)")
        out = np.empty(self.frames * self.channels(), dtype=np.float64)
        _lib.check(self._ctx.lib.aukit_cuda_audio_pcm(self._ctx.handle, self._h, int(bitDepth), _DATATYPES[dataType],
                                                      int(interleaved), C.c_void_p(out.ctypes.data)))
        return out

    def stream(self, chunkSize=None, bitDepth=None, dataType=None):        # A:921-937
        """Audio:stream: (iterator, total length in seconds).  Each step of the iterator returns
        (chunks, position): one float64 array of un-rounded PCM values per channel (encodePCM, A:868-894)
        and the position of the chunk in seconds -- (1-based pos) / sampleRate, as the reference computes it."""
        chunkSize = _expect(1, chunkSize, float, type(None)) or 131072
        bitDepth = _expect(2, bitDepth, float, type(None)) or 8
        dataType = _expect(3, dataType, str, type(None)) or "signed"
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        if dataType not in _DATATYPES:
            raise AukitError("bad argument #3 (invalid data type)")
        if dataType == "float" and bitDepth != 32:
            raise AukitError("bad argument #2 (float audio must have 32-bit depth)")
        chunk, bits, dt = int(chunkSize), int(bitDepth), _DATATYPES[dataType]
        rate = self._ctx.lib.aukit_cuda_audio_sample_rate(self._h)
        nch = self.channels()

        def it():
            pos = 1                                                       # the reference's 1-based frame index
            while True:
                buf = np.empty((nch, chunk), dtype=np.float64)
                got = C.c_size_t(0)
                _lib.check(self._ctx.lib.aukit_cuda_audio_stream_chunk(self._ctx.handle, self._h, bits, dt, pos - 1, chunk,
                                                                       C.c_void_p(buf.ctypes.data), C.byref(got)))
                if got.value == 0:
                    return
                yield [buf[c, : got.value].copy() for c in range(nch)], pos / rate
                pos += chunk
        return it(), self.frames / rate

    def pcm_bytes(self, bitDepth=16, dataType="signed", interleaved=True, rounding="truncate") -> bytes:
        """The packed little-endian samples Audio:wav writes (A:981-985); `rounding` is the host
        string.pack's: "truncate" (Cobalt), "floor" or "nearest"."""
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        if dataType not in _DATATYPES:
            raise AukitError("bad argument #3 (invalid data type)")
        out = np.empty(self.frames * self.channels() * (bitDepth // 8), dtype=np.uint8)
        _lib.check(self._ctx.lib.aukit_cuda_audio_pcm_bytes(self._ctx.handle, self._h, int(bitDepth), _DATATYPES[dataType],
                                                            int(bool(interleaved)), {"truncate": 0, "floor": 1, "nearest": 2}[rounding],
                                                            C.c_void_p(out.ctypes.data)))
        return out.tobytes()

    def wav(self, bitDepth=None, rounding="floor", dialect=DIALECT_GENERAL) -> bytes:   # A:942-1004
        """Audio:wav: a RIFF/WAVE file of the samples (8-bit unsigned, otherwise signed PCM).  The reference packs the
        un-rounded numbers of Audio:pcm with the host's string.pack; `rounding` says how that pack narrows them
        ("floor", "truncate" = a C cast, "nearest").  Header quirk kept: with metadata the RIFF size field still
        counts only the header and the samples (A:1001).
        dialect LITERAL also keeps the chunk arithmetic of A:981-985: values are packed 32768 at a time for
        i = 1, #data - 32768, 32768, then a tail of #data % 32768 values that starts ONE VALUE EARLY (at index
        floor(#data / 32768) * 32768) -- so the last value of the file is dropped, a whole final chunk is lost when
        #data is a multiple of 32768, and fewer than 32768 values raise (the tail's first argument is data[0] = nil)."""
        import struct
        bitDepth = _expect(1, bitDepth, float, type(None)) or 16
        if bitDepth == 1:
            raise AukitError("aukit_cuda: DFPWM output is outside the accelerated path")
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        bitDepth = int(bitDepth)
        body = self.pcm_bytes(bitDepth, "unsigned" if bitDepth == 8 else "signed", True, rounding)
        if dialect == DIALECT_LITERAL:
            B, nvals, cs = bitDepth // 8, self.frames * self.channels(), 32768
            k, rem = nvals // cs, nvals % cs
            if k == 0:
                raise AukitError("bad argument #2 to 'pack' (number expected, got nil)")
            nloop = len(range(1, nvals - cs + 1, cs))
            body = body[: nloop * cs * B] + body[(k * cs - 1) * B: (k * cs - 1 + rem) * B]
        ch, rate = self.channels(), int(np.floor(self.sampleRate))
        fmt = struct.pack("<HHIIHH", 1, ch, rate, int(self.sampleRate * ch * bitDepth / 8), ch * bitDepth // 8, bitDepth)
        head = b"RIFF" + struct.pack("<I", len(body) + 36) + b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt
        if self.metadata:
            inv = {}
            for tag, key in _WAV_METADATA.items():
                inv.setdefault(key, tag)
            lst = b"INFO"
            for k, v in self.metadata.items():
                if k in inv:
                    if isinstance(v, float) and v == int(v):
                        v = int(v)
                    vb = v if isinstance(v, bytes) else str(v).encode("latin-1")
                    lst += inv[k].encode() + struct.pack("<I", len(vb)) + vb
                    if len(lst) % 2:
                        lst += b"\0"
            head += b"LIST" + struct.pack("<I", len(lst)) + lst
        return head + b"data" + struct.pack("<I", len(body)) + body

    def numpy(self) -> np.ndarray:
        """[channels, frames] float32 copy on the host (channel 1's length, like #data[1])."""
        n = self.frames
        out = np.zeros((self.channels(), n), dtype=np.float32)
        for c in range(self.channels()):
            ch = self._channel(c)
            out[c, : ch.size] = ch[:n]
        return out

    @classmethod
    def from_numpy(cls, x: np.ndarray, sampleRate=48000, ctx: Optional[Context] = None) -> "Audio":
        ctx = ctx or context()
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
        h = C.c_void_p()
        _lib.check(ctx.lib.aukit_cuda_audio_new(ctx.handle, x.shape[0], x.shape[1], float(sampleRate), C.byref(h)))
        a = cls(ctx, h)
        for c in range(x.shape[0]):
            _lib.check(ctx.lib.aukit_cuda_audio_upload(ctx.handle, h, c, 0, x.shape[1], C.c_void_p(x[c].ctypes.data)))
        return a

    def __repr__(self):
        return "<aukit_b200.Audio %d ch x %d frames @ %s Hz on device>" % (self.channels(), self.frames, self.sampleRate)


def _expect_audio(n, v) -> Audio:
    if isinstance(v, Audio):
        return v
    raise AukitError("bad argument #%d (expected Audio, got %s)" % (n, type(v).__name__))   # A:234-237


# ------------------------------------------------------------------ loaders
def new(duration, channels=None, sampleRate=None, ctx: Optional[Context] = None) -> Audio:   # A:1784
    _expect(1, duration, float)
    channels = _expect(2, channels, int, type(None)) or 1
    sampleRate = _expect(3, sampleRate, float, type(None)) or 48000
    ctx = ctx or context()
    h = C.c_void_p()
    frames = int(np.floor(duration * sampleRate)) if duration * sampleRate >= 1 else 0
    _lib.check(ctx.lib.aukit_cuda_audio_new(ctx.handle, channels, frames, float(sampleRate), C.byref(h)))
    return Audio(ctx, h, {}, {})


def pcm(data, bitDepth=None, dataType=None, channels=None, sampleRate=None, interleaved=None, bigEndian=None,
        ctx: Optional[Context] = None) -> Audio:                          # A:1049
    if isinstance(data, (list, tuple)):
        raise AukitError("aukit_b200: table (pre-unpacked) PCM input is host-side only; pass the packed string")
    bitDepth = _expect(2, bitDepth, int, type(None)) or 8
    dataType = _expect(3, dataType, str, type(None)) or "signed"
    channels = _expect(4, channels, int, type(None)) or 1
    sampleRate = _expect(5, sampleRate, float, type(None)) or 48000
    _expect(6, interleaved, bool, type(None))
    if interleaved is None:
        interleaved = True
    _expect(7, bigEndian, bool, type(None))
    if bitDepth not in (8, 16, 24, 32):
        raise AukitError("bad argument #2 (invalid bit depth)")
    if dataType not in _DATATYPES:
        raise AukitError("bad argument #3 (invalid data type)")
    ctx = ctx or context()
    p, n, keep = _buf(_as_bytes(data))
    out = C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_pcm(ctx.handle, p, n, bitDepth, _DATATYPES[dataType], channels, float(sampleRate),
                                      int(interleaved), int(bool(bigEndian)), C.byref(out)))
    return Audio(ctx, out, {}, {"bitDepth": bitDepth, "dataType": dataType})   # A:1072


def g711(data, ulaw, channels=None, sampleRate=None, ctx: Optional[Context] = None) -> Audio:   # A:1361
    _expect(2, ulaw, bool)
    channels = _expect(3, channels, int, type(None)) or 1
    sampleRate = _expect(4, sampleRate, float, type(None)) or 8000
    ctx = ctx or context()
    p, n, keep = _buf(_as_bytes(data))
    out = C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_g711(ctx.handle, p, n, int(ulaw), channels, float(sampleRate), C.byref(out)))
    # A:1383 stores bitDepth/dataType in `metadata` and leaves `info` empty
    return Audio(ctx, out, {"bitDepth": 14 if ulaw else 13, "dataType": "signed"}, {})


def adpcm(data, channels=None, sampleRate=None, topFirst=None, interleaved=None, predictor=None, step_index=None,
          ctx: Optional[Context] = None) -> Audio:                        # A:1183
    channels = _expect(2, channels, int, type(None)) or 1
    sampleRate = _expect(3, sampleRate, float, type(None)) or 48000
    _expect(4, topFirst, bool, type(None))
    if topFirst is None:
        topFirst = True
    _expect(5, interleaved, bool, type(None))
    if interleaved is None:
        interleaved = True

    def _state(v, argn, lo, hi):
        if v is None:
            return None
        if isinstance(v, (int, float)) and not isinstance(v, bool):
            if channels != 1:
                raise AukitError("bad argument #%d (table too short)" % argn)
            v = [v]
        if channels > len(v):
            raise AukitError("bad argument #%d (table too short)" % argn)
        for x in v[:channels]:
            if not lo <= x <= hi:
                raise AukitError("number outside of range (expected %s to be within %d and %d)" % (x, lo, hi))
        return (C.c_int * channels)(*[int(x) for x in v[:channels]])

    pr, si = _state(predictor, 6, -32768, 32767), _state(step_index, 7, 0, 88)
    ctx = ctx or context()
    p, n, keep = _buf(_as_bytes(data))
    out = C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_adpcm(ctx.handle, p, n, channels, float(sampleRate), int(topFirst), int(interleaved),
                                        pr, si, C.byref(out)))
    return Audio(ctx, out, {}, {"bitDepth": 16, "dataType": "signed"})


def msadpcm(data, blockAlign, channels=None, sampleRate=None, coefficients=None, dialect=DIALECT_LITERAL,
            ctx: Optional[Context] = None) -> Audio:                      # A:1283
    _expect(2, blockAlign, int)
    channels = _expect(3, channels, int, type(None)) or 1
    sampleRate = _expect(4, sampleRate, float, type(None)) or 48000
    c1 = c2 = None
    nco = 0
    if coefficients is not None:
        if not isinstance(coefficients[0], (list, tuple)):
            raise AukitError("bad argument #5 (first entry is not a table)")
        if not isinstance(coefficients[1], (list, tuple)):
            raise AukitError("bad argument #5 (second entry is not a table)")
        if len(coefficients[0]) != len(coefficients[1]):
            raise AukitError("bad argument #5 (lists are not the same length)")
        nco = len(coefficients[0])
        c1 = (C.c_int * nco)(*[int(v) for v in coefficients[0]])
        c2 = (C.c_int * nco)(*[int(v) for v in coefficients[1]])
    ctx = ctx or context()
    p, n, keep = _buf(_as_bytes(data))
    out = C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_msadpcm(ctx.handle, p, n, blockAlign, channels, float(sampleRate), c1, c2, nco,
                                          dialect, C.byref(out)))
    return Audio(ctx, out, {}, {"bitDepth": 16, "dataType": "signed"})


def _lua_tonumber(s: bytes):
    """tonumber(str) for INFO tags (A:1566): decimal / hex numerals with surrounding whitespace."""
    try:
        t = s.decode("latin-1").strip(" \t\n\r\f\v")
        if not t or "\0" in t or "_" in t:
            return None
        if t.lower().lstrip("+-").startswith("0x"):
            return float(int(t, 16)) if "." not in t and "p" not in t.lower() else float.fromhex(t)
        if t.lower().lstrip("+-") in ("inf", "infinity", "nan"):
            return None
        return float(t)
    except ValueError:
        return None


def wav_info(data) -> dict:
    """Host-only container walk of aukit.wav (A:1456-1574)."""
    lib = _lib.load()
    p, n, keep = _buf(_as_bytes(data))
    info = WavInfo()
    _lib.check(lib.aukit_cuda_wav_parse(p, n, C.byref(info)))
    return _info_dict(info, keep)


def _info_dict(info: WavInfo, raw) -> dict:
    raw = raw.tobytes() if isinstance(raw, np.ndarray) else raw
    meta = {}
    for i in range(info.ntags):
        t = info.tags[i]
        key = _WAV_METADATA.get(t.id.decode("latin-1"))
        if key:
            s = raw[t.off: t.off + t.len]
            num = _lua_tonumber(s)
            meta[key] = num if num is not None else s                     # tonumber(str) or str
    return {
        "dataType": _WAV_TYPES[info.format], "channels": info.channels, "sampleRate": info.sampleRate,
        "blockAlign": info.blockAlign, "bitDepth": info.bitDepth if info.have_fmt else None,
        "coefficients": [list(info.coef1[: info.ncoef]), list(info.coef2[: info.ncoef])] if info.ncoef else None,
        "data_off": info.data_off, "data_size": info.data_size, "metadata": meta,
    }


def wav(data, head=False, dialect=DIALECT_LITERAL, ctx: Optional[Context] = None) -> Audio:   # A:1456
    ctx = ctx or context()
    p, n, keep = _buf(_as_bytes(data))
    info = WavInfo()
    out = C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_wav(ctx.handle, p, n, int(bool(head)), dialect, C.byref(info), C.byref(out)))
    d = _info_dict(info, keep)
    return Audio(ctx, out, d["metadata"], {"dataType": d["dataType"], "bitDepth": d["bitDepth"]})   # A:1553-1554


def _container_audio(ctx, out, ci: ContainerInfo, raw: bytes, meta_override: bool) -> Audio:
    if ci.codec:                                   # aukit.g711: bitDepth/dataType live in `metadata` (A:1383)
        metadata, info = {"bitDepth": 14 if ci.ulaw else 13, "dataType": "signed"}, {}
    else:                                          # aukit.pcm: in `info` (A:1171)
        metadata, info = {}, {"bitDepth": ci.bitDepth, "dataType": ["signed", "unsigned", "float"][ci.dataType]}
    if meta_override:                              # aiff: obj.metadata = meta (A:1617)
        metadata = {ci.meta[i].key.decode(): raw[ci.meta[i].off: ci.meta[i].off + ci.meta[i].len] for i in range(ci.nmeta)}
    return Audio(ctx, out, metadata, info)


def au(data, ctx: Optional[Context] = None) -> Audio:                       # A:1634
    ctx = ctx or context()
    b = _as_bytes(data)
    p, n, keep = _buf(b)
    ci, out = ContainerInfo(), C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_au(ctx.handle, p, n, C.byref(ci), C.byref(out)))
    return _container_audio(ctx, out, ci, bytes(b), False)


def aiff(data, head=False, ctx: Optional[Context] = None) -> Audio:         # A:1580
    ctx = ctx or context()
    b = _as_bytes(data)
    p, n, keep = _buf(b)
    ci, out = ContainerInfo(), C.c_void_p()
    _lib.check(ctx.lib.aukit_cuda_aiff(ctx.handle, p, n, int(bool(head)), C.byref(ci), C.byref(out)))
    a = _container_audio(ctx, out, ci, bytes(b), True)
    if head:
        a.info = {}                                # aukit.new leaves info empty (A:1610)
    return a


# ------------------------------------------------------------------ effects (in place)
class effects:
    @staticmethod
    def amplify(audio: Audio, multiplier) -> Audio:                        # A:3356
        _expect_audio(1, audio)
        _expect(2, multiplier, float)
        _lib.check(audio._ctx.lib.aukit_cuda_amplify(audio._ctx.handle, audio._h, float(multiplier)))
        return audio

    @staticmethod
    def invert(audio: Audio) -> Audio:                                     # A:3412
        _expect_audio(1, audio)
        _lib.check(audio._ctx.lib.aukit_cuda_invert(audio._ctx.handle, audio._h))
        return audio

    @staticmethod
    def fade(audio: Audio, startTime, startAmplitude, endTime, endAmplitude) -> Audio:   # A:3392
        _expect_audio(1, audio)
        for i, v in enumerate((startTime, startAmplitude, endTime, endAmplitude)):
            _expect(i + 2, v, float)
        _lib.check(audio._ctx.lib.aukit_cuda_fade(audio._ctx.handle, audio._h, float(startTime), float(startAmplitude),
                                                  float(endTime), float(endAmplitude)))
        return audio

    @staticmethod
    def delay(audio: Audio, delay, multiplier=None) -> Audio:              # A:3500
        _expect_audio(1, audio)
        _expect(2, delay, float)
        multiplier = _expect(3, multiplier, float, type(None))
        _lib.check(audio._ctx.lib.aukit_cuda_delay(audio._ctx.handle, audio._h, float(delay),
                                                   0.5 if multiplier is None else float(multiplier)))
        return audio

    @staticmethod
    def center(audio: Audio) -> Audio:                                     # A:3465
        _expect_audio(1, audio)
        _lib.check(audio._ctx.lib.aukit_cuda_center(audio._ctx.handle, audio._h))
        return audio

    @staticmethod
    def lowpass(audio: Audio, frequency) -> Audio:                         # A:3586
        _expect_audio(1, audio)
        _expect(2, frequency, float)
        _lib.check(audio._ctx.lib.aukit_cuda_lowpass(audio._ctx.handle, audio._h, float(frequency)))
        return audio

    @staticmethod
    def highpass(audio: Audio, frequency) -> Audio:                        # A:3605
        _expect_audio(1, audio)
        _expect(2, frequency, float)
        _lib.check(audio._ctx.lib.aukit_cuda_highpass(audio._ctx.handle, audio._h, float(frequency)))
        return audio

    @staticmethod
    def normalize(audio: Audio, peakAmplitude=None, independent=None) -> Audio:   # A:3431
        _expect_audio(1, audio)
        peakAmplitude = _expect(2, peakAmplitude, float, type(None))
        if peakAmplitude is None:
            peakAmplitude = 1
        _expect(3, independent, bool, type(None))
        _lib.check(audio._ctx.lib.aukit_cuda_normalize(audio._ctx.handle, audio._h, float(peakAmplitude),
                                                       int(bool(independent))))
        return audio


# ------------------------------------------------------------------ fused auplay chain
def preload(data, bitDepth=16, dataType="signed", channels=2, sampleRate=44100, targetRate=48000,
            interpolation=None, mono=True, peakAmplitude=0.8, bigEndian=False,
            ctx: Optional[Context] = None) -> np.ndarray:
    """auplay.lua:12-27 in one call on packed interleaved PCM held in HOST memory:
    aukit.pcm -> :resample(targetRate) -> :mono() -> effects.normalize(peak).  H2D copy, the two
    fused passes and the D2H copy all happen inside.  Returns [1 or channels, frames] float32."""
    interpolation = interpolation or defaultInterpolation
    if interpolation not in _INTERPS:
        raise AukitError("bad argument #2 (invalid interpolation type)")
    if dataType not in _DATATYPES:
        raise AukitError("bad argument #3 (invalid data type)")
    ctx = ctx or context()
    b = _as_bytes(data)
    p, n, keep = _buf(b)
    B = bitDepth // 8
    if bitDepth not in (8, 16, 24, 32):
        raise AukitError("bad argument #2 (invalid bit depth)")
    if n % (B * channels):
        raise AukitError("bad argument #1 (uneven amount of data per channel)")
    frames = n // (B * channels)
    n_out = int(ctx.lib.aukit_resample_out_len(frames, float(sampleRate), float(targetRate)))
    d = PipelineDesc(bitDepth, _DATATYPES[dataType], channels, int(bool(bigEndian)), float(sampleRate),
                     float(targetRate), _INTERPS[interpolation], int(bool(mono)), frames, 0, frames, 0, n_out)
    out = np.empty((1 if mono else channels, n_out), dtype=np.float32)
    _lib.check(ctx.lib.aukit_cuda_pipeline_host(ctx.handle, C.byref(d), p, n, float(peakAmplitude),
                                                C.c_void_p(out.ctypes.data)))
    return out


def preload_clips(clips: Sequence, sampleRates, bitDepth=16, dataType="signed", channels=2, targetRate=48000,
                  interpolation=None, multiplier=1.0, bigEndian=False, ctx: Optional[Context] = None):
    """A playlist of clips through aukit.pcm -> :resample(targetRate, interpolation) -> effects.amplify(multiplier)
    (A:1049, A:653, A:3356) in one fused pass per rate class (BASELINE config 3).  `clips`: packed interleaved PCM
    (bytes / uint8 arrays) sharing one sample format; `sampleRates`: one rate per clip (or a single number).
    Returns one Audio per clip."""
    interpolation = interpolation or defaultInterpolation
    if interpolation not in _INTERPS:
        raise AukitError("bad argument #2 (invalid interpolation type)")
    if dataType not in _DATATYPES:
        raise AukitError("bad argument #3 (invalid data type)")
    ctx = ctx or context()
    n = len(clips)
    if isinstance(sampleRates, (int, float)):
        sampleRates = [sampleRates] * n
    bufs = [_buf(_as_bytes(c)) for c in clips]
    ptrs = (C.c_void_p * max(n, 1))(*[b[0] for b in bufs])
    sizes = (C.c_size_t * max(n, 1))(*[b[1] for b in bufs])
    rates = (C.c_double * max(n, 1))(*[float(r) for r in sampleRates])
    outs = (C.c_void_p * max(n, 1))()
    _lib.check(ctx.lib.aukit_cuda_batch_resample_amplify(ctx.handle, ptrs, sizes, rates, n, int(bitDepth), _DATATYPES[dataType], int(channels),
                                                         int(bool(bigEndian)), float(targetRate), _INTERPS[interpolation], float(multiplier),
                                                         outs))
    return [Audio(ctx, C.c_void_p(outs[k]), {}, {"bitDepth": bitDepth, "dataType": dataType}) for k in range(n)]


class Group:
    """One host thread driving several GPUs (aukit_cuda_group_*): what the Lua module uses when more than one B200 is
    visible.  `devices`: list of device ordinals (None = all visible); the same ordinal may appear more than once."""

    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if devices is None:
            _lib.check(self.lib.aukit_cuda_group_create(None, 0, C.byref(h)))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _lib.check(self.lib.aukit_cuda_group_create(arr, len(devices), C.byref(h)))
        self.handle = h
        self.size = int(self.lib.aukit_cuda_group_size(h))
        self.contexts = [Context(_borrowed=self.lib.aukit_cuda_group_ctx(h, i)) for i in range(self.size)]

    def preload(self, data, bitDepth=16, dataType="signed", channels=2, sampleRate=44100, targetRate=48000,
                interpolation=None, mono=True, peakAmplitude=0.8, bigEndian=False, owner: Optional[Context] = None):
        """preload() time-sharded over the group's GPUs: same arguments, same bits as on one GPU.  Returns a host array,
        or -- with `owner` (a Context) -- one device-resident Audio on that context (what the Lua module's cu.preload does)."""
        interpolation = interpolation or defaultInterpolation
        if interpolation not in _INTERPS:
            raise AukitError("bad argument #2 (invalid interpolation type)")
        if dataType not in _DATATYPES:
            raise AukitError("bad argument #3 (invalid data type)")
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        p, n, keep = _buf(_as_bytes(data))
        B = bitDepth // 8
        if n % (B * channels):
            raise AukitError("bad argument #1 (uneven amount of data per channel)")
        frames = n // (B * channels)
        n_out = int(self.lib.aukit_resample_out_len(frames, float(sampleRate), float(targetRate)))
        d = PipelineDesc(bitDepth, _DATATYPES[dataType], channels, int(bool(bigEndian)), float(sampleRate), float(targetRate),
                         _INTERPS[interpolation], int(bool(mono)), frames, 0, frames, 0, n_out)
        if owner is not None:                             # gathered device-to-device into one Audio on `owner`'s device
            h = C.c_void_p()
            _lib.check(self.lib.aukit_cuda_group_preload_audio(self.handle, owner.handle, C.byref(d), p, n, float(peakAmplitude), C.byref(h)))
            return Audio(owner, h, {}, {"bitDepth": bitDepth, "dataType": dataType})
        out = np.empty((1 if mono else channels, n_out), dtype=np.float32)
        _lib.check(self.lib.aukit_cuda_group_preload(self.handle, C.byref(d), p, n, float(peakAmplitude), C.c_void_p(out.ctypes.data)))
        return out

    def normalize(self, shards: Sequence["Audio"], peakAmplitude=None, independent=None):
        """effects.normalize on an Audio held as one time shard per member (shards[i] created on self.contexts[i])."""
        if len(shards) != self.size:
            raise AukitError("aukit_b200: one shard per group member")
        arr = (C.c_void_p * self.size)(*[s._h for s in shards])
        _lib.check(self.lib.aukit_cuda_group_normalize(self.handle, arr, 1.0 if peakAmplitude is None else float(peakAmplitude),
                                                       int(bool(independent))))
        return shards

    def close(self):
        if self.handle:
            for c in self.contexts:
                c.handle = None
            self.lib.aukit_cuda_group_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Preloader:
    """Pipelined preload() for a playlist of clips (aukit_cuda_preloader_*): clip i's download overlaps
    clip i+1's upload, `slots` device buffers in rotation.  submit() is asynchronous; results are valid
    after drain().  Host arrays should come from pinned() so the copies really are asynchronous.

        pl = Preloader(max_in_bytes, max_out_samples)
        for clip, out in zip(clips, outs): pl.submit(clip, out, sampleRate=44100)
        pl.drain()
    """

    def __init__(self, max_in_bytes: int, max_out_samples: int, slots: int = 2, ctx: Optional[Context] = None):
        self.ctx = ctx or context()
        h = C.c_void_p()
        _lib.check(self.ctx.lib.aukit_cuda_preloader_create(self.ctx.handle, int(max_in_bytes), int(max_out_samples), int(slots),
                                                            C.byref(h)))
        self.handle = h
        self._keep = []

    @staticmethod
    def pinned(shape, dtype) -> np.ndarray:
        """numpy array over pinned host memory (aukit_cuda_host_alloc); freed with the array."""
        lib = _lib.load()
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        _lib.check(lib.aukit_cuda_host_alloc(n, C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
        import weakref
        weakref.finalize(buf, lib.aukit_cuda_host_free, C.c_void_p(p.value))
        return arr

    def describe(self, nbytes, bitDepth=16, dataType="signed", channels=2, sampleRate=44100, targetRate=48000,
                 interpolation=None, mono=True, bigEndian=False) -> PipelineDesc:
        interpolation = interpolation or defaultInterpolation
        if interpolation not in _INTERPS:
            raise AukitError("bad argument #2 (invalid interpolation type)")
        if dataType not in _DATATYPES:
            raise AukitError("bad argument #3 (invalid data type)")
        if bitDepth not in (8, 16, 24, 32):
            raise AukitError("bad argument #2 (invalid bit depth)")
        B = bitDepth // 8
        if nbytes % (B * channels):
            raise AukitError("bad argument #1 (uneven amount of data per channel)")
        frames = nbytes // (B * channels)
        n_out = int(self.ctx.lib.aukit_resample_out_len(frames, float(sampleRate), float(targetRate)))
        return PipelineDesc(bitDepth, _DATATYPES[dataType], channels, int(bool(bigEndian)), float(sampleRate),
                            float(targetRate), _INTERPS[interpolation], int(bool(mono)), frames, 0, frames, 0, n_out)

    def submit(self, data: np.ndarray, out: np.ndarray, peakAmplitude=0.8, desc: Optional[PipelineDesc] = None, **fmt):
        """Enqueue one clip: `data` packed interleaved PCM bytes (uint8 array), `out` float32
        [1 or channels, n_out].  Both must stay alive and untouched until drain()."""
        d = desc or self.describe(data.nbytes, **fmt)
        self._keep.append((data, out, d))
        _lib.check(self.ctx.lib.aukit_cuda_preloader_submit(self.handle, C.byref(d), C.c_void_p(data.ctypes.data), data.nbytes,
                                                            float(peakAmplitude), C.c_void_p(out.ctypes.data)))
        return d

    def begin(self, data: np.ndarray, desc: PipelineDesc) -> int:
        k = C.c_int(-1)
        self._keep.append((data, desc))
        _lib.check(self.ctx.lib.aukit_cuda_preloader_begin(self.handle, C.byref(desc), C.c_void_p(data.ctypes.data), data.nbytes,
                                                           C.byref(k)))
        return k.value

    def peak_ptr(self, slot: int) -> int:
        return int(self.ctx.lib.aukit_cuda_preloader_peak_ptr(self.handle, slot) or 0)

    @property
    def stream(self) -> int:
        return int(self.ctx.lib.aukit_cuda_preloader_stream(self.handle) or 0)

    def finish(self, slot: int, out: np.ndarray, peakAmplitude=0.8):
        self._keep.append(out)
        _lib.check(self.ctx.lib.aukit_cuda_preloader_finish(self.handle, slot, float(peakAmplitude), C.c_void_p(out.ctypes.data)))

    def drain(self):
        try:
            _lib.check(self.ctx.lib.aukit_cuda_preloader_drain(self.handle))
        finally:
            self._keep.clear()

    def close(self):
        if self.handle:
            self.ctx.lib.aukit_cuda_preloader_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
