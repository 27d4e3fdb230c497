"""Python stand-in for the Lua C module `aukit_cuda` (csrc/lua_binding.c), for running the Lua FACADE
(aukit_b200/lua/aukit.lua) inside oracle/luavm on the GPU box, where no Lua interpreter exists.  It
exposes exactly the functions luaopen_aukit_cuda registers, with the same argument conventions, but
reaches libaukit_cuda.so through the Python ctypes layer instead of the C binding."""
import ctypes as C

import numpy as np

from oracle.luavm.lua import LuaError, LuaFunction, LuaTable, to_lua


class Handle:
    """Lua userdata standing for an aukit_audio*."""

    def __init__(self, audio):
        self.audio = audio


def make_module(ak):
    ctx = ak.context()
    lib = ctx.lib

    def wrap(call, *args):
        out = C.c_void_p()
        rc = call(ctx.handle, *args, C.byref(out))
        if rc != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return Handle(ak.Audio(ctx, out))

    def num(v, default=None):
        return default if v is None else v

    def ints(t, n):
        if not isinstance(t, LuaTable):
            return None
        return (C.c_int * n)(*[int(t.get(i + 1)) for i in range(n)])

    def buf(b):
        return C.cast(C.c_char_p(b), C.c_void_p), len(b)

    def l_pcm(a):
        p, n = buf(a[0])
        return [wrap(lib.aukit_cuda_pcm, p, n, int(num(a[1] if len(a) > 1 else None, 8)), int(num(a[2] if len(a) > 2 else None, 0)),
                     int(num(a[3] if len(a) > 3 else None, 1)), float(num(a[4] if len(a) > 4 else None, 48000)),
                     int(bool(num(a[5] if len(a) > 5 else None, True))), int(bool(num(a[6] if len(a) > 6 else None, False))))]

    def l_g711(a):
        p, n = buf(a[0])
        return [wrap(lib.aukit_cuda_g711, p, n, int(bool(a[1])), int(num(a[2] if len(a) > 2 else None, 1)), float(num(a[3] if len(a) > 3 else None, 8000)))]

    def l_wav(a):
        data = a[0]
        p, n = buf(data)
        info = ak.WavInfo()
        out = C.c_void_p()
        rc = lib.aukit_cuda_wav(ctx.handle, p, n, int(bool(a[1] if len(a) > 1 else False)), int(num(a[2] if len(a) > 2 else None, 0)),
                                C.byref(info), C.byref(out))
        if rc != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        names = [b"signed", b"unsigned", b"float", b"alaw", b"ulaw", b"adpcm", b"msadpcm", b"dfpwm", None]
        t = LuaTable()
        if names[info.format] is not None:
            t.set(b"dataType", names[info.format])
        t.set(b"channels", float(info.channels))
        t.set(b"sampleRate", float(info.sampleRate))
        t.set(b"blockAlign", float(info.blockAlign))
        if info.have_fmt:
            t.set(b"bitDepth", float(info.bitDepth))
        tags = LuaTable()
        for i in range(info.ntags):
            tg = info.tags[i]
            e = LuaTable()
            e.arr = [tg.id, data[tg.off: tg.off + tg.len]]
            tags.set(i + 1, e)
        t.set(b"tags", tags)
        return [Handle(ak.Audio(ctx, out)), t]

    def container(fn, a, *extra):
        data = a[0]
        p, n = buf(data)
        ci = ak.ContainerInfo()
        out = C.c_void_p()
        if fn(ctx.handle, p, n, *extra, C.byref(ci), C.byref(out)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        t = LuaTable()
        t.set(b"codec", b"g711" if ci.codec else b"pcm")
        t.set(b"bitDepth", float(ci.bitDepth))
        t.set(b"dataType", [b"signed", b"unsigned", b"float"][ci.dataType])
        t.set(b"ulaw", bool(ci.ulaw))
        meta = LuaTable()
        for i in range(ci.nmeta):
            e = LuaTable()
            e.arr = [ci.meta[i].key, data[ci.meta[i].off: ci.meta[i].off + ci.meta[i].len]]
            meta.set(i + 1, e)
        t.set(b"meta", meta)
        return [Handle(ak.Audio(ctx, out)), t]

    def l_au(a):
        return container(lib.aukit_cuda_au, a)

    def l_aiff(a):
        return container(lib.aukit_cuda_aiff, a, int(bool(a[1] if len(a) > 1 else False)))

    def l_resample(a):
        return [wrap(lib.aukit_cuda_resample, a[0].audio._h, float(a[1]), int(a[2]))]

    def l_mono(a):
        return [wrap(lib.aukit_cuda_mono, a[0].audio._h)]

    def l_amplify(a):
        if lib.aukit_cuda_amplify(ctx.handle, a[0].audio._h, float(a[1])) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return []

    def l_lowpass(a):
        if lib.aukit_cuda_lowpass(ctx.handle, a[0].audio._h, float(a[1])) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return []

    def simple(fn, nargs, defaults=()):
        def g(a):
            vals = [float(num(a[i + 1] if len(a) > i + 1 else None, defaults[i] if i < len(defaults) else None)) for i in range(nargs)]
            if fn(ctx.handle, a[0].audio._h, *vals) != 0:
                raise LuaError(lib.aukit_cuda_last_error())
            return []
        return g

    def l_pcm_out(a):
        au = a[0].audio
        out = np.empty(au.frames * au.channels(), dtype=np.float64)
        il = True if len(a) < 4 or a[3] is None else bool(a[3])
        if lib.aukit_cuda_audio_pcm(ctx.handle, au._h, int(a[1]), int(a[2]), int(il), C.c_void_p(out.ctypes.data)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        t = LuaTable()
        t.arr = [float(v) for v in out]
        return [t]

    def l_pcm_bytes(a):
        au = a[0].audio
        bits = int(a[1])
        out = np.empty(au.frames * au.channels() * (bits // 8), dtype=np.uint8)
        il = True if len(a) < 4 or a[3] is None else bool(a[3])
        if lib.aukit_cuda_audio_pcm_bytes(ctx.handle, au._h, bits, int(a[2]), int(il), int(num(a[4] if len(a) > 4 else None, 1)),
                                          C.c_void_p(out.ctypes.data)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return [out.tobytes()]

    def l_normalize(a):
        if lib.aukit_cuda_normalize(ctx.handle, a[0].audio._h, float(num(a[1] if len(a) > 1 else None, 1.0)),
                                    int(bool(a[2] if len(a) > 2 else False))) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return []

    def l_frames(a):
        au = a[0].audio
        if len(a) > 1 and a[1] is not None:
            return [float(lib.aukit_cuda_audio_channel_frames(au._h, int(a[1]) - 1))]
        return [float(au.frames)]

    def l_read(a):
        au, c, first, count = a[0].audio, int(a[1]) - 1, int(a[2]) - 1, int(a[3])
        out = np.empty(count, dtype=np.float32)
        if lib.aukit_cuda_audio_download(ctx.handle, au._h, c, first, count, C.c_void_p(out.ctypes.data)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        t = LuaTable()
        t.arr = [float(v) for v in out]
        return [t]

    def l_new(a):
        out = C.c_void_p()
        if lib.aukit_cuda_audio_new(ctx.handle, int(a[0]), int(a[1]), float(a[2]), C.byref(out)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return [Handle(ak.Audio(ctx, out))]

    def l_set_sample_rate(a):
        if lib.aukit_cuda_audio_set_sample_rate(a[0].audio._h, float(a[1])) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return []

    def l_stream_chunk(a):
        au, bits, dt, first, count = a[0].audio, int(a[1]), int(a[2]), int(a[3]) - 1, int(a[4])
        nch = au.channels()
        out = np.empty((nch, max(count, 1)), dtype=np.float64)
        got = C.c_size_t(0)
        if lib.aukit_cuda_audio_stream_chunk(ctx.handle, au._h, bits, dt, first, count, C.c_void_p(out.ctypes.data), C.byref(got)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        if got.value == 0:
            return [None]
        t = LuaTable()
        for c in range(nch):
            e = LuaTable()
            e.arr = [float(v) for v in out[c, : got.value]]
            t.set(c + 1, e)
        return [t]

    def l_preload(a):
        a = list(a) + [None] * (10 - len(a))
        data = a[0]
        p, n = buf(data)
        bits, ch = int(num(a[1], 16)), int(num(a[3], 2))
        frames = n // (ch * bits // 8)
        src, dst = float(num(a[4], 44100)), float(num(a[5], 48000))
        d = ak.PipelineDesc(bits, int(num(a[2], 0)), ch, int(bool(num(a[9], False))), src, dst, int(num(a[6], 1)), int(bool(num(a[7], True))),
                            frames, 0, frames, 0, int(lib.aukit_resample_out_len(frames, src, dst)))
        out = C.c_void_p()
        if lib.aukit_cuda_preload_audio(ctx.handle, C.byref(d), p, n, float(num(a[8], 1.0)), C.byref(out)) != 0:
            raise LuaError(lib.aukit_cuda_last_error())
        return [Handle(ak.Audio(ctx, out))]

    mod = LuaTable()
    for name, f in {"preload": l_preload, "device_count": lambda a: [1.0], "set_sample_rate": l_set_sample_rate, "stream_chunk": l_stream_chunk, "pcm": l_pcm, "g711": l_g711, "wav": l_wav, "resample": l_resample, "mono": l_mono, "amplify": l_amplify,
                    "lowpass": l_lowpass, "pcm_out": l_pcm_out, "pcm_bytes": l_pcm_bytes, "invert": simple(lib.aukit_cuda_invert, 0),
                    "fade": simple(lib.aukit_cuda_fade, 4), "delay": simple(lib.aukit_cuda_delay, 2, (None, 0.5)),
                    "center": simple(lib.aukit_cuda_center, 0), "highpass": simple(lib.aukit_cuda_highpass, 1), "au": l_au, "aiff": l_aiff, "normalize": l_normalize, "frames": l_frames, "read": l_read, "new": l_new,
                    "channels": lambda a: [float(a[0].audio.channels())],
                    "sample_rate": lambda a: [float(lib.aukit_cuda_audio_sample_rate(a[0].audio._h))]}.items():
        mod.set(name.encode(), LuaFunction(f, "aukit_cuda." + name))
    mod.set(b"abi_version", 1.0)
    return mod
