"""csrc/lua_binding.c (luaopen_aukit_cuda) EXECUTED: the module's own object code runs inside tests/luahost/luahost.c,
a toy host that implements the Lua 5.2 C API calls the binding declares (stack, strings, tables, userdata + metatables,
__gc, luaL_check* errors, lua_error as a longjmp to the protected call).  CPU tests: registration, argument checking,
error propagation, stack discipline.  GPU tests: the auplay chain and the other entry points against the ctypes path."""
import os
import re

import numpy as np
import pytest

from util import TOL, f32_equal_bits, ms_blocks, tone_s16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(ak):
    import luahost
    h = luahost.LuaHost()
    yield h
    h.close()


def _top(host):
    return host.lib.lua_gettop(host.L)


def test_module_table_matches_the_source_and_the_shim(host, ak):
    src = open(os.path.join(ROOT, "aukit_b200", "csrc", "lua_binding.c")).read()
    table = src[src.index("static const luaL_Reg funcs[]"):]
    table = table[:table.index("{NULL, NULL}")]
    declared = re.findall(r'\{"(\w+)",\s*l_\w+\}', table)
    names = host.names()
    assert sorted(names) == sorted(declared + ["abi_version"])
    assert host.field("abi_version") == float(ak._lib.load().aukit_cuda_abi_version())
    shim = open(os.path.join(ROOT, "tests", "luashim.py")).read()
    for n in declared:
        if n in ("preload", "device_count", "adpcm", "ima_adpcm_wav", "msadpcm", "concat", "write"):
            continue                                   # only the C binding offers these
        assert '"%s"' % n in shim, n


def test_argument_errors_unwind_cleanly(host):
    import luahost
    for name, args, msg in (
            ("pcm", (), r"bad argument #1 to '\?' \(string expected, got no value\)"),
            ("pcm", (None,), r"bad argument #1 to '\?' \(string expected, got nil\)"),
            ("pcm", (b"\0\0", "x"), r"bad argument #2 to '\?' \(number expected, got string\)"),
            ("resample", (b"not an audio", 48000.0, 1), r"bad argument #1 to '\?' \(aukit_cuda.Audio expected, got string\)"),
            ("mono", (), r"aukit_cuda.Audio expected, got no value"),
            ("ima_adpcm_wav", (b"\0" * 8,), r"bad argument #2 to '\?' \(number expected, got no value\)"),
            ("nope", (), r"attempt to call field 'nope'")):
        with pytest.raises(luahost.HostError, match=msg):
            host.call(name, *args)
        assert _top(host) == 0                          # the protected call left the stack as it found it
    # numbers are accepted where strings are expected, and numeric strings as numbers (lua_tolstring / lua_tonumberx)
    try:
        got = host.call("pcm", 12.0, "16")                 # "12" = one 16-bit mono frame: succeeds where a GPU exists
        assert len(got) == 1
    except luahost.HostError as e:                          # ... and reaches the library (no device here), not an argument error
        assert "bad argument" not in str(e) and "no CPU fallback" in str(e)
    assert _top(host) == 0


@pytest.mark.gpu
def test_auplay_chain_through_the_c_binding(host, ak, O):
    pcm = tone_s16(6 * 8000 + 77, 2, 44100, seed=12)
    h = host.call("pcm", pcm.tobytes(), 16, 0, 2, 44100.0)[0]
    assert host.call("channels", h) == [2.0] and host.call("frames", h) == [float(len(pcm))] and host.call("sample_rate", h) == [44100.0]
    r = host.call("resample", h, 48000.0, 2)[0]
    m = host.call("mono", r)[0]
    assert host.call("normalize", m, 0.8) == []
    n = int(host.call("frames", m)[0])
    got = np.array(host.call("read", m, 1, 1, n)[0]["arr"], dtype=np.float32)
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= TOL
    via_ctypes = ak.effects.normalize(ak.pcm(pcm.tobytes(), 16, "signed", 2, 44100).resample(48000, "cubic").mono(), 0.8).numpy()[0]
    assert f32_equal_bits(got, via_ctypes)
    # the library's own error strings arrive as Lua errors
    import luahost
    with pytest.raises(luahost.HostError, match=r"bad argument #2 \(invalid bit depth\)"):
        host.call("pcm", b"\0\0", 12)
    with pytest.raises(luahost.HostError, match="frame range out of bounds"):
        host.call("read", m, 1, n, 5)
    assert _top(host) == 0


@pytest.mark.gpu
def test_userdata_gc_runs_the_finaliser(host):
    import gc
    gc.collect()
    live0, calls0 = host.live_userdata, host.gc_calls
    hs = [host.call("new", 2, 1000, 8000.0)[0] for _ in range(5)]
    assert host.live_userdata == live0 + 5
    extra = host.call("mono", hs[0])[0]
    hs[1].release()                                     # explicit: last reference gone -> __gc -> aukit_cuda_audio_free
    assert host.gc_calls == calls0 + 1 and host.live_userdata == live0 + 5
    import luahost
    with pytest.raises(luahost.HostError, match="userdata already released"):
        host.call("channels", hs[1])
    del hs, extra                                       # Python drops the references like Lua's collector would
    gc.collect()
    assert host.live_userdata == live0 and host.gc_calls == calls0 + 6
    assert _top(host) == 0


@pytest.mark.gpu
def test_tables_in_and_out(host, ak, O):
    # wav: nested result tables (info + tags)
    import struct
    pcm = tone_s16(500, 1, 8000, seed=2)
    info = b"INFO" + b"INAM" + struct.pack("<I", 4) + b"Song"
    body = b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8000, 16000, 2, 16) + b"LIST" + struct.pack("<I", len(info)) + info + \
        b"data" + struct.pack("<I", pcm.nbytes) + pcm.tobytes()
    wav = b"RIFF" + struct.pack("<I", len(body)) + body
    a, t = host.call("wav", wav)
    assert t["fields"][b"dataType"] == b"signed" and t["fields"][b"channels"] == 1.0 and t["fields"][b"bitDepth"] == 16.0
    assert t["fields"][b"tags"]["arr"][0]["arr"] == [b"INAM", b"Song"]
    got = np.array(host.call("read", a, 1, 1, 500)[0]["arr"], dtype=np.float32)
    assert f32_equal_bits(got, ak.wav(wav).numpy()[0])
    # msadpcm: coefficient tables travel as Lua arrays (int_table), dialect as a number
    raw = ms_blocks(6, 256, 2, seed=5)
    c1, c2 = [256, 512, 0, 192, 240, 460, 392], [0, -256, 0, 64, 0, -208, -232]
    m = host.call("msadpcm", raw.tobytes(), 256, 2, 22050.0, c1, c2, 0)[0]
    ref = O.msadpcm(raw, 256, 2, None, 0).astype(np.float32)
    n = int(host.call("frames", m)[0])
    for c in range(2):
        got = np.array(host.call("read", m, c + 1, 1, n)[0]["arr"], dtype=np.float32)
        assert f32_equal_bits(got, ref[c])
    # write (Lua array in) + stream_chunk (array of arrays out) + concat (varargs) + pcm_bytes (string out)
    z = host.call("new", 1, 8, 8000.0)[0]
    host.call("write", z, 1, 3, [0.5, -0.25, 1.0])
    assert host.call("read", z, 1, 1, 8)[0]["arr"] == [0, 0, 0.5, -0.25, 1.0, 0, 0, 0]
    both = host.call("concat", z, z)[0]
    assert host.call("frames", both) == [16.0]
    chunk = host.call("stream_chunk", both, 8, 0, 9, 100)[0]
    assert len(chunk["arr"]) == 1 and chunk["arr"][0]["arr"] == [0, 0, 63.5, -32, 127, 0, 0, 0]
    assert host.call("stream_chunk", both, 8, 0, 17, 4) == [None]
    assert host.call("pcm_bytes", z, 16, 0, True, 1)[0] == np.array([0, 0, 16383, -8192, 32767, 0, 0, 0], "<i2").tobytes()
    assert _top(host) == 0
