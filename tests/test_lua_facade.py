"""The Lua facade (aukit_b200/lua/aukit.lua) executed by oracle/luavm against the real CUDA library: the
auplay.lua call chain (auplay.lua:12-27) written in Lua, unchanged from how a ComputerCraft script would
write it, must reproduce the reference's own golden output.  Every test runs twice: with `require "aukit_cuda"`
served by the REAL C binding (csrc/lua_binding.c -> lib/aukit_cuda.so, executed inside tests/luahost/luahost.c, a toy
host for the Lua 5.2 C API -- no Lua interpreter that could dlopen() it exists in the image), and by tests/luashim.py,
a Python stand-in with the same function table that reaches the library through ctypes."""
import json
import os

import numpy as np
import pytest

from util import TOL, tone_s16

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

AUPLAY_CHAIN = r'''
local aukit = require "aukit"
local data = ...
local audio = aukit.wav(data)                    -- auplay.lua:12
local resamp = audio:resample(48000)             -- auplay.lua:21 (aukit.defaultInterpolation = "linear")
local mono = resamp:mono()                       -- auplay.lua:24
local ret = aukit.effects.normalize(mono, 0.8)   -- auplay.lua:27: relies on in-place mutation
assert(ret == mono)
local out = {}
for i = 1, #mono.data[1] do out[i] = mono.data[1][i] end
return out, #mono.data, mono:len(), mono.sampleRate, audio.info.dataType, audio.info.bitDepth, #audio.data, #audio.data[1]
'''


@pytest.fixture(scope="module", params=["cbinding", "shim"])
def lua(ak, request):
    from oracle.luavm.aukit_ref import EXPECT_LUA
    from oracle.luavm.lua import Interpreter
    I = Interpreter()
    I.preload[b"cc.expect"] = lambda: I.run(EXPECT_LUA, "cc.expect")[0]
    if request.param == "cbinding":
        import luahost
        host = luahost.LuaHost()
        I.preload[b"aukit_cuda"] = host.module
    else:
        import luashim
        I.preload[b"aukit_cuda"] = lambda: luashim.make_module(ak)
    src = open(os.path.join(ROOT, "aukit_b200", "lua", "aukit.lua"), "rb").read()
    I.preload[b"aukit"] = lambda: I.run(src, "aukit.lua(facade)")[0]
    return I


def test_auplay_chain_through_the_lua_facade(lua):
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
    manifest = json.loads(z["manifest"].tobytes().decode())
    i = [m["name"] for m in manifest].index("chain_c1_mini_linear")
    wav = z["c%d/in" % i].tobytes()
    ref = z["c%d/out0" % i]                          # what the reference's aukit.lua produced for the same chain
    r = lua.run(AUPLAY_CHAIN, "auplay_chain", [wav])
    got = np.array(r[0].arr)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= TOL
    assert r[1:] == [1.0, len(ref) / 48000.0, 48000.0, b"signed", 16.0, 2.0, 8820.0]


def test_facade_errors_and_effects_semantics(lua):
    from oracle.luavm.lua import LuaError
    with pytest.raises(LuaError, match=r"bad argument #2 \(invalid bit depth\)"):
        lua.run('local aukit = require "aukit" return aukit.pcm("\\0\\0", 12)')
    with pytest.raises(LuaError, match=r"uneven amount of data per channel"):
        lua.run('local aukit = require "aukit" return aukit.pcm("\\0\\0\\0\\0\\0\\0", 16, "signed", 2)')
    with pytest.raises(LuaError, match=r"invalid interpolation type"):
        lua.run('local aukit = require "aukit" return aukit.pcm("\\0\\0\\0\\0", 16):resample(48000, "bogus")')
    with pytest.raises(LuaError, match=r"not a WAV file"):
        lua.run('local aukit = require "aukit" return aukit.wav("RIFXxxxxxxxxxxxxxxxx")')
    r = lua.run('''
        local aukit = require "aukit"
        local a = aukit.pcm("\\0\\64\\0\\192", 16, "signed", 1, 8000)       -- 0.5, -0.5
        local b = aukit.effects.amplify(a, 1)                               -- multiplier 1: untouched, same object
        local c = aukit.effects.amplify(a, 4)                               -- in place, clamped
        return b == a, c == a, a.data[1][1], a.data[1][2], #a.data[1], a:len(), a:channels()
    ''')
    assert r == [True, True, 1.0, -1.0, 2.0, 2 / 8000.0, 1.0]


def test_facade_next_rows_lowpass_pcm_containers(lua, O):
    """SURVEY 8(f) rows through the Lua facade: auplay.lua:27-34 continues with effects.lowpass and the
    requantisation of Audio:pcm; aukit.au / aukit.aiff land on the same loaders."""
    import struct
    pcm = tone_s16(2000, 1, 8000, seed=8)
    au = b".snd" + struct.pack(">IIIII", 24, pcm.nbytes, 3, 8000, 1) + pcm.astype(">i2").tobytes()
    lua.G.set(b"AU_FILE", au)
    r = lua.run('''
        local aukit = require "aukit"
        local a = aukit.au(AU_FILE)
        local rate, ch, n = a.sampleRate, a:channels(), #a.data[1]
        aukit.effects.normalize(a, 0.8)
        local same = aukit.effects.lowpass(a, a.sampleRate / 2) == a
        local q = a:pcm(8, "signed", true)
        return rate, ch, n, same, q, a.info.bitDepth, a.info.dataType
    ''')
    ref_dec, info = O.au(au)
    ref = O.audio_pcm(O.lowpass(O.normalize(ref_dec, 0.8, False), 4000.0, 8000.0), 8, "signed", True)
    assert r[0] == 8000.0 and r[1] == 1.0 and r[2] == ref_dec.shape[1] and r[3] is True
    got = np.array(r[4].arr)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 128 * 2 * TOL
    assert r[5:] == [16.0, b"signed"]


def test_facade_audio_wav_writer(lua, ak):
    """Audio:wav through the facade (header packed by the host Lua, samples by the GPU): the same bytes as the
    Python mirror's general dialect, and a file the loader reads back to within one LSB."""
    pcm = tone_s16(3000, 2, 8000, seed=9)
    lua.G.set(b"PCM_BYTES", pcm.tobytes())
    r = lua.run('''
        local aukit = require "aukit"
        local a = aukit.pcm(PCM_BYTES, 16, "signed", 2, 8000)
        a.metadata.title = "T"
        return a:wav(16)
    ''')
    a = ak.pcm(pcm.tobytes(), 16, "signed", 2, 8000)
    a.metadata = {"title": "T"}
    assert r[0] == a.wav(16, "floor", ak.DIALECT_GENERAL)
    back = ak.wav(r[0])
    assert back.channels() == 2 and back.frames == 3000 and back.metadata == {"title": b"T"} or back.metadata == {"title": "T"}
    assert np.max(np.abs(back.numpy() - a.numpy())) <= 1.0 / 32767


def test_facade_preload_is_the_four_reference_calls(lua, O):
    """aukit.preload (the fused chain; through the C binding it also takes the all-GPUs path when the box has several)
    against the same chain written with the reference's four calls, and against the oracle."""
    pcm = tone_s16(5 * 44100 + 321, 2, 44100, seed=21)
    lua.G.set(b"PCM_BYTES", pcm.tobytes())
    r = lua.run('''
        local aukit = require "aukit"
        local fused = aukit.preload(PCM_BYTES, 16, "signed", 2, 44100, 48000, "cubic", true, 0.8)
        local four = aukit.pcm(PCM_BYTES, 16, "signed", 2, 44100):resample(48000, "cubic"):mono()
        aukit.effects.normalize(four, 0.8)
        local n, worst = #fused.data[1], 0
        for i = 1, n do worst = math.max(worst, math.abs(fused.data[1][i] - four.data[1][i])) end
        local out = {}
        for i = 1, n do out[i] = fused.data[1][i] end
        return n, #four.data[1], worst, fused.sampleRate, fused:channels(), aukit.deviceCount(), out
    ''')
    ref = O.chain_s16(pcm.tobytes(), 2, 44100, 48000, "cubic", 0.8)
    assert r[0] == r[1] == float(len(ref)) and r[2] <= 2 * TOL and r[3] == 48000.0 and r[4] == 1.0 and r[5] >= 1.0
    assert np.max(np.abs(np.array(r[6].arr) - ref)) <= TOL
    from oracle.luavm.lua import LuaError
    with pytest.raises(LuaError, match=r"uneven amount of data per channel"):
        lua.run('local aukit = require "aukit" return aukit.preload("\\0\\0\\0", 16, "signed", 2)')
