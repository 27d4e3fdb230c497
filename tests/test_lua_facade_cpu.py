"""The Lua facade over the REAL C binding without a GPU: `require "aukit"` loads (luaopen_aukit_cuda runs inside
tests/luahost), the argument checks that the facade performs itself raise the reference's messages before any device
work, and the first call that needs the device fails loudly through all three layers (Lua facade -> lua_binding.c ->
libaukit_cuda.so) -- there is no CPU fallback to fall into."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lua(ak):
    from oracle.luavm.aukit_ref import EXPECT_LUA
    from oracle.luavm.lua import Interpreter
    import luahost
    host = luahost.LuaHost()
    I = Interpreter()
    I.preload[b"cc.expect"] = lambda: I.run(EXPECT_LUA, "cc.expect")[0]
    I.preload[b"aukit_cuda"] = host.module
    src = open(os.path.join(ROOT, "aukit_b200", "lua", "aukit.lua"), "rb").read()
    I.preload[b"aukit"] = lambda: I.run(src, "aukit.lua(facade)")[0]
    return I


def test_facade_loads_and_checks_arguments_without_a_device(lua):
    from oracle.luavm.lua import LuaError
    r = lua.run('local aukit = require "aukit" return aukit._VERSION, aukit.defaultInterpolation, type(aukit.effects.normalize), type(aukit.preload)')
    assert r[0].startswith(b"1.10.0") and r[1:] == [b"linear", b"function", b"function"]
    for code, msg in (
            ('aukit.pcm("\\0\\0", 12)', r"bad argument #2 \(invalid bit depth\)"),
            ('aukit.pcm("\\0\\0\\0\\0", 32, "float", 1, 0)', r"number outside of range"),
            ('aukit.pcm("\\0\\0\\0", 16, "signed", 2)', r"uneven amount of data per channel"),
            ('aukit.pcm(42)', r"bad argument #1 \(expected string or table, got number\)"),
            ('aukit.preload("\\0\\0\\0", 16, "signed", 2)', r"uneven amount of data per channel"),
            ('aukit.preload("\\0\\0\\0\\0", 16, "signed", 2, 44100, 48000, "bogus")', r"invalid interpolation type"),
            ('aukit.wav("RIFXxxxxxxxxxxxxxxxx")', r"not a WAV file"),
            ('aukit.effects.normalize({}, 1)', r"expected Audio")):
        with pytest.raises(LuaError, match=msg):
            lua.run('local aukit = require "aukit" return ' + code)


def test_device_calls_fail_loudly_through_the_lua_boundary(lua):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oracle.luavm.lua import LuaError
    for code in ('aukit.pcm("\\0\\0\\0\\0", 16, "signed", 1, 8000)', 'aukit.new(1, 1, 8000)',
                 'aukit.preload("\\0\\0\\0\\0", 16, "signed", 2, 44100, 48000, "cubic")'):
        with pytest.raises(LuaError, match="no CPU fallback"):
            lua.run('local aukit = require "aukit" return ' + code)
