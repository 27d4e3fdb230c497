"""Runs the UNMODIFIED reference /root/reference/aukit.lua inside oracle/luavm (TEST INFRASTRUCTURE).

Only usable where /root/reference exists (this container).  tests/golden/generate.py uses it to
produce the committed golden vectors; nothing on the GPU box imports it.

Shims: cc.expect (callable table with .range), a stub cc.audio.dfpwm,
os.epoch (constant, so the reference's 3 s yield never fires), sleep (no-op).  string.pack /
string.unpack, bit32, utf8-free paths are provided by the interpreter itself.
"""
from __future__ import annotations

import os

import numpy as np

from .lua import Interpreter, LuaError, LuaFunction, LuaTable, call, index, to_lua, tostr, type_name

REFERENCE = os.environ.get("AUKIT_REFERENCE", "/root/reference/aukit.lua")

EXPECT_LUA = r'''
-- minimal cc.expect: expect(index, value, ...types) / expect.range(num, min, max)
local function expect(index, value, ...)
    local t = type(value)
    local n = select("#", ...)
    for i = 1, n do
        local want = select(i, ...)
        if t == want then return value end
    end
    local types = {...}
    local name
    if n > 1 then name = table.concat(types, ", ", 1, n - 1) .. " or " .. types[n] else name = types[1] end
    error(("bad argument #%d (expected %s, got %s)"):format(index, name, t), 3)
end
local function range(num, min, max)
    min = min or -math.huge
    max = max or math.huge
    if num ~= num or num < min or num > max then
        error(("number outside of range (expected %s to be within %s and %s)"):format(num, min, max), 3)
    end
    return num
end
return setmetatable({expect = expect, range = range, field = function() end}, {__call = function(_, ...) return expect(...) end})
'''


class Reference:
    def __init__(self, path: str = REFERENCE):
        self.I = Interpreter()
        I = self.I
        I.preload[b"cc.expect"] = lambda: I.run(EXPECT_LUA, "cc.expect")[0]

        def dfpwm_stub():
            t = LuaTable()
            for k in (b"make_decoder", b"make_encoder", b"encode", b"decode"):
                t.set(k, LuaFunction(lambda a: (_ for _ in ()).throw(LuaError(b"dfpwm is outside the hot path")), "dfpwm"))
            return t
        I.preload[b"cc.audio.dfpwm"] = dfpwm_stub
        src = open(path, "rb").read()
        self.aukit = I.run(src, "aukit.lua")[0]

    # --- generic call helpers
    def fn(self, *path):
        o = self.aukit
        for p in path:
            o = index(o, p.encode() if isinstance(p, str) else p)
        return o

    def call(self, path, *args):
        f = self.fn(*path) if isinstance(path, (tuple, list)) else self.fn(path)
        return call(f, [to_lua(a) for a in args])

    def method(self, obj, name, *args):
        return call(index(obj, name.encode()), [obj] + [to_lua(a) for a in args])

    # --- Audio -> numpy
    @staticmethod
    def audio_data(audio):
        data = audio.get(b"data")
        chans = []
        for ch in data.arr:
            chans.append(np.array([np.nan if v is None else v for v in ch.arr], dtype=np.float64))
        return chans

    @staticmethod
    def audio_fields(audio):
        def conv(t):
            out = {}
            if type(t) is LuaTable:
                k = None
                while True:
                    r = t.next(k)
                    if r[0] is None:
                        break
                    k = r[0]
                    key = k.decode("latin-1") if type(k) is bytes else k
                    v = r[1]
                    out[key] = v.decode("latin-1") if type(v) is bytes else v
            return out
        return {"sampleRate": audio.get(b"sampleRate"), "metadata": conv(audio.get(b"metadata")), "info": conv(audio.get(b"info"))}


def lua_error_message(ex: LuaError) -> str:
    return str(ex)
