"""TEST INFRASTRUCTURE: runs the real Lua C module (aukit_b200/lib/aukit_cuda.so, built from csrc/lua_binding.c)
inside tests/luahost/luahost.c, a toy host for the Lua 5.2 C API (there is no Lua interpreter in the image).

    host = LuaHost()                       # builds + loads libluahost.so (RTLD_GLOBAL), dlopen()s aukit_cuda.so,
                                           # calls luaopen_aukit_cuda
    host.call("pcm", data, 16, 0, 2, 44100.0)   -> [UserData]
    host.module()                          # the same functions as a luavm LuaTable: `require "aukit_cuda"` for the
                                           # Lua facade running in oracle/luavm

Values cross as: None <-> nil, bool, float/int <-> number, bytes <-> string, list / luavm LuaTable <-> table (array part
and string keys), UserData <-> full userdata (kept alive by a host reference; dropping the Python object releases it,
which runs the binding's __gc)."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "luahost", "luahost.c")
LIB = os.path.join(HERE, "luahost", "libluahost.so")
MODULE = os.path.join(ROOT, "aukit_b200", "lib", "aukit_cuda.so")

LUA_TNONE, LUA_TNIL, LUA_TBOOLEAN, LUA_TNUMBER, LUA_TSTRING, LUA_TTABLE, LUA_TFUNCTION, LUA_TUSERDATA = -1, 0, 1, 3, 4, 5, 6, 7


class HostError(Exception):
    """A Lua error raised by the C module (lua_error / luaL_error / a failed luaL_check*)."""


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        r = subprocess.run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-o", LIB, SRC, "-ldl"], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("luahost build failed:\n" + r.stdout + r.stderr)
    return LIB


class UserData:
    """A full userdata living in the host; the Python object is the reference a Lua variable would be."""

    def __init__(self, host, ref):
        self._host, self._ref = host, ref

    def release(self):
        if self._ref is not None and self._host.L:
            self._host.lib.lh_unref(self._host.L, self._ref)
        self._ref = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class LuaHost:
    def __init__(self, module_path=MODULE, openf="luaopen_aukit_cuda"):
        lib = C.CDLL(build(), mode=C.RTLD_GLOBAL)          # the module's lua_* / luaL_* symbols resolve here
        P, I, D, SZ = C.c_void_p, C.c_int, C.c_double, C.c_size_t
        for name, res, args in (
                ("lh_new", P, []), ("lh_close", None, [P]), ("lh_error", C.c_char_p, [P]), ("lh_gc_calls", C.c_long, [P]),
                ("lh_live_userdata", C.c_long, [P]), ("lh_open", I, [P, C.c_char_p, C.c_char_p]), ("lh_getmodule", I, [P, C.c_char_p]),
                ("lh_module_key", C.c_char_p, [P, I]), ("lh_call", I, [P, C.c_char_p, I]), ("lh_protected", I, [P, I, I, I, C.c_char_p]),
                ("lh_field", C.c_char_p, [P, I, I]), ("lh_ref", I, [P, I]), ("lh_pushref", I, [P, I]), ("lh_unref", None, [P, I]),
                ("lua_gettop", I, [P]), ("lua_type", I, [P, I]), ("lua_toboolean", I, [P, I]), ("lua_tonumberx", D, [P, I, P]),
                ("lua_tolstring", P, [P, I, C.POINTER(SZ)]), ("lua_rawlen", SZ, [P, I]), ("lua_pushnil", None, [P]),
                ("lh_array_numbers", C.c_long, [P, I, C.POINTER(D), C.c_long]), ("lh_push_number_array", I, [P, C.POINTER(D), C.c_long]),
                ("lua_pushnumber", None, [P, D]), ("lua_pushboolean", None, [P, I]), ("lua_pushlstring", P, [P, C.c_char_p, SZ])):
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        self.lib = lib
        self.L = lib.lh_new()
        if lib.lh_open(self.L, module_path.encode(), openf.encode()) != 0:
            raise HostError(lib.lh_error(self.L).decode("latin-1"))

    # ---- stack helpers
    def _prot(self, op, idx=0, n=0, key=None):
        if self.lib.lh_protected(self.L, op, idx, n, key) != 0:
            raise HostError(self.lib.lh_error(self.L).decode("latin-1"))

    def settop(self, n):
        self._prot(0, n)

    def push(self, v):
        lib, L = self.lib, self.L
        if v is None:
            lib.lua_pushnil(L)
        elif isinstance(v, bool):
            lib.lua_pushboolean(L, int(v))
        elif isinstance(v, (int, float)):
            lib.lua_pushnumber(L, float(v))
        elif isinstance(v, (bytes, bytearray)):
            lib.lua_pushlstring(L, bytes(v), len(v))
        elif isinstance(v, str):
            b = v.encode("latin-1")
            lib.lua_pushlstring(L, b, len(b))
        elif isinstance(v, UserData):
            if v._ref is None or lib.lh_pushref(L, v._ref) != 0:
                raise HostError("userdata already released")
        elif isinstance(v, (list, tuple)) and len(v) >= 16 and all(type(x) in (int, float) for x in v):
            arr = (C.c_double * len(v))(*v)
            if lib.lh_push_number_array(L, arr, len(v)) != 0:
                raise HostError("stack overflow")
        elif isinstance(v, (list, tuple)):
            self._prot(5, 0, len(v))
            for i, x in enumerate(v):
                self.push(x)
                self._prot(1, -2, i + 1)
        elif hasattr(v, "arr") and hasattr(v, "hash"):          # a luavm LuaTable
            if len(v.arr) >= 16 and not v.hash and all(type(x) in (int, float) for x in v.arr):
                return self.push(list(v.arr))
            self._prot(5, 0, len(v.arr))
            for i, x in enumerate(v.arr):
                self.push(x)
                self._prot(1, -2, i + 1)
            for k, x in v.hash.items():
                self.push(x)
                if isinstance(k, (bytes, str)):
                    self._prot(3, -2, 0, k if isinstance(k, bytes) else k.encode())
                elif isinstance(k, (int, float)) and int(k) == k and k >= 1:
                    self._prot(1, -2, int(k))
                else:
                    raise HostError("table key %r cannot cross into the host" % (k,))
        else:
            raise HostError("cannot push %r" % type(v))

    def value(self, idx, table=None):
        """The value at stack index idx as a Python object (tables become `table()` instances or dicts)."""
        lib, L = self.lib, self.L
        t = lib.lua_type(L, idx)
        if t in (LUA_TNIL, LUA_TNONE):
            return None
        if t == LUA_TBOOLEAN:
            return bool(lib.lua_toboolean(L, idx))
        if t == LUA_TNUMBER:
            return lib.lua_tonumberx(L, idx, None)
        if t == LUA_TSTRING:
            n = C.c_size_t(0)
            p = lib.lua_tolstring(L, idx, C.byref(n))
            return C.string_at(p, n.value)
        if t == LUA_TUSERDATA:
            return UserData(self, lib.lh_ref(L, idx))
        if t == LUA_TTABLE:
            if idx < 0:
                idx = lib.lua_gettop(L) + idx + 1
            arr = []
            n = lib.lua_rawlen(L, idx)
            got = -1
            if n >= 16:                                            # long arrays of numbers cross in one call
                buf = (C.c_double * n)()
                got = lib.lh_array_numbers(L, idx, buf, n)
                if got == n:
                    arr = list(buf)
            if got != n:
                for i in range(1, n + 1):
                    self._prot(2, idx, i)
                    arr.append(self.value(-1, table))
                    self.settop(-2)
            fields = {}
            i = 0
            while True:
                k = lib.lh_field(L, idx, i)
                if k is None:
                    break
                fields[k] = self.value(-1, table)
                self.settop(-2)
                i += 1
            if table is None:
                return {"arr": arr, "fields": fields}
            t_ = table()
            t_.arr = arr
            for k, x in fields.items():
                t_.set(k, x)
            return t_
        raise HostError("unsupported Lua type %d on the stack" % t)

    # ---- calls
    def call(self, name, *args, table=None):
        base = self.lib.lua_gettop(self.L)
        for a in args:
            self.push(a)
        n = self.lib.lh_call(self.L, name.encode(), len(args))
        if n < 0:
            msg = self.lib.lh_error(self.L)
            self.settop(base)
            raise HostError(msg.decode("latin-1"))
        out = [self.value(base + 1 + i, table) for i in range(n)]
        self.settop(base)
        return out

    def names(self):
        out, i = [], 0
        while True:
            k = self.lib.lh_module_key(self.L, i)
            if k is None:
                return out
            out.append(k.decode())
            i += 1

    def field(self, name):
        self.lib.lh_getmodule(self.L, name.encode())
        v = self.value(-1)
        self.settop(-2)
        return v

    @property
    def gc_calls(self):
        return int(self.lib.lh_gc_calls(self.L))

    @property
    def live_userdata(self):
        return int(self.lib.lh_live_userdata(self.L))

    def module(self):
        """`require "aukit_cuda"` for oracle/luavm: every C function of the module as a luavm function."""
        from oracle.luavm.lua import LuaError, LuaFunction, LuaTable
        mod = LuaTable()
        for name in self.names():
            self.lib.lh_getmodule(self.L, name.encode())
            t = self.lib.lua_type(self.L, -1)
            if t == LUA_TFUNCTION:
                self.settop(-2)

                def f(a, _n=name):
                    try:
                        return self.call(_n, *a, table=LuaTable)
                    except HostError as e:
                        raise LuaError(str(e).encode("latin-1"))
                mod.set(name.encode(), LuaFunction(f, "aukit_cuda." + name))
            else:
                mod.set(name.encode(), self.value(-1, LuaTable))
                self.settop(-2)
        return mod

    def close(self):
        if self.L:
            self.lib.lh_close(self.L)
            self.L = None
