"""A small Lua 5.2 interpreter -- just enough of the language and standard library to execute
the UNMODIFIED reference /root/reference/aukit.lua (which needs Lua 5.2 + string.pack/unpack +
bit32 + a few ComputerCraft shims).

TEST INFRASTRUCTURE ONLY (lives under oracle/).  No Lua interpreter exists in this image, so this
is how the reference itself is run here to generate the golden vectors under tests/golden/
(tests/golden/generate.py).  It is a from-scratch implementation of the Lua 5.2 reference
manual's semantics: all numbers are IEEE doubles, strings are byte strings, tables have
metatables, functions are closures with upvalues, multiple assignment / results / varargs.

Design: source -> tokens -> AST (nested tuples) -> Python closures.  Every local variable lives
in a one-element list (a cell) held in the activation's frame list, so closures capture cells
exactly like Lua upvalues.  A Lua function is a Python callable taking a list of arguments and
returning a list of results.
"""
from __future__ import annotations

import math
import struct
import sys
import time

sys.setrecursionlimit(10000)


class LuaError(Exception):
    def __init__(self, value, traceback=None):
        Exception.__init__(self, value)
        self.value = value
        self.lua_traceback = traceback

    def __str__(self):
        v = self.value
        return v.decode("latin-1") if isinstance(v, bytes) else str(v)


# ======================================================================================= values
class LuaTable:
    __slots__ = ("arr", "hash", "meta")

    def __init__(self):
        self.arr = []      # values for keys 1..len(arr)
        self.hash = {}
        self.meta = None

    def get(self, k):
        if type(k) is float:
            ik = int(k)
            if ik == k:
                if 1 <= ik <= len(self.arr):
                    return self.arr[ik - 1]
                return self.hash.get(ik)
            return self.hash.get(k)
        if type(k) is int:
            if 1 <= k <= len(self.arr):
                return self.arr[k - 1]
        return self.hash.get(k)

    def set(self, k, v):
        if type(k) is float:
            ik = int(k)
            if ik == k:
                k = ik
            elif k != k:
                raise LuaError(b"table index is NaN")
        elif k is None:
            raise LuaError(b"table index is nil")
        if type(k) is int:
            n = len(self.arr)
            if 1 <= k <= n:
                self.arr[k - 1] = v
                if v is None and k == n:
                    arr = self.arr
                    while arr and arr[-1] is None:
                        arr.pop()
                return
            if k == n + 1:
                if v is None:
                    self.hash.pop(k, None)
                    return
                self.arr.append(v)
                h = self.hash
                if h:
                    h.pop(k, None)
                    nk = k + 1
                    while nk in h:
                        self.arr.append(h.pop(nk))
                        nk += 1
                return
        if v is None:
            self.hash.pop(k, None)
        else:
            self.hash[k] = v

    def length(self):
        return len(self.arr)

    def next(self, k):
        """Lua's next(): iteration order = array part, then hash part."""
        n = len(self.arr)
        if k is None:
            i = 0
        else:
            if type(k) is float and int(k) == k:
                k = int(k)
            if type(k) is int and 1 <= k <= n:
                i = k
            else:
                keys = list(self.hash.keys())
                try:
                    pos = keys.index(k)
                except ValueError:
                    raise LuaError(b"invalid key to 'next'")
                if pos + 1 < len(keys):
                    nk = keys[pos + 1]
                    return [float(nk) if type(nk) is int else nk, self.hash[nk]]
                return [None]
        while i < n:
            if self.arr[i] is not None:
                return [float(i + 1), self.arr[i]]
            i += 1
        for nk, v in self.hash.items():
            return [float(nk) if type(nk) is int else nk, v]
        return [None]


class LuaFunction:
    __slots__ = ("call", "name")

    def __init__(self, call, name="?"):
        self.call = call
        self.name = name


def type_name(v):
    if v is None:
        return "nil"
    if v is True or v is False:
        return "boolean"
    if type(v) is float or type(v) is int:
        return "number"
    if type(v) is bytes:
        return "string"
    if type(v) is LuaTable:
        return "table"
    if type(v) is LuaFunction:
        return "function"
    return "userdata"


def fmt_number(x):
    if x != x:
        return b"nan" if math.copysign(1, x) > 0 else b"-nan"
    if x in (math.inf, -math.inf):
        return b"inf" if x > 0 else b"-inf"
    return ("%.14g" % x).encode()


def tostr(v):
    if type(v) is bytes:
        return v
    if type(v) is float or type(v) is int:
        return fmt_number(float(v))
    if v is None:
        return b"nil"
    if v is True:
        return b"true"
    if v is False:
        return b"false"
    if type(v) is LuaTable:
        mt = v.meta
        if mt is not None:
            h = mt.get(b"__tostring")
            if h is not None:
                return call(h, [v])[0]
        return ("table: 0x%08x" % (id(v) & 0xFFFFFFFF)).encode()
    if type(v) is LuaFunction:
        return ("function: 0x%08x" % (id(v) & 0xFFFFFFFF)).encode()
    return repr(v).encode()


def str2number(s):
    try:
        t = s.decode("latin-1").strip(" \t\n\r\f\v")
    except Exception:
        return None
    if not t:
        return None
    try:
        low = t.lower()
        body = low.lstrip("+-")
        if body.startswith("0x"):
            neg = low.startswith("-")
            if "." in body or "p" in body:
                v = float.fromhex(body)
            else:
                v = float(int(body, 16))
            return -v if neg else v
        if body in ("inf", "infinity", "nan") or "_" in body:
            return None
        return float(t)
    except ValueError:
        return None


def tonum(v):
    if type(v) is float:
        return v
    if type(v) is int:
        return float(v)
    if type(v) is bytes:
        return str2number(v)
    return None


def call(f, args):
    if type(f) is LuaFunction:
        return f.call(args)
    if type(f) is LuaTable and f.meta is not None:
        h = f.meta.get(b"__call")
        if h is not None:
            return call(h, [f] + args)
    raise LuaError(("attempt to call a %s value" % type_name(f)).encode())


def index(o, k):
    t = type(o)
    if t is LuaTable:
        v = o.get(k)
        if v is not None or o.meta is None:
            return v
        h = o.meta.get(b"__index")
        if h is None:
            return None
        if type(h) is LuaFunction:
            r = h.call([o, k])
            return r[0] if r else None
        return index(h, k)
    if t is bytes:
        return STRING_LIB.get(k)
    raise LuaError(("attempt to index a %s value" % type_name(o)).encode())


def setindex(o, k, v):
    if type(o) is LuaTable:
        if o.meta is not None:
            h = o.meta.get(b"__newindex")
            if h is not None and o.get(k) is None:
                if type(h) is LuaFunction:
                    h.call([o, k, v])
                else:
                    setindex(h, k, v)
                return
        o.set(k, v)
        return
    raise LuaError(("attempt to index a %s value" % type_name(o)).encode())


def arith_meta(op, a, b):
    for x in (a, b):
        if type(x) is LuaTable and x.meta is not None:
            h = x.meta.get(op)
            if h is not None:
                r = call(h, [a, b])
                return r[0] if r else None
    bad = b if tonum(a) is not None else a
    if op == b"__concat":
        raise LuaError(("attempt to concatenate a %s value" % type_name(bad)).encode())
    raise LuaError(("attempt to perform arithmetic on a %s value" % type_name(bad)).encode())


def arith(op, a, b):
    x, y = tonum(a), tonum(b)
    if x is None or y is None:
        return arith_meta(op, a, b)
    return ARITH[op](x, y)


def lua_div(x, y):
    try:
        return x / y
    except ZeroDivisionError:
        if x != x or x == 0:
            return math.nan
        neg = (x < 0) != (math.copysign(1.0, y) < 0)
        return -math.inf if neg else math.inf


def lua_mod(x, y):
    # luai_nummod: a - floor(a/b)*b
    try:
        return x - math.floor(x / y) * y
    except (ZeroDivisionError, OverflowError, ValueError):
        if y == 0 or x != x or y != y or x in (math.inf, -math.inf):
            return math.nan
        return x  # finite % inf (sign cases ignored)


def lua_pow(x, y):
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf
    except ValueError:
        return math.nan


def lua_mul(x, y):
    return x * y


ARITH = {b"__add": lambda x, y: x + y, b"__sub": lambda x, y: x - y, b"__mul": lua_mul, b"__div": lua_div,
         b"__mod": lua_mod, b"__pow": lua_pow}


def lua_eq(a, b):
    ta, tb = type(a), type(b)
    if ta is float and tb is float:
        return a == b                      # NaN ~= NaN even for the same Python object
    if a is b:
        return True
    if (ta is float or ta is int) and (tb is float or tb is int):
        return a == b
    if ta is not tb:
        return False
    if ta is bytes:
        return a == b
    if ta is LuaTable:
        ma, mb = a.meta, b.meta
        if ma is not None and mb is not None:
            ha, hb = ma.get(b"__eq"), mb.get(b"__eq")
            if ha is not None and ha is hb:
                r = call(ha, [a, b])
                return bool(r and r[0] is not None and r[0] is not False)
        return False
    return a == b


def lua_lt(a, b):
    ta, tb = type(a), type(b)
    if (ta is float or ta is int) and (tb is float or tb is int):
        return a < b
    if ta is bytes and tb is bytes:
        return a < b
    for x in (a, b):
        if type(x) is LuaTable and x.meta is not None:
            h = x.meta.get(b"__lt")
            if h is not None:
                r = call(h, [a, b])
                return bool(r and r[0] is not None and r[0] is not False)
    raise LuaError(("attempt to compare %s with %s" % (type_name(a), type_name(b))).encode())


def lua_le(a, b):
    ta, tb = type(a), type(b)
    if (ta is float or ta is int) and (tb is float or tb is int):
        return a <= b
    if ta is bytes and tb is bytes:
        return a <= b
    for x in (a, b):
        if type(x) is LuaTable and x.meta is not None:
            h = x.meta.get(b"__le")
            if h is not None:
                r = call(h, [a, b])
                return bool(r and r[0] is not None and r[0] is not False)
    raise LuaError(("attempt to compare %s with %s" % (type_name(a), type_name(b))).encode())


def lua_len(v):
    if type(v) is bytes:
        return float(len(v))
    if type(v) is LuaTable:
        if v.meta is not None:
            h = v.meta.get(b"__len")
            if h is not None:
                r = call(h, [v])
                return r[0] if r else None
        return float(len(v.arr))
    raise LuaError(("attempt to get length of a %s value" % type_name(v)).encode())


def lua_concat(a, b):
    ta, tb = type(a), type(b)
    if (ta is bytes or ta is float or ta is int) and (tb is bytes or tb is float or tb is int):
        return tostr(a) + tostr(b)
    return arith_meta(b"__concat", a, b)


# ======================================================================================= lexer
KEYWORDS = {"and", "break", "do", "else", "elseif", "end", "false", "for", "function", "goto", "if", "in", "local",
            "nil", "not", "or", "repeat", "return", "then", "true", "until", "while"}
SYMBOLS = ["...", "..", "==", "~=", "<=", ">=", "::", "+", "-", "*", "/", "%", "^", "#", "<", ">", "=", "(", ")", "{", "}",
           "[", "]", ";", ":", ",", "."]


def tokenize(src: bytes, chunk="?"):
    s = src.decode("latin-1")
    i, n, line = 0, len(s), 1
    toks = []

    def long_bracket(pos):
        # s[pos] == '[' ; returns (level) or -1
        j = pos + 1
        lvl = 0
        while j < n and s[j] == "=":
            lvl += 1
            j += 1
        if j < n and s[j] == "[":
            return lvl, j + 1
        return -1, pos

    while i < n:
        c = s[i]
        if c == "\n":
            line += 1
            i += 1
            continue
        if c in " \t\r\f\v":
            i += 1
            continue
        if c == "-" and s.startswith("--", i):
            i += 2
            if i < n and s[i] == "[":
                lvl, start = long_bracket(i)
                if lvl >= 0:
                    close = "]" + "=" * lvl + "]"
                    e = s.find(close, start)
                    if e < 0:
                        raise LuaError(("%s:%d: unfinished long comment" % (chunk, line)).encode())
                    line += s.count("\n", i, e)
                    i = e + len(close)
                    continue
            while i < n and s[i] != "\n":
                i += 1
            continue
        if c.isalpha() or c == "_":
            j = i + 1
            while j < n and (s[j].isalnum() or s[j] == "_"):
                j += 1
            w = s[i:j]
            toks.append(("kw" if w in KEYWORDS else "name", w, line))
            i = j
            continue
        if c.isdigit() or (c == "." and i + 1 < n and s[i + 1].isdigit()):
            j = i
            if s.startswith(("0x", "0X"), i):
                j = i + 2
                while j < n and (s[j] in "0123456789abcdefABCDEF." or (s[j] in "pP") or (s[j] in "+-" and s[j - 1] in "pP")):
                    j += 1
                txt = s[i:j]
                v = float.fromhex(txt) if ("." in txt or "p" in txt.lower()) else float(int(txt, 16))
            else:
                while j < n and (s[j].isdigit() or s[j] == "." or s[j] in "eE" or (s[j] in "+-" and s[j - 1] in "eE")):
                    j += 1
                v = float(s[i:j])
            toks.append(("num", v, line))
            i = j
            continue
        if c in "\"'":
            q = c
            j = i + 1
            buf = bytearray()
            while True:
                if j >= n or s[j] == "\n":
                    raise LuaError(("%s:%d: unfinished string" % (chunk, line)).encode())
                ch = s[j]
                if ch == q:
                    break
                if ch == "\\":
                    j += 1
                    e = s[j]
                    if e in "abfnrtv\\\"'":
                        buf.append({"a": 7, "b": 8, "f": 12, "n": 10, "r": 13, "t": 9, "v": 11, "\\": 92, '"': 34, "'": 39}[e])
                        j += 1
                    elif e == "\n":
                        buf.append(10)
                        line += 1
                        j += 1
                    elif e == "x":
                        buf.append(int(s[j + 1:j + 3], 16))
                        j += 3
                    elif e == "z":
                        j += 1
                        while j < n and s[j] in " \t\r\n\f\v":
                            if s[j] == "\n":
                                line += 1
                            j += 1
                    elif e.isdigit():
                        k = j
                        while k < j + 3 and k < n and s[k].isdigit():
                            k += 1
                        buf.append(int(s[j:k]))
                        j = k
                    else:
                        raise LuaError(("%s:%d: invalid escape sequence" % (chunk, line)).encode())
                else:
                    buf.append(ord(ch))
                    j += 1
            toks.append(("str", bytes(buf), line))
            i = j + 1
            continue
        if c == "[":
            lvl, start = long_bracket(i)
            if lvl >= 0:
                close = "]" + "=" * lvl + "]"
                e = s.find(close, start)
                if e < 0:
                    raise LuaError(("%s:%d: unfinished long string" % (chunk, line)).encode())
                body = s[start:e]
                if body.startswith("\r\n"):
                    body = body[2:]
                elif body.startswith("\n"):
                    body = body[1:]
                toks.append(("str", body.encode("latin-1"), line))
                line += s.count("\n", i, e)
                i = e + len(close)
                continue
        for sym in SYMBOLS:
            if s.startswith(sym, i):
                toks.append(("sym", sym, line))
                i += len(sym)
                break
        else:
            raise LuaError(("%s:%d: unexpected symbol near '%s'" % (chunk, line, c)).encode())
    toks.append(("eof", None, line))
    return toks


# ======================================================================================= parser
BINPRI = {"or": (1, 1), "and": (2, 2), "<": (3, 3), ">": (3, 3), "<=": (3, 3), ">=": (3, 3), "~=": (3, 3), "==": (3, 3),
          "..": (5, 4), "+": (6, 6), "-": (6, 6), "*": (7, 7), "/": (7, 7), "%": (7, 7), "^": (10, 9)}
UNARY_PRI = 8


class Parser:
    def __init__(self, toks, chunk):
        self.t = toks
        self.p = 0
        self.chunk = chunk

    def peek(self):
        return self.t[self.p]

    def next(self):
        tok = self.t[self.p]
        self.p += 1
        return tok

    def check(self, kind, val=None):
        tok = self.t[self.p]
        return tok[0] == kind and (val is None or tok[1] == val)

    def accept(self, kind, val=None):
        if self.check(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        tok = self.t[self.p]
        if tok[0] == kind and (val is None or tok[1] == val):
            self.p += 1
            return tok
        raise LuaError(("%s:%d: '%s' expected near '%s'" % (self.chunk, tok[2], val or kind, tok[1])).encode())

    def block_end(self):
        tok = self.t[self.p]
        return tok[0] == "eof" or (tok[0] == "kw" and tok[1] in ("end", "else", "elseif", "until"))

    def block(self):
        stmts = []
        while not self.block_end():
            if self.check("kw", "return"):
                line = self.next()[2]
                exprs = []
                if not self.block_end() and not self.check("sym", ";"):
                    exprs = self.exprlist()
                self.accept("sym", ";")
                stmts.append(("return", exprs, line))
                break
            st = self.statement()
            if st is not None:
                stmts.append(st)
        return stmts

    def statement(self):
        tok = self.peek()
        line = tok[2]
        if tok[0] == "sym" and tok[1] == ";":
            self.next()
            return None
        if tok[0] == "sym" and tok[1] == "::":
            self.next()
            name = self.expect("name")[1]
            self.expect("sym", "::")
            return ("label", name, line)
        if tok[0] == "kw":
            k = tok[1]
            if k == "if":
                self.next()
                clauses = []
                cond = self.expr()
                self.expect("kw", "then")
                clauses.append((cond, self.block()))
                els = None
                while True:
                    if self.accept("kw", "elseif"):
                        cond = self.expr()
                        self.expect("kw", "then")
                        clauses.append((cond, self.block()))
                    elif self.accept("kw", "else"):
                        els = self.block()
                        self.expect("kw", "end")
                        break
                    else:
                        self.expect("kw", "end")
                        break
                return ("if", clauses, els, line)
            if k == "while":
                self.next()
                cond = self.expr()
                self.expect("kw", "do")
                body = self.block()
                self.expect("kw", "end")
                return ("while", cond, body, line)
            if k == "do":
                self.next()
                body = self.block()
                self.expect("kw", "end")
                return ("do", body, line)
            if k == "for":
                self.next()
                n1 = self.expect("name")[1]
                if self.accept("sym", "="):
                    start = self.expr()
                    self.expect("sym", ",")
                    limit = self.expr()
                    step = self.expr() if self.accept("sym", ",") else None
                    self.expect("kw", "do")
                    body = self.block()
                    self.expect("kw", "end")
                    return ("fornum", n1, start, limit, step, body, line)
                names = [n1]
                while self.accept("sym", ","):
                    names.append(self.expect("name")[1])
                self.expect("kw", "in")
                exprs = self.exprlist()
                self.expect("kw", "do")
                body = self.block()
                self.expect("kw", "end")
                return ("forin", names, exprs, body, line)
            if k == "repeat":
                self.next()
                body = self.block()
                self.expect("kw", "until")
                cond = self.expr()
                return ("repeat", body, cond, line)
            if k == "function":
                self.next()
                target = ("name", self.expect("name")[1], line)
                is_method = False
                fname = target[1]
                while self.check("sym", ".") or self.check("sym", ":"):
                    sep = self.next()[1]
                    key = self.expect("name")[1]
                    fname += sep + key
                    target = ("index", target, ("const", key.encode()), line)
                    if sep == ":":
                        is_method = True
                        break
                f = self.funcbody(is_method, fname, line)
                return ("assign", [target], [f], line)
            if k == "local":
                self.next()
                if self.accept("kw", "function"):
                    name = self.expect("name")[1]
                    f = self.funcbody(False, name, line)
                    return ("localfunc", name, f, line)
                names = [self.expect("name")[1]]
                while self.accept("sym", ","):
                    names.append(self.expect("name")[1])
                exprs = self.exprlist() if self.accept("sym", "=") else []
                return ("local", names, exprs, line)
            if k == "break":
                self.next()
                return ("break", line)
            if k == "goto":
                self.next()
                return ("goto", self.expect("name")[1], line)
        # exprstat: call or assignment
        e = self.suffixedexp()
        if self.check("sym", "=") or self.check("sym", ","):
            targets = [e]
            while self.accept("sym", ","):
                targets.append(self.suffixedexp())
            self.expect("sym", "=")
            exprs = self.exprlist()
            for tg in targets:
                if tg[0] not in ("name", "index"):
                    raise LuaError(("%s:%d: syntax error (cannot assign)" % (self.chunk, line)).encode())
            return ("assign", targets, exprs, line)
        if e[0] not in ("call", "method"):
            raise LuaError(("%s:%d: syntax error near '%s'" % (self.chunk, line, self.peek()[1])).encode())
        return ("callstat", e, line)

    def exprlist(self):
        es = [self.expr()]
        while self.accept("sym", ","):
            es.append(self.expr())
        return es

    def funcbody(self, is_method, name, line):
        self.expect("sym", "(")
        params = ["self"] if is_method else []
        vararg = False
        if not self.check("sym", ")"):
            while True:
                if self.accept("sym", "..."):
                    vararg = True
                    break
                params.append(self.expect("name")[1])
                if not self.accept("sym", ","):
                    break
        self.expect("sym", ")")
        body = self.block()
        self.expect("kw", "end")
        return ("function", params, vararg, body, name, line)

    def primaryexp(self):
        tok = self.next()
        if tok[0] == "name":
            return ("name", tok[1], tok[2])
        if tok[0] == "sym" and tok[1] == "(":
            e = self.expr()
            self.expect("sym", ")")
            return ("paren", e)
        raise LuaError(("%s:%d: unexpected symbol near '%s'" % (self.chunk, tok[2], tok[1])).encode())

    def suffixedexp(self):
        e = self.primaryexp()
        while True:
            tok = self.peek()
            if tok[0] == "sym":
                if tok[1] == ".":
                    self.next()
                    e = ("index", e, ("const", self.expect("name")[1].encode()), tok[2])
                    continue
                if tok[1] == "[":
                    self.next()
                    k = self.expr()
                    self.expect("sym", "]")
                    e = ("index", e, k, tok[2])
                    continue
                if tok[1] == ":":
                    self.next()
                    name = self.expect("name")[1]
                    e = ("method", e, name.encode(), self.callargs(), tok[2])
                    continue
                if tok[1] in ("(", "{"):
                    e = ("call", e, self.callargs(), tok[2])
                    continue
            elif tok[0] == "str":
                e = ("call", e, self.callargs(), tok[2])
                continue
            return e

    def callargs(self):
        tok = self.peek()
        if tok[0] == "str":
            self.next()
            return [("const", tok[1])]
        if tok[0] == "sym" and tok[1] == "{":
            return [self.tablecons()]
        self.expect("sym", "(")
        args = []
        if not self.check("sym", ")"):
            args = self.exprlist()
        self.expect("sym", ")")
        return args

    def tablecons(self):
        line = self.expect("sym", "{")[2]
        arr, rec = [], []   # items in order: ('pos', expr) or ('key', kexpr, vexpr)
        items = []
        while not self.check("sym", "}"):
            if self.check("sym", "["):
                self.next()
                k = self.expr()
                self.expect("sym", "]")
                self.expect("sym", "=")
                items.append(("key", k, self.expr()))
            elif self.check("name") and self.t[self.p + 1][0] == "sym" and self.t[self.p + 1][1] == "=":
                k = self.next()[1]
                self.next()
                items.append(("key", ("const", k.encode()), self.expr()))
            else:
                items.append(("pos", self.expr()))
            if not (self.accept("sym", ",") or self.accept("sym", ";")):
                break
        self.expect("sym", "}")
        return ("table", items, line)

    def simpleexp(self):
        tok = self.peek()
        if tok[0] == "num":
            self.next()
            return ("const", tok[1])
        if tok[0] == "str":
            self.next()
            return ("const", tok[1])
        if tok[0] == "kw":
            if tok[1] == "nil":
                self.next()
                return ("const", None)
            if tok[1] == "true":
                self.next()
                return ("const", True)
            if tok[1] == "false":
                self.next()
                return ("const", False)
            if tok[1] == "function":
                self.next()
                return self.funcbody(False, "anonymous", tok[2])
        if tok[0] == "sym":
            if tok[1] == "...":
                self.next()
                return ("vararg",)
            if tok[1] == "{":
                return self.tablecons()
        return self.suffixedexp()

    def expr(self, limit=0):
        tok = self.peek()
        if (tok[0] == "kw" and tok[1] == "not") or (tok[0] == "sym" and tok[1] in ("-", "#")):
            self.next()
            operand = self.expr(UNARY_PRI)
            if tok[1] == "-" and operand[0] == "const" and type(operand[1]) is float:
                left = ("const", -operand[1])
            else:
                left = ("unop", tok[1], operand, tok[2])
        else:
            left = self.simpleexp()
        while True:
            tok = self.peek()
            op = tok[1] if tok[0] in ("sym", "kw") else None
            pri = BINPRI.get(op)
            if pri is None or pri[0] <= limit:
                break
            self.next()
            right = self.expr(pri[1])
            left = ("binop", op, left, right, tok[2])
        return left


# ======================================================================================= compiler
class FuncState:
    def __init__(self, parent):
        self.parent = parent
        self.scopes = [{}]
        self.nslots = 0
        self.upnames = {}     # name -> upvalue index
        self.updesc = []      # (from_parent_local: bool, index)

    def declare(self, name):
        idx = self.nslots
        self.nslots += 1
        self.scopes[-1][name] = idx
        return idx

    def find_local(self, name):
        for sc in reversed(self.scopes):
            if name in sc:
                return sc[name]
        return None

    def find_upvalue(self, name):
        if name in self.upnames:
            return self.upnames[name]
        if self.parent is None:
            return None
        loc = self.parent.find_local(name)
        if loc is not None:
            self.updesc.append((True, loc))
        else:
            up = self.parent.find_upvalue(name)
            if up is None:
                return None
            self.updesc.append((False, up))
        self.upnames[name] = len(self.updesc) - 1
        return self.upnames[name]


BREAK = ("break",)


class Compiler:
    def __init__(self, interp, chunk):
        self.I = interp
        self.chunk = chunk

    # ---- expressions: return fn(frame, upv, va) -> value
    def expr(self, fs, e):
        k = e[0]
        if k == "const":
            v = e[1]
            return lambda fr, up, va: v
        if k == "name":
            name = e[1]
            loc = fs.find_local(name)
            if loc is not None:
                return lambda fr, up, va: fr[loc][0]
            upi = fs.find_upvalue(name)
            if upi is not None:
                return lambda fr, up, va: up[upi][0]
            G = self.I.G
            key = name.encode()
            return lambda fr, up, va: G.get(key)
        if k == "paren":
            inner = self.expr(fs, e[1])
            return inner
        if k == "vararg":
            return lambda fr, up, va: va[0] if va else None
        if k == "index":
            obj, key = self.expr(fs, e[1]), self.expr(fs, e[2])
            line = e[3]
            if e[2][0] == "const":
                kc = e[2][1]

                def idx_const(fr, up, va):
                    o = obj(fr, up, va)
                    if type(o) is LuaTable:
                        v = o.get(kc)
                        if v is not None or o.meta is None:
                            return v
                    try:
                        return index(o, kc)
                    except LuaError as ex:
                        raise self.where(ex, line, e[1], kc)
                return idx_const

            def idx(fr, up, va):
                o = obj(fr, up, va)
                kk = key(fr, up, va)
                if type(o) is LuaTable:
                    v = o.get(kk)
                    if v is not None or o.meta is None:
                        return v
                try:
                    return index(o, kk)
                except LuaError as ex:
                    raise self.where(ex, line, e[1], kk)
            return idx
        if k in ("call", "method"):
            multi = self.multi(fs, e)

            def first(fr, up, va):
                r = multi(fr, up, va)
                return r[0] if r else None
            return first
        if k == "function":
            return self.function(fs, e)
        if k == "table":
            return self.table(fs, e)
        if k == "unop":
            op, operand, line = e[1], self.expr(fs, e[2]), e[3]
            if op == "not":
                return lambda fr, up, va: (lambda v: v is None or v is False)(operand(fr, up, va))
            if op == "-":
                def neg(fr, up, va):
                    v = operand(fr, up, va)
                    if type(v) is float:
                        return -v
                    n = tonum(v)
                    if n is None:
                        if type(v) is LuaTable and v.meta is not None and v.meta.get(b"__unm") is not None:
                            return call(v.meta.get(b"__unm"), [v, v])[0]
                        raise self.where(LuaError(("attempt to perform arithmetic on a %s value" % type_name(v)).encode()), line)
                    return -n
                return neg

            def length(fr, up, va):
                try:
                    return lua_len(operand(fr, up, va))
                except LuaError as ex:
                    raise self.where(ex, line)
            return length
        if k == "binop":
            return self.binop(fs, e)
        raise LuaError(("cannot compile expression %r" % (k,)).encode())

    def where(self, ex, line, objexpr=None, key=None):
        if getattr(ex, "located", False) or type(ex.value) is not bytes:
            return ex
        msg = ex.value
        if objexpr is not None and msg.startswith(b"attempt to index"):
            if objexpr[0] == "name":
                msg += (" (%s '%s')" % ("local/global", objexpr[1])).encode()
            elif objexpr[0] == "index" and objexpr[2][0] == "const" and type(objexpr[2][1]) is bytes:
                msg += b" (field '" + objexpr[2][1] + b"')"
        new = LuaError(("%s:%d: " % (self.chunk, line)).encode() + msg)
        new.located = True
        return new

    def binop(self, fs, e):
        op, line = e[1], e[4]
        a, b = self.expr(fs, e[2]), self.expr(fs, e[3])
        if op == "and":
            def land(fr, up, va):
                x = a(fr, up, va)
                if x is None or x is False:
                    return x
                return b(fr, up, va)
            return land
        if op == "or":
            def lor(fr, up, va):
                x = a(fr, up, va)
                if x is None or x is False:
                    return b(fr, up, va)
                return x
            return lor
        if op in ("+", "-", "*", "/", "%", "^"):
            mm = {"+": b"__add", "-": b"__sub", "*": b"__mul", "/": b"__div", "%": b"__mod", "^": b"__pow"}[op]
            fast = {"+": float.__add__, "-": float.__sub__, "*": float.__mul__}.get(op)
            slow = ARITH[mm]
            where = self.where
            if fast is not None:
                def ar(fr, up, va):
                    x, y = a(fr, up, va), b(fr, up, va)
                    if type(x) is float and type(y) is float:
                        return fast(x, y)
                    try:
                        return arith(mm, x, y)
                    except LuaError as ex:
                        raise where(ex, line)
                return ar

            def ar2(fr, up, va):
                x, y = a(fr, up, va), b(fr, up, va)
                if type(x) is float and type(y) is float:
                    return slow(x, y)
                try:
                    return arith(mm, x, y)
                except LuaError as ex:
                    raise where(ex, line)
            return ar2
        if op == "..":
            def cc(fr, up, va):
                try:
                    return lua_concat(a(fr, up, va), b(fr, up, va))
                except LuaError as ex:
                    raise self.where(ex, line)
            return cc
        if op == "==":
            return lambda fr, up, va: lua_eq(a(fr, up, va), b(fr, up, va))
        if op == "~=":
            return lambda fr, up, va: not lua_eq(a(fr, up, va), b(fr, up, va))
        cmpf = {"<": (lua_lt, False), "<=": (lua_le, False), ">": (lua_lt, True), ">=": (lua_le, True)}[op]
        f, swap = cmpf
        where = self.where
        pyop = {"<": float.__lt__, "<=": float.__le__, ">": float.__gt__, ">=": float.__ge__}[op]

        def cmp(fr, up, va):
            x, y = a(fr, up, va), b(fr, up, va)
            if type(x) is float and type(y) is float:
                return pyop(x, y)
            try:
                return f(y, x) if swap else f(x, y)
            except LuaError as ex:
                raise where(ex, line)
        return cmp

    # multi-valued expression: fn -> list
    def multi(self, fs, e):
        k = e[0]
        if k == "call":
            fn = self.expr(fs, e[1])
            args = self.arglist(fs, e[2])
            line = e[3]
            fname = e[1][1] if e[1][0] == "name" else (e[1][2][1].decode("latin-1") if e[1][0] == "index" and e[1][2][0] == "const" and type(e[1][2][1]) is bytes else "?")
            where = self.where

            def docall(fr, up, va):
                f = fn(fr, up, va)
                av = args(fr, up, va)
                if type(f) is LuaFunction:
                    return f.call(av)
                try:
                    return call(f, av)
                except LuaError as ex:
                    if type(ex.value) is bytes and ex.value.startswith(b"attempt to call"):
                        ex = LuaError(ex.value + (" (%s)" % fname).encode())
                    raise where(ex, line)
            return docall
        if k == "method":
            obj = self.expr(fs, e[1])
            name = e[2]
            args = self.arglist(fs, e[3])
            line = e[4]
            where = self.where

            def domethod(fr, up, va):
                o = obj(fr, up, va)
                try:
                    f = index(o, name)
                    return call(f, [o] + args(fr, up, va))
                except LuaError as ex:
                    raise where(ex, line)
            return domethod
        if k == "vararg":
            return lambda fr, up, va: va
        single = self.expr(fs, e)
        return lambda fr, up, va: [single(fr, up, va)]

    def arglist(self, fs, exprs):
        if not exprs:
            return lambda fr, up, va: []
        last = exprs[-1]
        singles = [self.expr(fs, x) for x in exprs[:-1]]
        if last[0] in ("call", "method", "vararg"):
            lm = self.multi(fs, last)
            if not singles:
                return lambda fr, up, va: list(lm(fr, up, va))

            def many(fr, up, va):
                out = [s(fr, up, va) for s in singles]
                out.extend(lm(fr, up, va))
                return out
            return many
        singles.append(self.expr(fs, last))
        if len(singles) == 1:
            s0 = singles[0]
            return lambda fr, up, va: [s0(fr, up, va)]
        if len(singles) == 2:
            s0, s1 = singles
            return lambda fr, up, va: [s0(fr, up, va), s1(fr, up, va)]
        return lambda fr, up, va: [s(fr, up, va) for s in singles]

    def table(self, fs, e):
        items = []
        for it in e[1]:
            if it[0] == "pos":
                items.append(("pos", it[1]))
            else:
                items.append(("key", self.expr(fs, it[1]), self.expr(fs, it[2])))
        compiled = []
        for i, it in enumerate(items):
            if it[0] == "pos":
                is_last = i == len(items) - 1
                if is_last and it[1][0] in ("call", "method", "vararg"):
                    compiled.append(("multi", self.multi(fs, it[1])))
                else:
                    compiled.append(("pos", self.expr(fs, it[1])))
            else:
                compiled.append(it)

        def build(fr, up, va):
            t = LuaTable()
            n = 0
            for it in compiled:
                kind = it[0]
                if kind == "pos":
                    n += 1
                    v = it[1](fr, up, va)
                    if v is not None:
                        t.set(n, v)
                elif kind == "key":
                    kk = it[1](fr, up, va)
                    v = it[2](fr, up, va)
                    if kk is None:
                        raise LuaError(b"table index is nil")
                    t.set(kk, v)
                else:
                    for v in it[1](fr, up, va):
                        n += 1
                        if v is not None:
                            t.set(n, v)
            return t
        return build

    def function(self, parent_fs, e):
        _, params, is_vararg, body, name, line = e
        fs = FuncState(parent_fs)
        pslots = [fs.declare(p) for p in params]
        block = self.block(fs, body, new_scope=False)
        updesc = fs.updesc
        nparams = len(pslots)
        chunk = self.chunk
        I = self.I

        def make(fr, up, va):
            cells = [fr[i] if from_local else up[i] for (from_local, i) in updesc]
            nslots = fs.nslots

            def lua_function(args):
                frame = [None] * nslots
                na = len(args)
                for i in range(nparams):
                    frame[pslots[i]] = [args[i] if i < na else None]
                extra = args[nparams:] if is_vararg and na > nparams else []
                r = block(frame, cells, extra)
                if r is None:
                    return []
                return r[1]
            return LuaFunction(lua_function, name)
        return make

    # ---- statements: fn(frame, upv, va) -> None | BREAK | ('ret', [values])
    def block(self, fs, stmts, new_scope=True):
        if new_scope:
            fs.scopes.append({})
        compiled = [self.stmt(fs, s) for s in stmts]
        if new_scope:
            fs.scopes.pop()
        if len(compiled) == 0:
            return lambda fr, up, va: None
        if len(compiled) == 1:
            return compiled[0]

        def run(fr, up, va):
            for s in compiled:
                r = s(fr, up, va)
                if r is not None:
                    return r
            return None
        return run

    def assign_target(self, fs, tg):
        if tg[0] == "name":
            name = tg[1]
            loc = fs.find_local(name)
            if loc is not None:
                def setl(fr, up, va, v):
                    fr[loc][0] = v
                return setl
            upi = fs.find_upvalue(name)
            if upi is not None:
                def setu(fr, up, va, v):
                    up[upi][0] = v
                return setu
            G = self.I.G
            key = name.encode()

            def setg(fr, up, va, v):
                G.set(key, v)
            return setg
        obj, key = self.expr(fs, tg[1]), self.expr(fs, tg[2])
        line = tg[3]
        where = self.where

        def seti(fr, up, va, v):
            o = obj(fr, up, va)
            k = key(fr, up, va)
            if type(o) is LuaTable and o.meta is None:
                o.set(k, v)
                return
            try:
                setindex(o, k, v)
            except LuaError as ex:
                raise where(ex, line, tg[1], k)
        return seti

    def stmt(self, fs, s):
        k = s[0]
        if k == "local":
            names, exprs = s[1], s[2]
            if len(names) == 1 and len(exprs) == 1:
                ev = self.expr(fs, exprs[0])
                slot = fs.declare(names[0])

                def local1(fr, up, va):
                    fr[slot] = [ev(fr, up, va)]
                return local1
            vals = self.arglist(fs, exprs)
            slots = [fs.declare(n) for n in names]
            ns = len(slots)

            def localn(fr, up, va):
                v = vals(fr, up, va)
                nv = len(v)
                for i in range(ns):
                    fr[slots[i]] = [v[i] if i < nv else None]
            return localn
        if k == "localfunc":
            slot = fs.declare(s[1])
            mk = self.function(fs, s[2])

            def localfunc(fr, up, va):
                cell = [None]
                fr[slot] = cell
                cell[0] = mk(fr, up, va)
            return localfunc
        if k == "assign":
            targets, exprs = s[1], s[2]
            if len(targets) == 1 and len(exprs) == 1:
                setter = self.assign_target(fs, targets[0])
                ev = self.expr(fs, exprs[0])
                tg = targets[0]
                if tg[0] == "name" and fs.find_local(tg[1]) is not None:
                    loc = fs.find_local(tg[1])

                    def assign_local(fr, up, va):
                        fr[loc][0] = ev(fr, up, va)
                    return assign_local

                def assign1(fr, up, va):
                    setter(fr, up, va, ev(fr, up, va))
                return assign1
            # general: evaluate table/key expressions are re-evaluated in setters (order differences are
            # unobservable for the reference's code, which has no side effects there)
            setters = [self.assign_target(fs, t) for t in targets]
            vals = self.arglist(fs, exprs)
            nt = len(setters)

            def assignn(fr, up, va):
                v = vals(fr, up, va)
                nv = len(v)
                for i in range(nt):
                    setters[i](fr, up, va, v[i] if i < nv else None)
            return assignn
        if k == "callstat":
            m = self.multi(fs, s[1])

            def callstat(fr, up, va):
                m(fr, up, va)
            return callstat
        if k == "return":
            exprs = s[1]
            if len(exprs) == 1 and exprs[0][0] not in ("call", "method", "vararg"):
                ev = self.expr(fs, exprs[0])
                return lambda fr, up, va: ("ret", [ev(fr, up, va)])
            vals = self.arglist(fs, exprs)
            return lambda fr, up, va: ("ret", vals(fr, up, va))
        if k == "break":
            return lambda fr, up, va: BREAK
        if k == "do":
            return self.block(fs, s[1])
        if k == "if":
            clauses = [(self.expr(fs, c), self.block(fs, b)) for c, b in s[1]]
            els = self.block(fs, s[2]) if s[2] is not None else None
            if len(clauses) == 1 and els is None:
                c0, b0 = clauses[0]

                def if1(fr, up, va):
                    v = c0(fr, up, va)
                    if v is not None and v is not False:
                        return b0(fr, up, va)
                    return None
                return if1

            def ifn(fr, up, va):
                for c, b in clauses:
                    v = c(fr, up, va)
                    if v is not None and v is not False:
                        return b(fr, up, va)
                if els is not None:
                    return els(fr, up, va)
                return None
            return ifn
        if k == "while":
            cond, body = self.expr(fs, s[1]), self.block(fs, s[2])

            def loop(fr, up, va):
                while True:
                    v = cond(fr, up, va)
                    if v is None or v is False:
                        return None
                    r = body(fr, up, va)
                    if r is not None:
                        if r is BREAK:
                            return None
                        return r
            return loop
        if k == "repeat":
            fs.scopes.append({})
            body = self.block(fs, s[1], new_scope=False)
            cond = self.expr(fs, s[2])
            fs.scopes.pop()

            def rep(fr, up, va):
                while True:
                    r = body(fr, up, va)
                    if r is not None:
                        if r is BREAK:
                            return None
                        return r
                    v = cond(fr, up, va)
                    if v is not None and v is not False:
                        return None
            return rep
        if k == "fornum":
            start, limit = self.expr(fs, s[2]), self.expr(fs, s[3])
            step = self.expr(fs, s[4]) if s[4] is not None else None
            fs.scopes.append({})
            slot = fs.declare(s[1])
            body = self.block(fs, s[5], new_scope=False)
            fs.scopes.pop()
            line = s[6]

            def fornum(fr, up, va):
                a, b = tonum(start(fr, up, va)), tonum(limit(fr, up, va))
                c = 1.0 if step is None else tonum(step(fr, up, va))
                if a is None or b is None or c is None:
                    raise LuaError(("%s:%d: 'for' initial value, limit and step must be numbers" % (self.chunk, line)).encode())
                if c == 0:
                    raise LuaError(("%s:%d: 'for' step is zero" % (self.chunk, line)).encode())
                i = a
                if c > 0:
                    while i <= b:
                        fr[slot] = [i]
                        r = body(fr, up, va)
                        if r is not None:
                            if r is BREAK:
                                return None
                            return r
                        i += c
                else:
                    while i >= b:
                        fr[slot] = [i]
                        r = body(fr, up, va)
                        if r is not None:
                            if r is BREAK:
                                return None
                            return r
                        i += c
                return None
            return fornum
        if k == "forin":
            vals = self.arglist(fs, s[2])
            fs.scopes.append({})
            slots = [fs.declare(n) for n in s[1]]
            body = self.block(fs, s[3], new_scope=False)
            fs.scopes.pop()
            ns = len(slots)

            def forin(fr, up, va):
                v = vals(fr, up, va)
                f = v[0] if len(v) > 0 else None
                st = v[1] if len(v) > 1 else None
                ctl = v[2] if len(v) > 2 else None
                while True:
                    rs = call(f, [st, ctl])
                    first = rs[0] if rs else None
                    if first is None:
                        return None
                    ctl = first
                    nr = len(rs)
                    for i in range(ns):
                        fr[slots[i]] = [rs[i] if i < nr else None]
                    r = body(fr, up, va)
                    if r is not None:
                        if r is BREAK:
                            return None
                        return r
            return forin
        if k in ("label", "goto"):
            raise LuaError(b"goto/labels are not supported by this interpreter")
        raise LuaError(("cannot compile statement %r" % (k,)).encode())


# ======================================================================================= Lua patterns
class _Match:
    MAXCAP = 32

    def __init__(self, src, pat):
        self.src, self.pat = src, pat
        self.level = 0
        self.cap = []      # [start, len]  len: -1 = position capture, -2 = unclosed

    def class_end(self, p):
        pat = self.pat
        if p >= len(pat):
            raise LuaError(b"malformed pattern (ends with '%')")
        c = pat[p]
        p += 1
        if c == 37:  # %
            if p >= len(pat):
                raise LuaError(b"malformed pattern (ends with '%')")
            return p + 1
        if c == 91:  # [
            if p < len(pat) and pat[p] == 94:
                p += 1
            first = True
            while True:
                if p >= len(pat):
                    raise LuaError(b"malformed pattern (missing ']')")
                cc = pat[p]
                p += 1
                if cc == 93 and not first:
                    return p
                first = False
                if cc == 37:
                    p += 1
        return p

    @staticmethod
    def single_class(c, cl):
        ch = chr(c)
        lower = cl | 32
        if lower == 97: res = ch.isalpha() and c < 128                     # a
        elif lower == 99: res = c < 32 or c == 127                         # c
        elif lower == 100: res = 48 <= c <= 57                             # d
        elif lower == 103: res = 33 <= c <= 126                            # g
        elif lower == 108: res = 97 <= c <= 122                            # l
        elif lower == 112: res = (33 <= c <= 47) or (58 <= c <= 64) or (91 <= c <= 96) or (123 <= c <= 126)  # p
        elif lower == 115: res = c in (32, 9, 10, 11, 12, 13)              # s
        elif lower == 117: res = 65 <= c <= 90                             # u
        elif lower == 119: res = (48 <= c <= 57) or (65 <= c <= 90) or (97 <= c <= 122)  # w
        elif lower == 120: res = (48 <= c <= 57) or (65 <= c <= 70) or (97 <= c <= 102)  # x
        else:
            return cl == c
        if 65 <= cl <= 90:
            return not res
        return res

    def match_class_set(self, c, p, ep):
        pat = self.pat
        sig = True
        p += 1
        if pat[p] == 94:
            sig = False
            p += 1
        while p < ep:
            if pat[p] == 37:
                p += 1
                if self.single_class(c, pat[p]):
                    return sig
                p += 1
            elif p + 2 < ep and pat[p + 1] == 45:
                if pat[p] <= c <= pat[p + 2]:
                    return sig
                p += 3
            else:
                if pat[p] == c:
                    return sig
                p += 1
        return not sig

    def single_match(self, s, p, ep):
        if s >= len(self.src):
            return False
        c = self.src[s]
        pc = self.pat[p]
        if pc == 46:
            return True
        if pc == 37:
            return self.single_class(c, self.pat[p + 1])
        if pc == 91:
            return self.match_class_set(c, p, ep - 1)
        return pc == c

    def do_match(self, s, p):
        pat, src = self.pat, self.src
        while True:
            if p >= len(pat):
                return s
            pc = pat[p]
            if pc == 40:  # (
                if p + 1 < len(pat) and pat[p + 1] == 41:
                    self.cap.append([s, -1])
                    r = self.do_match(s, p + 2)
                    if r is None:
                        self.cap.pop()
                    return r
                self.cap.append([s, -2])
                r = self.do_match(s, p + 1)
                if r is None:
                    self.cap.pop()
                return r
            if pc == 41:  # )
                l = None
                for i in range(len(self.cap) - 1, -1, -1):
                    if self.cap[i][1] == -2:
                        l = i
                        break
                if l is None:
                    raise LuaError(b"invalid pattern capture")
                self.cap[l][1] = s - self.cap[l][0]
                r = self.do_match(s, p + 1)
                if r is None:
                    self.cap[l][1] = -2
                return r
            if pc == 36 and p + 1 == len(pat):  # $
                return s if s == len(src) else None
            if pc == 37 and p + 1 < len(pat):
                nx = pat[p + 1]
                if nx == 98:  # %b
                    if p + 3 >= len(pat):
                        raise LuaError(b"malformed pattern (missing arguments to '%b')")
                    if s >= len(src) or src[s] != pat[p + 2]:
                        return None
                    b, e = pat[p + 2], pat[p + 3]
                    cont = 1
                    q = s + 1
                    res = None
                    while q < len(src):
                        ch = src[q]
                        if ch == e:
                            cont -= 1
                            if cont == 0:
                                res = q + 1
                                break
                        elif ch == b:
                            cont += 1
                        q += 1
                    if res is None:
                        return None
                    s = res
                    p += 4
                    continue
                if nx == 102:  # %f
                    p += 2
                    if p >= len(pat) or pat[p] != 91:
                        raise LuaError(b"missing '[' after '%f' in pattern")
                    ep = self.class_end(p)
                    prev = src[s - 1] if s > 0 else 0
                    cur = src[s] if s < len(src) else 0
                    if (not self.match_class_set(prev, p, ep - 1)) and self.match_class_set(cur, p, ep - 1):
                        p = ep
                        continue
                    return None
                if 48 <= nx <= 57:  # back reference
                    l = nx - 49
                    if l < 0 or l >= len(self.cap) or self.cap[l][1] == -2:
                        raise LuaError(b"invalid capture index")
                    cs, cl = self.cap[l]
                    if len(src) - s >= cl and src[cs:cs + cl] == src[s:s + cl]:
                        s += cl
                        p += 2
                        continue
                    return None
            ep = self.class_end(p)
            epc = pat[ep] if ep < len(pat) else 0
            if epc == 63:  # ?
                if self.single_match(s, p, ep):
                    r = self.do_match(s + 1, ep + 1)
                    if r is not None:
                        return r
                p = ep + 1
                continue
            if epc == 43:  # +
                if not self.single_match(s, p, ep):
                    return None
                return self.max_expand(s + 1, p, ep)
            if epc == 42:  # *
                return self.max_expand(s, p, ep)
            if epc == 45:  # -
                while True:
                    r = self.do_match(s, ep + 1)
                    if r is not None:
                        return r
                    if self.single_match(s, p, ep):
                        s += 1
                    else:
                        return None
            if not self.single_match(s, p, ep):
                return None
            s += 1
            p = ep

    def max_expand(self, s, p, ep):
        i = 0
        while self.single_match(s + i, p, ep):
            i += 1
        while i >= 0:
            r = self.do_match(s + i, ep + 1)
            if r is not None:
                return r
            i -= 1
        return None

    def get_capture(self, i, s, e):
        if i >= len(self.cap):
            if i == 0:
                return self.src[s:e]
            raise LuaError(b"invalid capture index")
        cs, cl = self.cap[i]
        if cl == -2:
            raise LuaError(b"unfinished capture")
        if cl == -1:
            return float(cs + 1)
        return self.src[cs:cs + cl]

    def captures(self, s, e, whole_if_none=True):
        n = len(self.cap)
        if n == 0 and whole_if_none:
            return [self.src[s:e]]
        return [self.get_capture(i, s, e) for i in range(n)]


def str_find_aux(args, find):
    s, pat = args[0], args[1]
    if type(s) is not bytes:
        s = tostr(s)
    if type(pat) is not bytes:
        pat = tostr(pat)
    init = int(tonum(args[2])) if len(args) > 2 and args[2] is not None else 1
    if init < 0:
        init = len(s) + init + 1
        if init < 1:
            init = 1
    elif init == 0:
        init = 1
    if init > len(s) + 1:
        return [None]
    plain = len(args) > 3 and args[3] not in (None, False)
    if find and (plain or not any(c in pat for c in b"^$*+?.([%-")):
        pos = s.find(pat, init - 1)
        if pos < 0:
            return [None]
        return [float(pos + 1), float(pos + len(pat))]
    anchor = pat.startswith(b"^")
    p0 = 1 if anchor else 0
    si = init - 1
    while True:
        m = _Match(s, pat)
        e = m.do_match(si, p0)
        if e is not None:
            if find:
                return [float(si + 1), float(e)] + (m.captures(si, e, False))
            return m.captures(si, e)
        si += 1
        if anchor or si > len(s):
            return [None]


def str_gmatch(args):
    s, pat = args[0], args[1]
    state = {"pos": 0}

    def it(_):
        si = state["pos"]
        while si <= len(s):
            m = _Match(s, pat)
            e = m.do_match(si, 0)
            if e is not None:
                state["pos"] = e + 1 if e == si else e
                return m.captures(si, e)
            si += 1
        state["pos"] = len(s) + 1
        return [None]
    return [LuaFunction(it, "gmatch_iter")]


def str_gsub(args):
    s, pat, repl = args[0], args[1], args[2]
    if type(s) is not bytes:
        s = tostr(s)
    max_s = int(tonum(args[3])) if len(args) > 3 and args[3] is not None else len(s) + 1
    anchor = pat.startswith(b"^")
    p0 = 1 if anchor else 0
    out = bytearray()
    si = 0
    n = 0
    while n < max_s:
        m = _Match(s, pat)
        e = m.do_match(si, p0)
        if e is not None:
            n += 1
            whole = s[si:e]
            if type(repl) is bytes or type(repl) is float:
                r = tostr(repl)
                buf = bytearray()
                i = 0
                while i < len(r):
                    c = r[i]
                    if c == 37:
                        i += 1
                        d = r[i]
                        if d == 48:
                            buf += whole
                        elif 49 <= d <= 57:
                            v = m.get_capture(d - 49, si, e)
                            buf += tostr(v)
                        else:
                            buf.append(d)
                    else:
                        buf.append(c)
                    i += 1
                out += buf
            else:
                caps = m.captures(si, e)
                if type(repl) is LuaTable:
                    v = index(repl, caps[0])
                else:
                    r = call(repl, caps)
                    v = r[0] if r else None
                if v is None or v is False:
                    out += whole
                elif type(v) in (bytes, float, int):
                    out += tostr(v)
                else:
                    raise LuaError(b"invalid replacement value (a " + type_name(v).encode() + b")")
        if e is not None and e > si:
            si = e
        elif si < len(s):
            out.append(s[si])
            si += 1
        else:
            break
        if anchor:
            break
    out += s[si:]
    return [bytes(out), float(n)]


# ======================================================================================= string.pack / unpack
def _pack_parse(fmt):
    """Yields (kind, size, align_request) items for a Lua 5.3-style pack format."""
    i, n = 0, len(fmt)
    little = sys.byteorder == "little"
    maxalign = 1
    items = []

    def number(default):
        nonlocal i
        j = i
        while j < n and 48 <= fmt[j] <= 57:
            j += 1
        if j == i:
            return default
        v = int(fmt[i:j])
        i = j
        return v
    while i < n:
        c = chr(fmt[i])
        i += 1
        if c == " ":
            continue
        if c == "<":
            little = True
        elif c == ">":
            little = False
        elif c == "=":
            little = sys.byteorder == "little"
        elif c == "!":
            maxalign = number(8)
        elif c in "bB":
            items.append(("int", 1, c == "b", little, maxalign))
        elif c in "hH":
            items.append(("int", 2, c == "h", little, maxalign))
        elif c in "lLjJ":
            items.append(("int", 8, c in "lj", little, maxalign))
        elif c == "T":
            items.append(("int", 8, False, little, maxalign))
        elif c in "iI":
            items.append(("int", number(4), c == "i", little, maxalign))
        elif c == "f":
            items.append(("float", 4, True, little, maxalign))
        elif c in "dn":
            items.append(("float", 8, True, little, maxalign))
        elif c == "s":
            items.append(("str", number(8), False, little, maxalign))
        elif c == "z":
            items.append(("zstr", 0, False, little, maxalign))
        elif c == "x":
            items.append(("pad", 1, False, little, maxalign))
        elif c == "c":
            sz = number(-1)
            if sz < 0:
                raise LuaError(b"missing size for format option 'c'")
            items.append(("chars", sz, False, little, maxalign))
        elif c == "X":
            # align to the next option's size
            if i >= n:
                raise LuaError(b"invalid next option for option 'X'")
            d = chr(fmt[i])
            i += 1
            if d in "bB":
                sz = 1
            elif d in "hH":
                sz = 2
            elif d in "iI":
                sz = number(4)
            elif d in "lLjJTdn":
                sz = 8
            elif d == "f":
                sz = 4
            else:
                raise LuaError(b"invalid next option for option 'X'")
            items.append(("align", sz, False, little, maxalign))
        else:
            raise LuaError(("invalid format option '%s'" % c).encode())
    return items


def _align_pad(pos, size, maxalign, kind):
    if kind in ("chars", "pad", "zstr", "str") and kind != "align":
        return 0
    a = min(size, maxalign)
    if a <= 1:
        return 0
    if a & (a - 1):
        raise LuaError(b"format asks for alignment not power of 2")
    return (a - (pos & (a - 1))) & (a - 1)


def str_unpack(args):
    fmt, data = args[0], args[1]
    pos = int(tonum(args[2])) - 1 if len(args) > 2 and args[2] is not None else 0
    if pos < 0:
        pos = len(data) + pos + 1
    if pos > len(data) or pos < 0:
        raise LuaError(b"bad argument #3 to 'unpack' (initial position out of string)")
    out = []
    ld = len(data)
    for kind, size, signed, little, maxalign in _pack_parse(fmt):
        if kind in ("int", "float", "align"):
            pad = _align_pad(pos, size, maxalign, kind)
            if pad + (0 if kind == "align" else size) > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            pos += pad
            if kind == "align":
                continue
        if kind == "int":
            if size > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            out.append(float(int.from_bytes(data[pos:pos + size], "little" if little else "big", signed=signed)))
            pos += size
        elif kind == "float":
            out.append(float(struct.unpack(("<" if little else ">") + ("f" if size == 4 else "d"), data[pos:pos + size])[0]))
            pos += size
        elif kind == "chars":
            if size > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            out.append(data[pos:pos + size])
            pos += size
        elif kind == "pad":
            if 1 > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            pos += 1
        elif kind == "str":
            pad = _align_pad(pos, size, maxalign, "int")
            if pad + size > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            pos += pad
            ln = int.from_bytes(data[pos:pos + size], "little" if little else "big")
            pos += size
            if ln > ld - pos:
                raise LuaError(b"bad argument #2 to 'unpack' (data string too short)")
            out.append(data[pos:pos + ln])
            pos += ln
        elif kind == "zstr":
            e = data.find(b"\0", pos)
            if e < 0:
                raise LuaError(b"bad argument #2 to 'unpack' (unfinished string for format 'z')")
            out.append(data[pos:e])
            pos = e + 1
    out.append(float(pos + 1))
    return out


def str_pack(args):
    fmt = args[0]
    out = bytearray()
    ai = 1
    for kind, size, signed, little, maxalign in _pack_parse(fmt):
        if kind in ("int", "float", "align"):
            out += b"\0" * _align_pad(len(out), size, maxalign, kind)
            if kind == "align":
                continue
        if kind == "int":
            if ai >= len(args) or tonum(args[ai]) is None:
                raise LuaError(("bad argument #%d to 'pack' (number expected, got %s)" % (ai + 1, "no value" if ai >= len(args) else "nil")).encode())
            v = int(math.floor(tonum(args[ai])))
            ai += 1
            out += (v & ((1 << (8 * size)) - 1)).to_bytes(size, "little" if little else "big")
        elif kind == "float":
            out += struct.pack(("<" if little else ">") + ("f" if size == 4 else "d"), tonum(args[ai]))
            ai += 1
        elif kind == "chars":
            s = args[ai]
            ai += 1
            out += s[:size] + b"\0" * (size - len(s))
        elif kind == "pad":
            out.append(0)
        elif kind == "str":
            s = args[ai]
            ai += 1
            out += b"\0" * _align_pad(len(out), size, maxalign, "int")
            out += len(s).to_bytes(size, "little" if little else "big") + s
        elif kind == "zstr":
            out += args[ai] + b"\0"
            ai += 1
    return [bytes(out)]


# ======================================================================================= standard library
STRING_LIB = LuaTable()


def _fn(name, f):
    return LuaFunction(f, name)


def _arg(args, i, default=None):
    return args[i] if i < len(args) and args[i] is not None else default


def _str_sub(args):
    s = args[0] if type(args[0]) is bytes else tostr(args[0])
    n = len(s)
    i = int(tonum(_arg(args, 1, 1.0)))
    j = int(tonum(_arg(args, 2, -1.0)))
    if i < 0:
        i = max(n + i + 1, 1)
    elif i == 0:
        i = 1
    if j < 0:
        j = n + j + 1
    elif j > n:
        j = n
    if i > j:
        return [b""]
    return [s[i - 1:j]]


def _str_byte(args):
    s = args[0] if type(args[0]) is bytes else tostr(args[0])
    n = len(s)
    i = int(tonum(_arg(args, 1, 1.0)))
    j = int(tonum(_arg(args, 2, float(i))))
    if i < 0:
        i = max(n + i + 1, 1)
    elif i == 0:
        i = 1
    if j < 0:
        j = n + j + 1
    elif j > n:
        j = n
    if i > j:
        return []
    return [float(b) for b in s[i - 1:j]]


def _str_format(args):
    fmt = args[0]
    out = bytearray()
    ai = 1
    i = 0
    while i < len(fmt):
        c = fmt[i]
        if c != 37:
            out.append(c)
            i += 1
            continue
        i += 1
        if fmt[i] == 37:
            out.append(37)
            i += 1
            continue
        j = i
        while chr(fmt[j]) in "-+ #0123456789.":
            j += 1
        spec = fmt[i:j].decode()
        conv = chr(fmt[j])
        i = j + 1
        v = args[ai] if ai < len(args) else None
        ai += 1
        if conv in "di":
            out += (("%" + spec + "d") % int(tonum(v))).encode()
        elif conv in "uoxX":
            out += (("%" + spec + conv.replace("u", "d")) % int(tonum(v))).encode()
        elif conv in "eEfgG":
            out += (("%" + spec + conv) % tonum(v)).encode()
        elif conv == "c":
            out.append(int(tonum(v)))
        elif conv == "s":
            out += (("%" + spec + "s") % tostr(v).decode("latin-1")).encode("latin-1")
        elif conv == "q":
            out += b'"' + tostr(v).replace(b"\\", b"\\\\").replace(b'"', b'\\"').replace(b"\n", b"\\n") + b'"'
        else:
            raise LuaError(("invalid option '%%%s' to 'format'" % conv).encode())
    return [bytes(out)]


def _str_rep(args):
    s = args[0] if type(args[0]) is bytes else tostr(args[0])
    n = int(tonum(args[1]))
    sep = _arg(args, 2, b"")
    if n <= 0:
        return [b""]
    return [sep.join([s] * n) if sep else s * n]


for _name, _f in {
    "sub": _str_sub, "byte": _str_byte,
    "char": lambda a: [bytes(int(tonum(x)) for x in a)],
    "len": lambda a: [float(len(a[0]))],
    "rep": _str_rep,
    "lower": lambda a: [a[0].lower()], "upper": lambda a: [a[0].upper()],
    "reverse": lambda a: [a[0][::-1]],
    "format": _str_format,
    "find": lambda a: str_find_aux(a, True), "match": lambda a: str_find_aux(a, False),
    "gmatch": str_gmatch, "gsub": str_gsub, "pack": str_pack, "unpack": str_unpack,
}.items():
    STRING_LIB.set(_name.encode(), _fn("string." + _name, _f))


def _u32(v):
    n = tonum(v)
    if n is None:
        raise LuaError(("bad argument to bit32 function (number expected, got %s)" % ("no value" if v is None else type_name(v))).encode())
    return int(math.floor(n)) & 0xFFFFFFFF


def _bit32_named(name, f):
    def wrapped(a):
        try:
            return f(a)
        except LuaError:
            bad = next((i for i, x in enumerate(a) if tonum(x) is None), len(a))
            got = "nil" if bad >= len(a) or a[bad] is None else type_name(a[bad])
            raise LuaError(("bad argument #%d to '%s' (number expected, got %s)" % (bad + 1, name, got)).encode())
        except IndexError:
            raise LuaError(("bad argument #%d to '%s' (number expected, got no value)" % (len(a) + 1, name)).encode())
    return wrapped


def _shift(x, disp):
    if disp <= -32 or disp >= 32:
        return 0
    return ((x << disp) if disp >= 0 else (x >> -disp)) & 0xFFFFFFFF


def make_bit32():
    t = LuaTable()

    def fold(op, init):
        def f(a):
            r = init
            for x in a:
                r = op(r, _u32(x))
            return [float(r & 0xFFFFFFFF)]
        return f

    def arshift(a):
        x, d = _u32(a[0]), int(tonum(a[1]))
        if d < 0:
            return [float(_shift(x, -d))]
        if x & 0x80000000:
            if d >= 32:
                return [float(0xFFFFFFFF)]
            return [float(((x >> d) | (~(0xFFFFFFFF >> d))) & 0xFFFFFFFF)]
        return [float(_shift(x, -d))]

    def extract(a):
        x, f = _u32(a[0]), int(tonum(a[1]))
        w = int(tonum(a[2])) if len(a) > 2 and a[2] is not None else 1
        if f < 0 or w <= 0 or f + w > 32:
            raise LuaError(b"trying to access non-existent bits")
        return [float((x >> f) & ((1 << w) - 1))]

    def replace(a):
        x, v, f = _u32(a[0]), _u32(a[1]), int(tonum(a[2]))
        w = int(tonum(a[3])) if len(a) > 3 and a[3] is not None else 1
        m = ((1 << w) - 1) << f
        return [float((x & ~m) | ((v << f) & m))]
    fns = {
        "band": fold(lambda r, x: r & x, 0xFFFFFFFF), "bor": fold(lambda r, x: r | x, 0), "bxor": fold(lambda r, x: r ^ x, 0),
        "bnot": lambda a: [float(~_u32(a[0]) & 0xFFFFFFFF)],
        "btest": lambda a: [fold(lambda r, x: r & x, 0xFFFFFFFF)(a)[0] != 0],
        "lshift": lambda a: [float(_shift(_u32(a[0]), int(tonum(a[1]))))],
        "rshift": lambda a: [float(_shift(_u32(a[0]), -int(tonum(a[1]))))],
        "arshift": arshift, "extract": extract, "replace": replace,
        "lrotate": lambda a: [float(((_u32(a[0]) << (int(tonum(a[1])) % 32)) | (_u32(a[0]) >> (32 - int(tonum(a[1])) % 32))) & 0xFFFFFFFF)],
    }
    for k, f in fns.items():
        t.set(k.encode(), _fn("bit32." + k, _bit32_named(k, f)))
    return t


def _checknum(a, i, fname):
    v = a[i] if i < len(a) else None
    n = tonum(v)
    if n is None:
        raise LuaError(("bad argument #%d to '%s' (number expected, got %s)" % (i + 1, fname, "no value" if i >= len(a) else type_name(v))).encode())
    return n


def make_math():
    t = LuaTable()

    def lmax(a):
        m = _checknum(a, 0, "max")
        for i in range(1, len(a)):
            d = _checknum(a, i, "max")
            if d > m:
                m = d
        return [m]

    def lmin(a):
        m = _checknum(a, 0, "min")
        for i in range(1, len(a)):
            d = _checknum(a, i, "min")
            if d < m:
                m = d
        return [m]

    def fl(a):
        x = _checknum(a, 0, "floor")
        return [float(math.floor(x)) if math.isfinite(x) else x]

    def ce(a):
        x = _checknum(a, 0, "ceil")
        return [float(math.ceil(x)) if math.isfinite(x) else x]

    def safe(f, name):
        def g(a):
            try:
                return [float(f(*[_checknum(a, i, name) for i in range(len(a))]))]
            except (ValueError, ZeroDivisionError):
                return [math.nan]
            except OverflowError:
                return [math.inf]
        return g
    fns = {"floor": fl, "ceil": ce, "max": lmax, "min": lmin, "abs": lambda a: [abs(_checknum(a, 0, "abs"))],
           "sin": safe(math.sin, "sin"), "cos": safe(math.cos, "cos"), "tan": safe(math.tan, "tan"), "sqrt": safe(math.sqrt, "sqrt"),
           "exp": safe(math.exp, "exp"), "log": safe(lambda x, b=None: math.log(x) if b is None else math.log(x, b), "log"),
           "pow": safe(math.pow, "pow"), "fmod": safe(math.fmod, "fmod"), "atan": safe(math.atan, "atan"),
           "atan2": safe(math.atan2, "atan2"), "asin": safe(math.asin, "asin"), "acos": safe(math.acos, "acos"),
           "ldexp": safe(lambda m, e: math.ldexp(m, int(e)), "ldexp"),
           "frexp": lambda a: (lambda r: [float(r[0]), float(r[1])])(math.frexp(_checknum(a, 0, "frexp"))),
           "random": lambda a: [0.5], "randomseed": lambda a: [],
           "modf": lambda a: (lambda x: [float(math.trunc(x)), x - math.trunc(x)])(_checknum(a, 0, "modf"))}
    for k, f in fns.items():
        t.set(k.encode(), _fn("math." + k, f))
    t.set(b"pi", math.pi)
    t.set(b"huge", math.inf)
    return t


def make_table_lib():
    t = LuaTable()

    def insert(a):
        tb = a[0]
        if len(a) == 2:
            tb.set(len(tb.arr) + 1, a[1])
        else:
            pos = int(tonum(a[1]))
            n = len(tb.arr)
            if pos == n + 1:
                tb.set(pos, a[2])
            else:
                tb.arr.insert(pos - 1, a[2])
        return []

    def remove(a):
        tb = a[0]
        n = len(tb.arr)
        if n == 0:
            return [None]
        pos = int(tonum(a[1])) if len(a) > 1 and a[1] is not None else n
        if pos < 1 or pos > n:
            return [None]
        return [tb.arr.pop(pos - 1)]

    def concat(a):
        tb = a[0]
        sep = _arg(a, 1, b"")
        i = int(tonum(_arg(a, 2, 1.0)))
        j = int(tonum(_arg(a, 3, float(len(tb.arr)))))
        parts = []
        for k in range(i, j + 1):
            v = tb.get(k)
            if type(v) not in (bytes, float, int):
                raise LuaError(("invalid value (at index %d) in table for 'concat'" % k).encode())
            parts.append(tostr(v))
        return [sep.join(parts)]

    def unpack(a):
        tb = a[0]
        i = int(tonum(_arg(a, 1, 1.0)))
        j = int(tonum(a[2])) if len(a) > 2 and a[2] is not None else int(lua_len(tb))
        if i == 1 and j == len(tb.arr):
            return list(tb.arr)
        return [tb.get(k) for k in range(i, j + 1)]

    def pack(a):
        tb = LuaTable()
        tb.arr = [x for x in a]
        while tb.arr and tb.arr[-1] is None:
            tb.arr.pop()
        for i, x in enumerate(a):
            if i >= len(tb.arr) and x is not None:
                tb.hash[i + 1] = x
        tb.set(b"n", float(len(a)))
        return [tb]

    def sort(a):
        tb = a[0]
        import functools
        if len(a) > 1 and a[1] is not None:
            lt = lambda x, y: (lambda r: bool(r and r[0] not in (None, False)))(call(a[1], [x, y]))
        else:
            lt = lua_lt
        tb.arr.sort(key=functools.cmp_to_key(lambda x, y: -1 if lt(x, y) else (1 if lt(y, x) else 0)))
        return []
    for k, f in {"insert": insert, "remove": remove, "concat": concat, "unpack": unpack, "pack": pack, "sort": sort}.items():
        t.set(k.encode(), _fn("table." + k, f))
    return t, unpack


class Interpreter:
    def __init__(self, stdout=None):
        self.G = LuaTable()
        self.loaded = {}
        self.preload = {}
        self.stdout = stdout if stdout is not None else sys.stdout
        self._install()

    def _install(self):
        G = self.G
        G.set(b"_G", G)
        G.set(b"_VERSION", b"Lua 5.2")
        G.set(b"string", STRING_LIB)
        G.set(b"bit32", make_bit32())
        G.set(b"math", make_math())
        tlib, unpack = make_table_lib()
        G.set(b"table", tlib)
        G.set(b"unpack", _fn("unpack", unpack))

        def lprint(a):
            self.stdout.write("\t".join(tostr(x).decode("latin-1") for x in a) + "\n")
            return []

        def lerror(a):
            raise LuaError(a[0] if a else None)

        def lassert(a):
            if not a or a[0] is None or a[0] is False:
                raise LuaError(a[1] if len(a) > 1 else b"assertion failed!")
            return a

        def lpcall(a):
            try:
                return [True] + call(a[0], a[1:])
            except LuaError as ex:
                return [False, ex.value]
            except RecursionError:
                return [False, b"stack overflow"]

        def lselect(a):
            n = a[0]
            if n == b"#":
                return [float(len(a) - 1)]
            n = int(tonum(n))
            if n < 0:
                n = len(a) + n
                return a[n:]
            return a[n:]

        def lsetmetatable(a):
            if type(a[0]) is not LuaTable:
                raise LuaError(b"bad argument #1 to 'setmetatable' (table expected)")
            a[0].meta = a[1] if len(a) > 1 else None
            return [a[0]]

        def lgetmetatable(a):
            o = a[0] if a else None
            if type(o) is LuaTable:
                if o.meta is not None:
                    protected = o.meta.get(b"__metatable")
                    return [protected if protected is not None else o.meta]
                return [None]
            if type(o) is bytes:
                mt = LuaTable()
                mt.set(b"__index", STRING_LIB)
                return [mt]
            return [None]

        def lnext(a):
            return a[0].next(a[1] if len(a) > 1 else None)
        nextf = _fn("next", lnext)

        def lpairs(a):
            o = a[0]
            if type(o) is LuaTable and o.meta is not None and o.meta.get(b"__pairs") is not None:
                return call(o.meta.get(b"__pairs"), [o])[:3]
            if type(o) is not LuaTable:
                raise LuaError(("bad argument #1 to 'pairs' (table expected, got %s)" % type_name(o)).encode())
            return [nextf, o, None]

        def ipairs_iter(a):
            t, i = a[0], a[1] + 1
            v = index(t, i)
            if v is None:
                return [None]
            return [i, v]
        ipf = _fn("ipairs_iter", ipairs_iter)

        def ltonumber(a):
            v = a[0] if a else None
            if len(a) > 1 and a[1] is not None:
                try:
                    return [float(int(tostr(v).decode("latin-1").strip(), int(tonum(a[1]))))]
                except ValueError:
                    return [None]
            return [tonum(v)]

        def lrequire(a):
            name = a[0]
            if name in self.loaded:
                return [self.loaded[name]]
            if name in self.preload:
                v = self.preload[name]()
                self.loaded[name] = v
                return [v]
            raise LuaError(b"module '" + name + b"' not found")

        def lrawlen(a):
            return [float(len(a[0].arr)) if type(a[0]) is LuaTable else float(len(a[0]))]
        os_t = LuaTable()
        os_t.set(b"clock", _fn("os.clock", lambda a: [time.process_time()]))
        os_t.set(b"time", _fn("os.time", lambda a: [float(int(time.time()))]))
        os_t.set(b"epoch", _fn("os.epoch", lambda a: [0.0]))   # constant: the reference's 3 s yield check never fires
        os_t.set(b"queueEvent", _fn("os.queueEvent", lambda a: []))
        os_t.set(b"pullEvent", _fn("os.pullEvent", lambda a: [b"nosleep"]))
        G.set(b"os", os_t)
        for k, f in {"print": lprint, "error": lerror, "assert": lassert, "pcall": lpcall, "select": lselect,
                     "setmetatable": lsetmetatable, "getmetatable": lgetmetatable, "next": lnext, "pairs": lpairs,
                     "ipairs": lambda a: [ipf, a[0], 0.0], "tonumber": ltonumber, "tostring": lambda a: [tostr(a[0] if a else None)],
                     "type": lambda a: [type_name(a[0] if a else None).encode()], "require": lrequire,
                     "rawget": lambda a: [a[0].get(a[1])], "rawset": lambda a: (a[0].set(a[1], a[2]), [a[0]])[1],
                     "rawequal": lambda a: [a[0] is a[1] or (type(a[0]) in (float, bytes) and type(a[0]) is type(a[1]) and a[0] == a[1])],
                     "rawlen": lrawlen, "sleep": lambda a: []}.items():
            G.set(k.encode(), _fn(k, f))

    def load(self, src, chunkname="chunk"):
        if isinstance(src, str):
            src = src.encode("latin-1")
        ast = Parser(tokenize(src, chunkname), chunkname).block()
        comp = Compiler(self, chunkname)
        fs = FuncState(None)
        block = comp.block(fs, ast, new_scope=False)

        def main(args):
            frame = [None] * max(fs.nslots, 1)
            r = block(frame, [], list(args))
            return [] if r is None else r[1]
        return LuaFunction(main, chunkname)

    def run(self, src, chunkname="chunk", args=()):
        return self.load(src, chunkname).call(list(args))


# ---- conversions for the Python side ----
def to_lua(v):
    if isinstance(v, bool) or v is None:
        return v
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, str):
        return v.encode("latin-1")
    if isinstance(v, (bytes, bytearray)):
        return bytes(v)
    if isinstance(v, (list, tuple)):
        t = LuaTable()
        t.arr = [to_lua(x) for x in v]
        return t
    if isinstance(v, dict):
        t = LuaTable()
        for k, x in v.items():
            t.set(to_lua(k), to_lua(x))
        return t
    return v
