/*
 * luahost.c -- TEST INFRASTRUCTURE: a toy host for Lua 5.2 C modules.
 *
 * No Lua interpreter exists in the build image, so aukit_b200/csrc/lua_binding.c (luaopen_aukit_cuda) could only be
 * compiled, never run.  This file implements the part of the Lua 5.2 C API that the binding declares -- value stack,
 * strings, tables, full userdata with metatables and __gc, luaL_check* argument errors, lua_error as a longjmp to the
 * protected call -- with the semantics of lua.h / lauxlib.h, so the binding's own object code runs unmodified:
 *
 *     libluahost.so (this file, loaded RTLD_GLOBAL)  <-  dlopen(aukit_cuda.so)  ->  luaopen_aukit_cuda(L)
 *
 * The driver (tests/luahost.py, ctypes) pushes arguments with the lua_* calls below, invokes a module function with
 * lh_call(), and reads the results back.  It is not a Lua implementation: there is no parser, no VM, no closures; the
 * Lua FACADE (aukit_b200/lua/aukit.lua) still runs in oracle/luavm, which calls the module through this host.
 * Memory: reference counts (tables cannot form cycles here); a userdata's __gc runs when its last reference goes.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct lua_State lua_State;
typedef double lua_Number;
typedef ptrdiff_t lua_Integer;
typedef int (*lua_CFunction)(lua_State *L);
typedef struct luaL_Reg { const char *name; lua_CFunction func; } luaL_Reg;

enum { LUA_TNONE = -1, LUA_TNIL = 0, LUA_TBOOLEAN = 1, LUA_TLIGHTUSERDATA = 2, LUA_TNUMBER = 3, LUA_TSTRING = 4,
       LUA_TTABLE = 5, LUA_TFUNCTION = 6, LUA_TUSERDATA = 7 };
static const char *const type_names[] = {"nil", "boolean", "userdata", "number", "string", "table", "function", "userdata"};

typedef struct obj obj;
typedef struct value {
    int t;
    union { int b; lua_Number n; obj *o; lua_CFunction f; } u;
} value;

typedef struct field { char *key; value v; struct field *next; } field;

struct obj {
    int t;              /* LUA_TSTRING, LUA_TTABLE, LUA_TUSERDATA */
    long refs;
    /* string */
    size_t len;
    char *s;
    /* table */
    value *arr;         /* arr[i - 1] = t[i], i = 1 .. narr */
    size_t narr, cap;
    field *fields;
    /* userdata */
    void *block;
    size_t size;
    obj *meta;
};

#define STACK_MAX 4096
#define MAX_METAS 16
#define MAX_REFS 65536

struct lua_State {
    value stack[STACK_MAX];
    int base, top;                 /* stack[base .. top) is the running function's frame */
    jmp_buf *jmp;
    char err[1024];
    struct { char name[64]; obj *mt; } metas[MAX_METAS];
    int nmetas;
    value refs[MAX_REFS];
    int ref_free[MAX_REFS], nfree, nrefs;
    value module;
    void *dl;
    long gc_calls, live_udata;
};

/* ------------------------------------------------------------------ values */
static void release(lua_State *L, value v);

static value nilv(void) { value v; v.t = LUA_TNIL; v.u.n = 0; return v; }
static int is_obj(value v) { return v.t == LUA_TSTRING || v.t == LUA_TTABLE || v.t == LUA_TUSERDATA; }
static value retain(value v) { if (is_obj(v)) v.u.o->refs++; return v; }

static void throw_msg(lua_State *L, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(L->err, sizeof L->err, fmt, ap);
    va_end(ap);
    if (!L->jmp) { fprintf(stderr, "luahost: unprotected error: %s\n", L->err); abort(); }
    longjmp(*L->jmp, 1);
}

static int call_function(lua_State *L, lua_CFunction f, int nargs);

static void free_obj(lua_State *L, obj *o) {
    if (o->t == LUA_TUSERDATA) {
        /* __gc, as lua_close / the collector would run it: the userdata is alive for the duration of the call */
        if (o->meta) {
            for (field *f = o->meta->fields; f; f = f->next)
                if (!strcmp(f->key, "__gc") && f->v.t == LUA_TFUNCTION) {
                    value self; self.t = LUA_TUSERDATA; self.u.o = o;
                    o->refs = 1 << 20;                      /* resurrected while the finaliser runs */
                    if (L->top >= STACK_MAX - 8) break;
                    L->stack[L->top++] = self;
                    jmp_buf jb, *prev = L->jmp;
                    const int base = L->base, top = L->top - 1;
                    L->jmp = &jb;
                    if (setjmp(jb) == 0) call_function(L, f->v.u.f, 1);
                    else fprintf(stderr, "luahost: error in __gc: %s\n", L->err);
                    L->jmp = prev;
                    /* drop whatever the finaliser left, without releasing `self` again */
                    while (L->top > top) { value x = L->stack[--L->top]; if (!(x.t == LUA_TUSERDATA && x.u.o == o)) release(L, x); }
                    L->base = base;
                    L->gc_calls++;
                    break;
                }
            value m; m.t = LUA_TTABLE; m.u.o = o->meta;
            release(L, m);
        }
        L->live_udata--;
        free(o->block);
    } else if (o->t == LUA_TTABLE) {
        for (size_t i = 0; i < o->narr; i++) release(L, o->arr[i]);
        free(o->arr);
        for (field *f = o->fields; f;) { field *n = f->next; release(L, f->v); free(f->key); free(f); f = n; }
    } else {
        free(o->s);
    }
    free(o);
}

static void release(lua_State *L, value v) {
    if (is_obj(v) && --v.u.o->refs == 0) free_obj(L, v.u.o);
}

static obj *new_obj(lua_State *L, int t) {
    obj *o = (obj *)calloc(1, sizeof(obj));
    if (!o) throw_msg(L, "not enough memory");
    o->t = t;
    return o;
}

static value *slot(lua_State *L, int idx) {              /* NULL: not an acceptable index (LUA_TNONE) */
    if (idx > 0) { const int i = L->base + idx - 1; return i < L->top ? &L->stack[i] : NULL; }
    if (idx < 0) { const int i = L->top + idx; return i >= L->base ? &L->stack[i] : NULL; }
    return NULL;
}

static value *need(lua_State *L, int idx) {
    value *v = slot(L, idx);
    if (!v) throw_msg(L, "luahost: invalid stack index %d", idx);
    return v;
}

static void push(lua_State *L, value v) {
    if (L->top >= STACK_MAX) throw_msg(L, "stack overflow");
    L->stack[L->top++] = v;
}

static obj *table_at(lua_State *L, int idx) {
    value *v = need(L, idx);
    if (v->t != LUA_TTABLE) throw_msg(L, "luahost: table expected at index %d, got %s", idx, type_names[v->t]);
    return v->u.o;
}

static void table_seti(lua_State *L, obj *t, size_t i, value v) {   /* takes ownership of v */
    if (i < 1) { release(L, v); throw_msg(L, "luahost: table index %zu not supported", i); }
    if (i > t->narr) {
        if (v.t == LUA_TNIL) return;
        if (i > t->cap) {
            size_t cap = t->cap ? t->cap : 4;
            while (cap < i) cap *= 2;
            value *a = (value *)realloc(t->arr, cap * sizeof(value));
            if (!a) { release(L, v); throw_msg(L, "not enough memory"); }
            t->arr = a; t->cap = cap;
        }
        for (size_t k = t->narr; k < i; k++) t->arr[k] = nilv();
        t->narr = i;
    }
    release(L, t->arr[i - 1]);
    t->arr[i - 1] = v;
    while (t->narr && t->arr[t->narr - 1].t == LUA_TNIL) t->narr--;   /* keep narr a border */
}

static void table_setfield(lua_State *L, obj *t, const char *k, value v) {   /* takes ownership of v */
    for (field *f = t->fields; f; f = f->next)
        if (!strcmp(f->key, k)) { release(L, f->v); f->v = v; return; }
    field *f = (field *)calloc(1, sizeof(field));
    f->key = strdup(k); f->v = v; f->next = NULL;
    field **tail = &t->fields;                 /* insertion order, so the driver can walk fields deterministically */
    while (*tail) tail = &(*tail)->next;
    *tail = f;
}

/* ------------------------------------------------------------------ lua.h subset */
int lua_gettop(lua_State *L) { return L->top - L->base; }

void lua_settop(lua_State *L, int idx) {
    int newtop = idx >= 0 ? L->base + idx : L->top + idx + 1;
    if (newtop < L->base || newtop > STACK_MAX) throw_msg(L, "luahost: lua_settop(%d) out of range", idx);
    while (L->top > newtop) release(L, L->stack[--L->top]);
    while (L->top < newtop) L->stack[L->top++] = nilv();
}

int lua_type(lua_State *L, int idx) { value *v = slot(L, idx); return v ? v->t : LUA_TNONE; }
int lua_toboolean(lua_State *L, int idx) { value *v = slot(L, idx); return v && !(v->t == LUA_TNIL || (v->t == LUA_TBOOLEAN && !v->u.b)); }

static int str2number(const obj *s, lua_Number *out) {
    char *end;
    if (!s->len) return 0;
    const double d = strtod(s->s, &end);
    while (*end == ' ' || *end == '\t' || *end == '\n') end++;
    if (end != s->s + s->len) return 0;
    *out = d;
    return 1;
}

lua_Number lua_tonumberx(lua_State *L, int idx, int *isnum) {
    value *v = slot(L, idx);
    lua_Number n = 0;
    int ok = 0;
    if (v && v->t == LUA_TNUMBER) { n = v->u.n; ok = 1; }
    else if (v && v->t == LUA_TSTRING) ok = str2number(v->u.o, &n);
    if (isnum) *isnum = ok;
    return ok ? n : 0;
}

void *lua_touserdata(lua_State *L, int idx) { value *v = slot(L, idx); return v && v->t == LUA_TUSERDATA ? v->u.o->block : NULL; }

size_t lua_rawlen(lua_State *L, int idx) {
    value *v = slot(L, idx);
    if (!v) return 0;
    if (v->t == LUA_TSTRING) return v->u.o->len;
    if (v->t == LUA_TTABLE) return v->u.o->narr;
    if (v->t == LUA_TUSERDATA) return v->u.o->size;
    return 0;
}

void lua_pushnil(lua_State *L) { push(L, nilv()); }
void lua_pushnumber(lua_State *L, lua_Number n) { value v; v.t = LUA_TNUMBER; v.u.n = n; push(L, v); }
void lua_pushinteger(lua_State *L, lua_Integer n) { lua_pushnumber(L, (lua_Number)n); }
void lua_pushboolean(lua_State *L, int b) { value v; v.t = LUA_TBOOLEAN; v.u.b = b != 0; push(L, v); }
void lua_pushcclosure(lua_State *L, lua_CFunction f, int n) { (void)n; value v; v.t = LUA_TFUNCTION; v.u.f = f; push(L, v); }

const char *lua_pushlstring(lua_State *L, const char *s, size_t len) {
    obj *o = new_obj(L, LUA_TSTRING);
    o->s = (char *)malloc(len + 1);
    if (!o->s) { free(o); throw_msg(L, "not enough memory"); }
    if (len) memcpy(o->s, s, len);
    o->s[len] = 0; o->len = len; o->refs = 1;
    value v; v.t = LUA_TSTRING; v.u.o = o;
    push(L, v);
    return o->s;
}

const char *lua_pushstring(lua_State *L, const char *s) {
    if (!s) { lua_pushnil(L); return NULL; }
    return lua_pushlstring(L, s, strlen(s));
}

void lua_pushvalue(lua_State *L, int idx) { push(L, retain(*need(L, idx))); }

void lua_createtable(lua_State *L, int narr, int nrec) {
    (void)nrec;
    obj *o = new_obj(L, LUA_TTABLE);
    o->refs = 1;
    if (narr > 0) { o->arr = (value *)malloc(sizeof(value) * (size_t)narr); o->cap = o->arr ? (size_t)narr : 0; }
    value v; v.t = LUA_TTABLE; v.u.o = o;
    push(L, v);
}

void lua_setfield(lua_State *L, int idx, const char *k) {           /* t[k] = top; pops */
    obj *t = table_at(L, idx);
    if (L->top <= L->base) throw_msg(L, "luahost: lua_setfield on an empty stack");
    table_setfield(L, t, k, L->stack[--L->top]);
}

void lua_getfield(lua_State *L, int idx, const char *k) {
    obj *t = table_at(L, idx);
    for (field *f = t->fields; f; f = f->next)
        if (!strcmp(f->key, k)) { push(L, retain(f->v)); return; }
    lua_pushnil(L);
}

void lua_rawseti(lua_State *L, int idx, int n) {                    /* t[n] = top; pops */
    obj *t = table_at(L, idx);
    if (L->top <= L->base) throw_msg(L, "luahost: lua_rawseti on an empty stack");
    const value v = L->stack[--L->top];
    table_seti(L, t, (size_t)n, v);
}

void lua_rawgeti(lua_State *L, int idx, int n) {
    obj *t = table_at(L, idx);
    if (n >= 1 && (size_t)n <= t->narr) push(L, retain(t->arr[n - 1]));
    else lua_pushnil(L);
}

void *lua_newuserdata(lua_State *L, size_t sz) {
    obj *o = new_obj(L, LUA_TUSERDATA);
    o->block = calloc(1, sz ? sz : 1);
    o->size = sz; o->refs = 1;
    L->live_udata++;
    value v; v.t = LUA_TUSERDATA; v.u.o = o;
    push(L, v);
    return o->block;
}

const char *lua_tolstring(lua_State *L, int idx, size_t *len) {
    value *v = slot(L, idx);
    if (v && v->t == LUA_TNUMBER) {                                  /* converted in place, as lua_tolstring does */
        char buf[64];
        const int n = snprintf(buf, sizeof buf, "%.14g", v->u.n);
        lua_pushlstring(L, buf, (size_t)n);
        const value s = L->stack[--L->top];
        *v = s;
    }
    if (!v || v->t != LUA_TSTRING) { if (len) *len = 0; return NULL; }
    if (len) *len = v->u.o->len;
    return v->u.o->s;
}

int lua_error(lua_State *L) {
    size_t n = 0;
    const char *s = L->top > L->base ? lua_tolstring(L, -1, &n) : NULL;
    char msg[1024];
    snprintf(msg, sizeof msg, "%s", s ? s : "(error object is not a string)");
    throw_msg(L, "%s", msg);
    return 0;
}

/* ------------------------------------------------------------------ lauxlib.h subset */
int luaL_error(lua_State *L, const char *fmt, ...) {
    /* luaL_error understands %s %d %f %c %% (lua_pushvfstring); vsnprintf is a superset.  No "file:line:" position is
     * prepended: luaL_where(L, 1) is empty for a C function called from the host */
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(L->err, sizeof L->err, fmt, ap);
    va_end(ap);
    if (!L->jmp) { fprintf(stderr, "luahost: unprotected error: %s\n", L->err); abort(); }
    longjmp(*L->jmp, 1);
    return 0;
}

static int arg_error(lua_State *L, int arg, const char *extra) { return luaL_error(L, "bad argument #%d to '?' (%s)", arg, extra); }

static int type_error(lua_State *L, int arg, const char *tname) {
    char extra[128];
    const int t = lua_type(L, arg);
    snprintf(extra, sizeof extra, "%s expected, got %s", tname, t == LUA_TNONE ? "no value" : type_names[t]);
    return arg_error(L, arg, extra);
}

const char *luaL_checklstring(lua_State *L, int arg, size_t *l) {
    const char *s = lua_tolstring(L, arg, l);
    if (!s) type_error(L, arg, "string");
    return s;
}

lua_Number luaL_checknumber(lua_State *L, int arg) {
    int ok;
    const lua_Number n = lua_tonumberx(L, arg, &ok);
    if (!ok) type_error(L, arg, "number");
    return n;
}

lua_Number luaL_optnumber(lua_State *L, int arg, lua_Number def) { return lua_type(L, arg) <= LUA_TNIL ? def : luaL_checknumber(L, arg); }
lua_Integer luaL_checkinteger(lua_State *L, int arg) { return (lua_Integer)luaL_checknumber(L, arg); }
lua_Integer luaL_optinteger(lua_State *L, int arg, lua_Integer def) { return lua_type(L, arg) <= LUA_TNIL ? def : luaL_checkinteger(L, arg); }

static obj *find_meta(lua_State *L, const char *tname) {
    for (int i = 0; i < L->nmetas; i++) if (!strcmp(L->metas[i].name, tname)) return L->metas[i].mt;
    return NULL;
}

int luaL_newmetatable(lua_State *L, const char *tname) {
    obj *mt = find_meta(L, tname);
    if (mt) { value v; v.t = LUA_TTABLE; v.u.o = mt; push(L, retain(v)); return 0; }
    if (L->nmetas >= MAX_METAS) throw_msg(L, "luahost: too many metatables");
    lua_createtable(L, 0, 4);
    mt = L->stack[L->top - 1].u.o;
    mt->refs++;                                                       /* the registry's reference */
    snprintf(L->metas[L->nmetas].name, sizeof L->metas[0].name, "%s", tname);
    L->metas[L->nmetas++].mt = mt;
    return 1;
}

void luaL_setmetatable(lua_State *L, const char *tname) {
    value *v = need(L, -1);
    obj *mt = find_meta(L, tname);
    if (v->t != LUA_TUSERDATA) throw_msg(L, "luahost: luaL_setmetatable on a %s", type_names[v->t]);
    if (v->u.o->meta) { value m; m.t = LUA_TTABLE; m.u.o = v->u.o->meta; release(L, m); }
    v->u.o->meta = mt;
    if (mt) mt->refs++;
}

void *luaL_checkudata(lua_State *L, int ud, const char *tname) {
    value *v = slot(L, ud);
    if (!v || v->t != LUA_TUSERDATA || !v->u.o->meta || v->u.o->meta != find_meta(L, tname)) type_error(L, ud, tname);
    return v->u.o->block;
}

void luaL_setfuncs(lua_State *L, const luaL_Reg *l, int nup) {
    if (nup) throw_msg(L, "luahost: upvalues are not supported");
    obj *t = table_at(L, -1);
    for (; l->name; l++) { value v; v.t = LUA_TFUNCTION; v.u.f = l->func; table_setfield(L, t, l->name, v); }
}

/* ------------------------------------------------------------------ calls */
/* Calls f with the top nargs values as its frame; on return the results replace function arguments. */
static int call_function(lua_State *L, lua_CFunction f, int nargs) {
    const int prev_base = L->base;
    const int frame = L->top - nargs;
    L->base = frame;
    const int nres = f(L);
    if (nres < 0 || nres > L->top - L->base) throw_msg(L, "luahost: C function returned %d results with %d values on its stack", nres, L->top - L->base);
    const int first = L->top - nres;
    for (int i = frame; i < first; i++) release(L, L->stack[i]);
    for (int i = 0; i < nres; i++) L->stack[frame + i] = L->stack[first + i];
    L->top = frame + nres;
    L->base = prev_base;
    return nres;
}

/* ------------------------------------------------------------------ driver API (tests/luahost.py) */
lua_State *lh_new(void) {
    lua_State *L = (lua_State *)calloc(1, sizeof(lua_State));
    if (L) L->module = nilv();
    return L;
}

const char *lh_error(lua_State *L) { return L->err; }
long lh_gc_calls(lua_State *L) { return L->gc_calls; }
long lh_live_userdata(lua_State *L) { return L->live_udata; }

/* dlopen(path), call `sym` (a luaopen_* function) in protected mode, keep the module table it returns */
int lh_open(lua_State *L, const char *path, const char *sym) {
    L->dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!L->dl) { snprintf(L->err, sizeof L->err, "dlopen: %s", dlerror()); return -1; }
    lua_CFunction f = (lua_CFunction)dlsym(L->dl, sym);
    if (!f) { snprintf(L->err, sizeof L->err, "dlsym: %s", dlerror()); return -1; }
    jmp_buf jb;
    L->jmp = &jb;
    const int base = L->base, top = L->top;
    if (setjmp(jb)) {
        while (L->top > top) release(L, L->stack[--L->top]);
        L->base = base; L->jmp = NULL;
        return -1;
    }
    const int n = call_function(L, f, 0);
    if (n != 1 || L->stack[L->top - 1].t != LUA_TTABLE) throw_msg(L, "luahost: %s did not return a table", sym);
    release(L, L->module);
    L->module = L->stack[--L->top];
    L->jmp = NULL;
    return 0;
}

/* pushes module[name]; returns its type */
int lh_getmodule(lua_State *L, const char *name) {
    if (L->module.t != LUA_TTABLE) { lua_pushnil(L); return LUA_TNIL; }
    for (field *f = L->module.u.o->fields; f; f = f->next)
        if (!strcmp(f->key, name)) { push(L, retain(f->v)); return f->v.t; }
    lua_pushnil(L);
    return LUA_TNIL;
}

/* names of the module's fields, in registration order; NULL past the end */
const char *lh_module_key(lua_State *L, int i) {
    if (L->module.t != LUA_TTABLE) return NULL;
    field *f = L->module.u.o->fields;
    while (f && i-- > 0) f = f->next;
    return f ? f->key : NULL;
}

/* module[name](top nargs values) in protected mode: the number of results (left on the stack), or -1 + lh_error() */
int lh_call(lua_State *L, const char *name, int nargs) {
    jmp_buf jb;
    const int base = L->base, floor_ = L->top - nargs;
    L->jmp = &jb;
    if (setjmp(jb)) {
        while (L->top > floor_) release(L, L->stack[--L->top]);
        L->base = base; L->jmp = NULL;
        return -1;
    }
    if (nargs < 0 || floor_ < L->base) throw_msg(L, "luahost: lh_call with %d arguments on a stack of %d", nargs, L->top - L->base);
    lua_CFunction f = NULL;
    if (L->module.t == LUA_TTABLE)
        for (field *fl = L->module.u.o->fields; fl; fl = fl->next)
            if (!strcmp(fl->key, name) && fl->v.t == LUA_TFUNCTION) f = fl->v.u.f;
    if (!f) throw_msg(L, "attempt to call field '%s' (a nil value)", name);
    const int n = call_function(L, f, nargs);
    L->jmp = NULL;
    return n;
}

/* protected wrappers for the stack calls the driver makes itself (an error in an unprotected call would abort) */
int lh_protected(lua_State *L, int op, int idx, int n, const char *k) {
    jmp_buf jb;
    const int base = L->base;
    L->jmp = &jb;
    if (setjmp(jb)) { L->base = base; L->jmp = NULL; return -1; }
    switch (op) {
        case 0: lua_settop(L, idx); break;
        case 1: lua_rawseti(L, idx, n); break;
        case 2: lua_rawgeti(L, idx, n); break;
        case 3: lua_setfield(L, idx, k); break;
        case 4: lua_getfield(L, idx, k); break;
        case 5: lua_createtable(L, n, 0); break;
        case 6: lua_pushvalue(L, idx); break;
        default: throw_msg(L, "luahost: unknown op %d", op);
    }
    L->jmp = NULL;
    return 0;
}

/* key of the i-th string field of the table at idx (insertion order), pushing its value; NULL past the end */
const char *lh_field(lua_State *L, int idx, int i) {
    value *v = slot(L, idx);
    if (!v || v->t != LUA_TTABLE) return NULL;
    field *f = v->u.o->fields;
    while (f && i-- > 0) f = f->next;
    if (!f) return NULL;
    push(L, retain(f->v));
    return f->key;
}

/* bulk transfers for the driver: the array part of the table at idx as doubles (-1 if an element is not a number) ... */
long lh_array_numbers(lua_State *L, int idx, double *out, long cap) {
    value *v = slot(L, idx);
    if (!v || v->t != LUA_TTABLE) return -1;
    const obj *t = v->u.o;
    if ((long)t->narr > cap) return -1;
    for (size_t i = 0; i < t->narr; i++) {
        if (t->arr[i].t != LUA_TNUMBER) return -1;
        out[i] = t->arr[i].u.n;
    }
    return (long)t->narr;
}

/* ... and a new table {v[0], v[1], ...} pushed on the stack */
int lh_push_number_array(lua_State *L, const double *v, long n) {
    if (L->top >= STACK_MAX) return -1;
    obj *o = (obj *)calloc(1, sizeof(obj));
    if (!o) return -1;
    o->t = LUA_TTABLE; o->refs = 1;
    if (n > 0) {
        o->arr = (value *)malloc(sizeof(value) * (size_t)n);
        if (!o->arr) { free(o); return -1; }
        o->cap = o->narr = (size_t)n;
        for (long i = 0; i < n; i++) { o->arr[i].t = LUA_TNUMBER; o->arr[i].u.n = v[i]; }
    }
    value tv; tv.t = LUA_TTABLE; tv.u.o = o;
    L->stack[L->top++] = tv;
    return 0;
}

/* keeps the value at idx alive outside the stack (what a Lua variable holding it would do) */
int lh_ref(lua_State *L, int idx) {
    value *v = slot(L, idx);
    if (!v) return -1;
    int id;
    if (L->nfree) id = L->ref_free[--L->nfree];
    else if (L->nrefs < MAX_REFS) id = L->nrefs++;
    else return -1;
    L->refs[id] = retain(*v);
    return id;
}

int lh_pushref(lua_State *L, int id) {
    if (id < 0 || id >= L->nrefs || L->top >= STACK_MAX) return -1;
    L->stack[L->top++] = retain(L->refs[id]);
    return 0;
}

void lh_unref(lua_State *L, int id) {
    if (id < 0 || id >= L->nrefs) return;
    const value v = L->refs[id];
    L->refs[id] = nilv();
    L->ref_free[L->nfree++] = id;
    release(L, v);                                                   /* last reference to a userdata: __gc runs here */
}

/* lua_close: everything goes, finalisers run */
void lh_close(lua_State *L) {
    if (!L) return;
    L->base = 0;
    while (L->top > 0) release(L, L->stack[--L->top]);
    for (int i = 0; i < L->nrefs; i++) { const value v = L->refs[i]; L->refs[i] = nilv(); release(L, v); }
    release(L, L->module);
    L->module = nilv();
    for (int i = 0; i < L->nmetas; i++) { value m; m.t = LUA_TTABLE; m.u.o = L->metas[i].mt; release(L, m); }
    /* the module stays mapped: unloading a CUDA-using library at this point buys nothing */
    free(L);
}
