/*
 * lua_binding.c -- luaopen_aukit_cuda: the thin Lua 5.2 C module over libaukit_cuda.so.
 *
 *   local cu = require "aukit_cuda"      -- aukit_b200/lib/aukit_cuda.so on package.cpath
 *
 * aukit_b200/lua/aukit.lua (the facade with the reference's API) is the only intended caller.
 * Every function maps 1:1 onto one entry point of include/aukit_cuda.h; failures raise Lua
 * errors carrying the C library's message, which already uses the reference's error strings.
 *
 * No lua.h exists in the build image, so the handful of Lua 5.2 C-API prototypes used here are
 * declared by hand below (they match lua.h / lauxlib.h of Lua 5.2.x: lua_Number = double,
 * lua_Integer = ptrdiff_t).  The symbols stay undefined in this shared object and resolve
 * against the host interpreter when it dlopen()s the module, which is how every Lua C module
 * links.  This file is compile-checked here but can only be EXERCISED where a Lua 5.2
 * interpreter exists (none does in this image).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/aukit_cuda.h"

/* ---- Lua 5.2 C API subset (hand-declared) ---- */
typedef struct lua_State lua_State;
typedef double lua_Number;
typedef ptrdiff_t lua_Integer;
typedef int (*lua_CFunction)(lua_State *L);
typedef struct luaL_Reg { const char *name; lua_CFunction func; } luaL_Reg;
#define LUA_TNIL 0
#define LUA_TBOOLEAN 1
#define LUA_TNUMBER 3
#define LUA_TTABLE 5
extern int lua_gettop(lua_State *L);
extern void lua_settop(lua_State *L, int idx);
extern int lua_type(lua_State *L, int idx);
extern int lua_toboolean(lua_State *L, int idx);
extern lua_Number lua_tonumberx(lua_State *L, int idx, int *isnum);
extern void *lua_touserdata(lua_State *L, int idx);
extern size_t lua_rawlen(lua_State *L, int idx);
extern void lua_pushnil(lua_State *L);
extern void lua_pushnumber(lua_State *L, lua_Number n);
extern void lua_pushinteger(lua_State *L, lua_Integer n);
extern const char *lua_pushlstring(lua_State *L, const char *s, size_t len);
extern const char *lua_pushstring(lua_State *L, const char *s);
extern void lua_pushboolean(lua_State *L, int b);
extern void lua_pushvalue(lua_State *L, int idx);
extern void lua_createtable(lua_State *L, int narr, int nrec);
extern void lua_setfield(lua_State *L, int idx, const char *k);
extern void lua_rawseti(lua_State *L, int idx, int n);
extern void lua_rawgeti(lua_State *L, int idx, int n);
extern void *lua_newuserdata(lua_State *L, size_t sz);
extern int lua_error(lua_State *L);
extern const char *luaL_checklstring(lua_State *L, int arg, size_t *l);
extern lua_Number luaL_checknumber(lua_State *L, int arg);
extern lua_Number luaL_optnumber(lua_State *L, int arg, lua_Number def);
extern lua_Integer luaL_checkinteger(lua_State *L, int arg);
extern lua_Integer luaL_optinteger(lua_State *L, int arg, lua_Integer def);
extern void *luaL_checkudata(lua_State *L, int ud, const char *tname);
extern int luaL_newmetatable(lua_State *L, const char *tname);
extern void luaL_setmetatable(lua_State *L, const char *tname);
extern void luaL_setfuncs(lua_State *L, const luaL_Reg *l, int nup);
extern int luaL_error(lua_State *L, const char *fmt, ...);

#define AUDIO_MT "aukit_cuda.Audio"

typedef struct { aukit_audio *a; } audio_ud;

static aukit_ctx *g_ctx; /* single Lua coroutine, single device (no threads in the reference) */

static aukit_ctx *ctx(lua_State *L) {
    if (!g_ctx && aukit_cuda_init(-1, &g_ctx)) luaL_error(L, "%s", aukit_cuda_last_error());
    return g_ctx;
}

static int fail(lua_State *L) { return luaL_error(L, "%s", aukit_cuda_last_error()); }

static aukit_audio *check_audio(lua_State *L, int idx) {
    audio_ud *u = (audio_ud *)luaL_checkudata(L, idx, AUDIO_MT);
    if (!u->a) luaL_error(L, "aukit_cuda: Audio handle already released");
    return u->a;
}

static int push_audio(lua_State *L, aukit_audio *a) {
    audio_ud *u = (audio_ud *)lua_newuserdata(L, sizeof *u);
    u->a = a;
    luaL_setmetatable(L, AUDIO_MT);
    return 1;
}

static int optbool(lua_State *L, int idx, int def) { return lua_type(L, idx) <= LUA_TNIL ? def : lua_toboolean(L, idx); }

/* int table (1-based) at idx -> malloc'd array of n ints, or NULL when nil */
static int *int_table(lua_State *L, int idx, int n) {
    if (lua_type(L, idx) != LUA_TTABLE) return NULL;
    int *v = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        lua_rawgeti(L, idx, i + 1);
        v[i] = (int)lua_tonumberx(L, -1, NULL);
        lua_settop(L, -2);
    }
    return v;
}

/* cu.pcm(data, bitDepth, dataType(0 signed,1 unsigned,2 float), channels, sampleRate, interleaved, bigEndian) */
static int l_pcm(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_audio *a = NULL;
    if (aukit_cuda_pcm(ctx(L), d, n, (int)luaL_optinteger(L, 2, 8), (int)luaL_optinteger(L, 3, 0), (int)luaL_optinteger(L, 4, 1),
                       luaL_optnumber(L, 5, 48000), optbool(L, 6, 1), optbool(L, 7, 0), &a))
        return fail(L);
    return push_audio(L, a);
}

/* cu.g711(data, ulaw, channels, sampleRate) */
static int l_g711(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_audio *a = NULL;
    if (aukit_cuda_g711(ctx(L), d, n, lua_toboolean(L, 2), (int)luaL_optinteger(L, 3, 1), luaL_optnumber(L, 4, 8000), &a)) return fail(L);
    return push_audio(L, a);
}

/* cu.adpcm(data, channels, sampleRate, topFirst, interleaved, predictor{}, step_index{}) */
static int l_adpcm(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    const int ch = (int)luaL_optinteger(L, 2, 1);
    int *pr = int_table(L, 6, ch), *si = int_table(L, 7, ch);
    aukit_audio *a = NULL;
    const int rc = aukit_cuda_adpcm(ctx(L), d, n, ch, luaL_optnumber(L, 3, 48000), optbool(L, 4, 1), optbool(L, 5, 1), pr, si, &a);
    free(pr); free(si);
    if (rc) return fail(L);
    return push_audio(L, a);
}

/* cu.ima_adpcm_wav(data, blockAlign, channels, sampleRate, dialect) */
static int l_ima_wav(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_audio *a = NULL;
    if (aukit_cuda_ima_adpcm_wav(ctx(L), d, n, (int)luaL_checkinteger(L, 2), (int)luaL_optinteger(L, 3, 1), luaL_optnumber(L, 4, 48000),
                                 (int)luaL_optinteger(L, 5, 0), &a))
        return fail(L);
    return push_audio(L, a);
}

/* cu.msadpcm(data, blockAlign, channels, sampleRate, coef1{}, coef2{}, dialect) */
static int l_msadpcm(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    const int nco = lua_type(L, 5) == LUA_TTABLE ? (int)lua_rawlen(L, 5) : 0;
    int *c1 = int_table(L, 5, nco), *c2 = int_table(L, 6, nco);
    aukit_audio *a = NULL;
    const int rc = aukit_cuda_msadpcm(ctx(L), d, n, (int)luaL_checkinteger(L, 2), (int)luaL_optinteger(L, 3, 1),
                                      luaL_optnumber(L, 4, 48000), c1, c2, nco, (int)luaL_optinteger(L, 7, 0), &a);
    free(c1); free(c2);
    if (rc) return fail(L);
    return push_audio(L, a);
}

/* cu.wav(data, head, dialect) -> audio, info{format, channels, sampleRate, blockAlign, bitDepth, tags = {{id, value}, ...}} */
static int l_wav(lua_State *L) {
    static const char *fmt_names[] = {"signed", "unsigned", "float", "alaw", "ulaw", "adpcm", "msadpcm", "dfpwm", NULL};
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_wav_info *info = (aukit_wav_info *)malloc(sizeof *info);
    aukit_audio *a = NULL;
    if (!info) return luaL_error(L, "out of memory");
    /* the RIFF walk is host work: malformed files raise the reference's errors (A:1460-1507) before the device is touched */
    if (aukit_cuda_wav_parse(d, n, info)) { free(info); return fail(L); }
    if (aukit_cuda_wav(ctx(L), d, n, optbool(L, 2, 0), (int)luaL_optinteger(L, 3, 0), info, &a)) { free(info); return fail(L); }
    push_audio(L, a);
    lua_createtable(L, 0, 8);
    if (fmt_names[info->format]) { lua_pushstring(L, fmt_names[info->format]); lua_setfield(L, -2, "dataType"); }
    lua_pushinteger(L, info->channels); lua_setfield(L, -2, "channels");
    lua_pushinteger(L, info->sampleRate); lua_setfield(L, -2, "sampleRate");
    lua_pushinteger(L, info->blockAlign); lua_setfield(L, -2, "blockAlign");
    if (info->have_fmt) { lua_pushinteger(L, info->bitDepth); lua_setfield(L, -2, "bitDepth"); }
    lua_createtable(L, info->ntags, 0);
    for (int i = 0; i < info->ntags; i++) {
        lua_createtable(L, 2, 0);
        lua_pushstring(L, info->tags[i].id); lua_rawseti(L, -2, 1);
        lua_pushlstring(L, d + info->tags[i].off, info->tags[i].len); lua_rawseti(L, -2, 2);
        lua_rawseti(L, -2, i + 1);
    }
    lua_setfield(L, -2, "tags");
    free(info);
    return 2;
}

/* cu.au(data) / cu.aiff(data, head) -> audio, {codec, bitDepth, dataType, ulaw, meta = {{key, value}, ...}} */
static int push_container(lua_State *L, aukit_audio *a, const aukit_container_info *ci, const char *d) {
    static const char *dt_names[] = {"signed", "unsigned", "float"};
    push_audio(L, a);
    lua_createtable(L, 0, 6);
    lua_pushstring(L, ci->codec == AUKIT_CODEC_G711 ? "g711" : "pcm"); lua_setfield(L, -2, "codec");
    lua_pushinteger(L, ci->bitDepth); lua_setfield(L, -2, "bitDepth");
    lua_pushstring(L, dt_names[ci->dataType]); lua_setfield(L, -2, "dataType");
    lua_pushboolean(L, ci->ulaw); lua_setfield(L, -2, "ulaw");
    lua_createtable(L, ci->nmeta, 0);
    for (int i = 0; i < ci->nmeta; i++) {
        lua_createtable(L, 2, 0);
        lua_pushstring(L, ci->meta[i].key); lua_rawseti(L, -2, 1);
        lua_pushlstring(L, d + ci->meta[i].off, ci->meta[i].len); lua_rawseti(L, -2, 2);
        lua_rawseti(L, -2, i + 1);
    }
    lua_setfield(L, -2, "meta");
    return 2;
}

static int l_au(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_container_info ci;
    aukit_audio *a = NULL;
    if (aukit_cuda_au_parse(d, n, &ci)) return fail(L);           /* header errors before the device is touched */
    if (aukit_cuda_au(ctx(L), d, n, &ci, &a)) return fail(L);
    return push_container(L, a, &ci, d);
}

static int l_aiff(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_container_info ci;
    aukit_audio *a = NULL;
    if (aukit_cuda_aiff_parse(d, n, &ci)) return fail(L);
    if (aukit_cuda_aiff(ctx(L), d, n, optbool(L, 2, 0), &ci, &a)) return fail(L);
    return push_container(L, a, &ci, d);
}

/* cu.new(channels, frames, sampleRate) */
static int l_new(lua_State *L) {
    aukit_audio *a = NULL;
    if (aukit_cuda_audio_new(ctx(L), (int)luaL_checkinteger(L, 1), (size_t)luaL_checknumber(L, 2), luaL_checknumber(L, 3), &a)) return fail(L);
    return push_audio(L, a);
}

/* cu.resample(audio, sampleRate, interpolation(0 none, 1 linear, 2 cubic)) */
static int l_resample(lua_State *L) {
    aukit_audio *a = NULL;
    if (aukit_cuda_resample(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2), (int)luaL_checkinteger(L, 3), &a)) return fail(L);
    return push_audio(L, a);
}

static int l_mono(lua_State *L) {
    aukit_audio *a = NULL;
    if (aukit_cuda_mono(ctx(L), check_audio(L, 1), &a)) return fail(L);
    return push_audio(L, a);
}

/* cu.concat(a, b, ...) -- same-rate parts (the facade resamples first, A:702) */
static int l_concat(lua_State *L) {
    const int n = lua_gettop(L);
    const aukit_audio **parts = (const aukit_audio **)malloc(sizeof(*parts) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) parts[i] = check_audio(L, i + 1);
    aukit_audio *a = NULL;
    const int rc = aukit_cuda_concat(ctx(L), parts, n, &a);
    free(parts);
    if (rc) return fail(L);
    return push_audio(L, a);
}

/* in-place effects return nothing; the facade returns its own Audio table (A:3368, A:3458) */
static int l_amplify(lua_State *L) {
    if (aukit_cuda_amplify(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2))) return fail(L);
    return 0;
}

static int l_invert(lua_State *L) {
    if (aukit_cuda_invert(ctx(L), check_audio(L, 1))) return fail(L);
    return 0;
}

static int l_fade(lua_State *L) {
    if (aukit_cuda_fade(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2), luaL_checknumber(L, 3), luaL_checknumber(L, 4),
                        luaL_checknumber(L, 5))) return fail(L);
    return 0;
}

static int l_delay(lua_State *L) {
    if (aukit_cuda_delay(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2), luaL_optnumber(L, 3, 0.5))) return fail(L);
    return 0;
}

static int l_center(lua_State *L) {
    if (aukit_cuda_center(ctx(L), check_audio(L, 1))) return fail(L);
    return 0;
}

static int l_lowpass(lua_State *L) {
    if (aukit_cuda_lowpass(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2))) return fail(L);
    return 0;
}

static int l_highpass(lua_State *L) {
    if (aukit_cuda_highpass(ctx(L), check_audio(L, 1), luaL_checknumber(L, 2))) return fail(L);
    return 0;
}

static int l_normalize(lua_State *L) {
    if (aukit_cuda_normalize(ctx(L), check_audio(L, 1), luaL_optnumber(L, 2, 1.0), optbool(L, 3, 0))) return fail(L);
    return 0;
}

/* cu.set_sample_rate(audio, rate): audio.sampleRate is a plain writable field in the reference (A:3383) */
static int l_set_sample_rate(lua_State *L) {
    if (aukit_cuda_audio_set_sample_rate(check_audio(L, 1), luaL_checknumber(L, 2))) return fail(L);
    return 0;
}

/* cu.stream_chunk(audio, bitDepth, dataType, pos (1-based), count) -> {{ch1 values}, {ch2 values}, ...} or nil past the
 * end: one step of Audio:stream's iterator (A:921-937, encodePCM's `multiple` branch A:881-891) */
static int l_stream_chunk(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    const size_t first = (size_t)luaL_checknumber(L, 4) - 1, count = (size_t)luaL_checknumber(L, 5);
    const int nch = aukit_cuda_audio_channels(a);
    double *buf = (double *)malloc(sizeof(double) * (count ? count : 1) * (size_t)(nch > 0 ? nch : 1));
    size_t got = 0;
    if (!buf) return luaL_error(L, "out of memory");
    if (aukit_cuda_audio_stream_chunk(ctx(L), a, (int)luaL_checkinteger(L, 2), (int)luaL_checkinteger(L, 3), first, count, buf, &got)) {
        free(buf);
        return fail(L);
    }
    if (got == 0) { free(buf); lua_pushnil(L); return 1; }
    lua_createtable(L, nch, 0);
    for (int c = 0; c < nch; c++) {
        lua_createtable(L, (int)got, 0);
        for (size_t i = 0; i < got; i++) { lua_pushnumber(L, buf[(size_t)c * count + i]); lua_rawseti(L, -2, (int)i + 1); }
        lua_rawseti(L, -2, c + 1);
    }
    free(buf);
    return 1;
}

static int l_channels(lua_State *L) { lua_pushinteger(L, aukit_cuda_audio_channels(check_audio(L, 1))); return 1; }
static int l_sample_rate(lua_State *L) { lua_pushnumber(L, aukit_cuda_audio_sample_rate(check_audio(L, 1))); return 1; }

/* cu.frames(audio [, channel (1-based)]) */
static int l_frames(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    if (lua_type(L, 2) <= LUA_TNIL) lua_pushnumber(L, (lua_Number)aukit_cuda_audio_frames(a));
    else lua_pushnumber(L, (lua_Number)aukit_cuda_audio_channel_frames(a, (int)luaL_checkinteger(L, 2) - 1));
    return 1;
}

/* cu.read(audio, channel (1-based), first (1-based), count) -> array of numbers (D2H copy) */
static int l_read(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    const int c = (int)luaL_checkinteger(L, 2) - 1;
    const size_t first = (size_t)luaL_checknumber(L, 3) - 1, count = (size_t)luaL_checknumber(L, 4);
    float *buf = (float *)malloc(sizeof(float) * (count ? count : 1));
    if (!buf) return luaL_error(L, "out of memory");
    if (aukit_cuda_audio_download(ctx(L), a, c, first, count, buf)) { free(buf); return fail(L); }
    lua_createtable(L, (int)count, 0);
    for (size_t i = 0; i < count; i++) { lua_pushnumber(L, buf[i]); lua_rawseti(L, -2, (int)i + 1); }
    free(buf);
    return 1;
}

/* cu.pcm_out(audio, bitDepth, dataType, interleaved) -> {numbers}: Audio:pcm (A:901), un-rounded values */
static int l_pcm_out(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    const size_t total = aukit_cuda_audio_frames(a) * (size_t)aukit_cuda_audio_channels(a);
    double *buf = (double *)malloc(sizeof(double) * (total ? total : 1));
    if (!buf) return luaL_error(L, "out of memory");
    if (aukit_cuda_audio_pcm(ctx(L), a, (int)luaL_checkinteger(L, 2), (int)luaL_checkinteger(L, 3), optbool(L, 4, 1), buf)) {
        free(buf);
        return fail(L);
    }
    lua_createtable(L, (int)total, 0);
    for (size_t i = 0; i < total; i++) { lua_pushnumber(L, buf[i]); lua_rawseti(L, -2, (int)i + 1); }
    free(buf);
    return 1;
}

/* cu.pcm_bytes(audio, bitDepth, dataType, interleaved, rounding) -> string of packed little-endian samples */
static int l_pcm_bytes(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    const int bits = (int)luaL_checkinteger(L, 2);
    const size_t total = aukit_cuda_audio_frames(a) * (size_t)aukit_cuda_audio_channels(a) * (size_t)(bits > 0 ? bits / 8 : 0);
    char *buf = (char *)malloc(total ? total : 1);
    if (!buf) return luaL_error(L, "out of memory");
    if (aukit_cuda_audio_pcm_bytes(ctx(L), a, bits, (int)luaL_checkinteger(L, 3), optbool(L, 4, 1), (int)luaL_optinteger(L, 5, 1), buf)) {
        free(buf);
        return fail(L);
    }
    lua_pushlstring(L, buf, total);
    free(buf);
    return 1;
}

/* cu.write(audio, channel, first, {numbers}) -- audio.data[c][i] = v */
static int l_write(lua_State *L) {
    aukit_audio *a = check_audio(L, 1);
    const int c = (int)luaL_checkinteger(L, 2) - 1;
    const size_t first = (size_t)luaL_checknumber(L, 3) - 1, count = lua_rawlen(L, 4);
    float *buf = (float *)malloc(sizeof(float) * (count ? count : 1));
    if (!buf) return luaL_error(L, "out of memory");
    for (size_t i = 0; i < count; i++) { lua_rawgeti(L, 4, (int)i + 1); buf[i] = (float)lua_tonumberx(L, -1, NULL); lua_settop(L, -2); }
    const int rc = aukit_cuda_audio_upload(ctx(L), a, c, first, count, buf);
    free(buf);
    if (rc) return fail(L);
    return 0;
}

/* cu.device_count() -> number of GPUs cu.preload spreads a buffer over */
static aukit_group *g_group; /* created by the first cu.preload on a box with more than one GPU */
static int g_ndev = -1;

static int device_count(lua_State *L) {
    if (g_ndev < 0) {
        ctx(L);
        g_ndev = 1;
        aukit_group *g = NULL;
        if (aukit_cuda_group_create(NULL, 0, &g) == 0) {
            g_ndev = aukit_cuda_group_size(g);
            if (g_ndev > 1) g_group = g; else aukit_cuda_group_destroy(g);
        }
    }
    return g_ndev;
}

static int l_device_count(lua_State *L) { lua_pushinteger(L, device_count(L)); return 1; }

/* cu.preload(data, bitDepth, dataType, channels, sampleRate, targetRate, interpolation, mono, peakAmplitude, bigEndian)
 * -> audio: auplay.lua:12-27 (aukit.pcm -> :resample -> :mono -> effects.normalize) as the two fused passes on the host
 * string; time-sharded over every visible GPU when there are several (same bits as one GPU). */
static int l_preload(lua_State *L) {
    size_t n;
    const char *d = luaL_checklstring(L, 1, &n);
    aukit_pipeline_desc p;
    memset(&p, 0, sizeof p);
    p.bitDepth = (int)luaL_optinteger(L, 2, 16);
    p.dataType = (int)luaL_optinteger(L, 3, 0);
    p.channels = (int)luaL_optinteger(L, 4, 2);
    p.srcRate = luaL_optnumber(L, 5, 44100);
    p.dstRate = luaL_optnumber(L, 6, 48000);
    p.interpolation = (int)luaL_optinteger(L, 7, 1);
    p.mono = optbool(L, 8, 1);
    const double peak = luaL_optnumber(L, 9, 1.0);
    p.bigEndian = optbool(L, 10, 0);
    if (p.bitDepth != 8 && p.bitDepth != 16 && p.bitDepth != 24 && p.bitDepth != 32) return luaL_error(L, "bad argument #2 (invalid bit depth)");
    if (p.dataType < 0 || p.dataType > 2) return luaL_error(L, "bad argument #3 (invalid data type)");
    if (p.dataType == 2 && p.bitDepth != 32) return luaL_error(L, "bad argument #2 (float audio must have 32-bit depth)");
    if (p.channels < 1) return luaL_error(L, "number outside of range (expected %d to be at least 1)", p.channels);
    const size_t fb = (size_t)p.channels * (size_t)(p.bitDepth / 8);
    if (n % fb) return luaL_error(L, "bad argument #1 (uneven amount of data per channel)");
    p.n_in_total = n / fb;
    p.in_first = 0; p.in_avail = (size_t)p.n_in_total;
    p.out_first = 0; p.n_out = (size_t)aukit_resample_out_len(p.n_in_total, p.srcRate, p.dstRate);
    aukit_audio *a = NULL;
    aukit_ctx *c = ctx(L);
    const int rc = device_count(L) > 1 ? aukit_cuda_group_preload_audio(g_group, c, &p, d, n, peak, &a)
                                       : aukit_cuda_preload_audio(c, &p, d, n, peak, &a);
    if (rc) return fail(L);
    return push_audio(L, a);
}

static int l_gc(lua_State *L) {
    audio_ud *u = (audio_ud *)luaL_checkudata(L, 1, AUDIO_MT);
    if (u->a) { aukit_cuda_audio_free(g_ctx, u->a); u->a = NULL; }
    return 0;
}

static const luaL_Reg funcs[] = {
    {"pcm", l_pcm}, {"g711", l_g711}, {"adpcm", l_adpcm}, {"ima_adpcm_wav", l_ima_wav}, {"msadpcm", l_msadpcm},
    {"wav", l_wav}, {"new", l_new}, {"resample", l_resample}, {"mono", l_mono}, {"concat", l_concat},
    {"au", l_au}, {"aiff", l_aiff}, {"amplify", l_amplify}, {"invert", l_invert}, {"fade", l_fade}, {"delay", l_delay}, {"center", l_center}, {"lowpass", l_lowpass}, {"highpass", l_highpass}, {"pcm_out", l_pcm_out}, {"pcm_bytes", l_pcm_bytes}, {"normalize", l_normalize}, {"channels", l_channels}, {"sample_rate", l_sample_rate},
    {"frames", l_frames}, {"read", l_read}, {"write", l_write}, {"set_sample_rate", l_set_sample_rate},
    {"stream_chunk", l_stream_chunk}, {"preload", l_preload}, {"device_count", l_device_count}, {NULL, NULL}};

int luaopen_aukit_cuda(lua_State *L) {
    luaL_newmetatable(L, AUDIO_MT);
    lua_pushvalue(L, -1);
    lua_setfield(L, -2, "__index");
    {
        static const luaL_Reg mt[] = {{"__gc", l_gc}, {NULL, NULL}};
        luaL_setfuncs(L, mt, 0);
    }
    lua_settop(L, -2);
    lua_createtable(L, 0, 20);
    luaL_setfuncs(L, funcs, 0);
    lua_pushinteger(L, AUKIT_CUDA_ABI_VERSION);
    lua_setfield(L, -2, "abi_version");
    return 1;
}
