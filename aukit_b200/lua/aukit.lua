--- aukit.lua (B200 facade) -- drop-in replacement for the PRELOAD path of MCJack123/AUKit.
--
-- Same module table, same function names, positional arguments, defaults and error strings as
-- the reference's aukit.lua (cited below as A:line), but every hot loop runs on the GPU through
-- the C module `aukit_cuda` (luaopen_aukit_cuda, csrc/lua_binding.c -> libaukit_cuda.so).
-- auplay.lua's load -> :resample(48000) -> :mono() -> effects.normalize(mono, 0.8) runs unchanged.
--
-- In scope (device-backed): aukit.preload (the fused auplay chain, all GPUs of the box), aukit.pcm, aukit.g711, aukit.adpcm, aukit.msadpcm, aukit.wav, aukit.au, aukit.aiff, aukit.new,
--   Audio:len/channels/resample (none, linear, cubic, sinc)/mono/concat/pcm/stream/wav,
--   aukit.effects.amplify/normalize/lowpass/highpass/invert/fade/delay/center, aukit.defaultInterpolation.
-- Everything else of the reference (players, streams, FLAC/QOA/DFPWM, editing ops, writers) is out of
-- scope of this accelerated path; load the reference module alongside for those.
--
-- An Audio is a table {sampleRate, data, metadata, info}.  `data` is a proxy whose `_h` field is the
-- device handle: #audio.data is the channel count, #audio.data[c] the frame count and
-- audio.data[c][i] reads sample i (block-cached device -> host copies).  `audio.data = other.data`
-- (done by effects.speed, A:3383) rebinds the handle, exactly like sharing the Lua tables would.

local cu = require "aukit_cuda"
local expect = require "cc.expect"

local aukit = {_VERSION = "1.10.0-b200", defaultInterpolation = "linear"}                 -- A:97-99
aukit.effects = {}

local Audio = {}
local Audio_mt

local DATATYPE = {signed = 0, unsigned = 1, float = 2}
local INTERP = {none = 0, linear = 1, cubic = 2, sinc = 3}
aukit.DIALECT_LITERAL, aukit.DIALECT_GENERAL = 0, 1
--- ADPCM block layout: LITERAL reproduces the reference for 1/2 channels (A:1544, A:1331 quirks
--- included); GENERAL is the standard N-channel layout the reference rejects (A:1349, A:1199).
aukit.adpcmDialect = aukit.DIALECT_LITERAL

local wavMetadata = {                                                                      -- A:198-220
    IPRD = "album", INAM = "title", IART = "artist", IWRI = "author", IMUS = "composer", IPRO = "producer",
    IPRT = "trackNumber", ITRK = "trackNumber", IFRM = "trackCount", PRT1 = "partNumber", PRT2 = "partCount",
    TLEN = "length", IRTD = "rating", ICRD = "date", ITCH = "encodedBy", ISFT = "encoder", ISRF = "media",
    IGNR = "genre", ICMT = "comment", ICOP = "copyright", ILNG = "language"
}

local function copy(tab)                                                                   -- A:239-243
    local t = {}
    for k, v in pairs(tab) do t[k] = v end
    return t
end

local function expectAudio(n, var)                                                         -- A:234-237
    if type(var) == "table" and getmetatable(var) == Audio_mt then return var end
    expect(n, var, "Audio") -- always fails
end

-- ---------------------------------------------------------------------------------------------
-- data proxies: lazy device -> host reads, 4096 samples per copy
local BLOCK = 4096
local channel_mt = {
    __len = function(self) return cu.frames(self._h, self._c) end,
    __index = function(self, i)
        if type(i) ~= "number" or i < 1 or i > cu.frames(self._h, self._c) or i % 1 ~= 0 then return nil end
        local b = math.floor((i - 1) / BLOCK)
        local cache = rawget(self, "_cache")
        if not cache or cache.b ~= b then
            local first = b * BLOCK + 1
            local count = math.min(BLOCK, cu.frames(self._h, self._c) - first + 1)
            cache = {b = b, v = cu.read(self._h, self._c, first, count)}
            rawset(self, "_cache", cache)
        end
        return cache.v[i - b * BLOCK]
    end,
    __newindex = function(self, i, v)
        cu.write(self._h, self._c, i, {v})
        rawset(self, "_cache", nil)
    end
}
local data_mt = {
    __len = function(self) return cu.channels(self._h) end,
    __index = function(self, c)
        if type(c) ~= "number" or c < 1 or c > cu.channels(self._h) then return nil end
        local ch = setmetatable({_h = self._h, _c = c}, channel_mt)
        rawset(self, c, ch)                      -- kept: a loop over audio.data[c][i] then reads one 4096-sample block per 4096 samples
        return ch
    end
}

local function wrap(handle, metadata, info)
    return setmetatable({sampleRate = cu.sample_rate(handle), data = setmetatable({_h = handle}, data_mt),
                         metadata = metadata or {}, info = info or {}}, Audio_mt)
end

-- `sampleRate` is a plain writable field, as in the reference (effects.speed assigns it, A:3383), and `data` may have
-- been rebound to another Audio's: every device call goes through handle(), which first tells the C side the rate
-- this Audio table currently carries.
local function handle(audio)
    local h = audio.data._h
    if type(audio.sampleRate) == "number" then cu.set_sample_rate(h, audio.sampleRate) end
    return h
end
-- after an in-place device operation: the cached host blocks of every channel proxy are stale
local function invalidate(audio)
    local d = audio.data
    for c = 1, cu.channels(d._h) do
        local ch = rawget(d, c)
        if ch then rawset(ch, "_cache", nil) end
    end
end

-- ---------------------------------------------------------------------------------------------
-- Audio methods

--- Returns the length of the audio object in seconds. (A:638)
function Audio:len() return #self.data[1] / self.sampleRate end

--- Returns the number of channels in the audio object. (A:644)
function Audio:channels() return #self.data end

--- Creates a new audio object with the data resampled to a different sample rate. (A:653)
function Audio:resample(sampleRate, interpolation)
    expect(1, sampleRate, "number")
    interpolation = expect(2, interpolation, "string", "nil") or aukit.defaultInterpolation
    if not INTERP[interpolation] then error("bad argument #2 (invalid interpolation type)", 2) end
    local out = wrap(cu.resample(handle(self), sampleRate, INTERP[interpolation]), copy(self.metadata), copy(self.info))
    out.sampleRate = sampleRate
    return out
end

--- Mixes down all channels to a new mono-channel audio object. (A:677)
function Audio:mono()
    local out = wrap(cu.mono(handle(self)), copy(self.metadata), copy(self.info))
    out.sampleRate = self.sampleRate
    return out
end

--- Converts the audio data to raw PCM samples: un-rounded numbers, like the reference. (A:901)
function Audio:pcm(bitDepth, dataType, interleaved)
    bitDepth = expect(1, bitDepth, "number", "nil") or 8
    dataType = expect(2, dataType, "string", "nil") or "signed"
    expect(3, interleaved, "boolean", "nil")
    if interleaved == nil then interleaved = true end
    if bitDepth ~= 8 and bitDepth ~= 16 and bitDepth ~= 24 and bitDepth ~= 32 then error("bad argument #2 (invalid bit depth)", 2) end
    if dataType ~= "signed" and dataType ~= "unsigned" and dataType ~= "float" then error("bad argument #3 (invalid data type)", 2) end
    if dataType == "float" and bitDepth ~= 32 then error("bad argument #2 (float audio must have 32-bit depth)", 2) end
    return cu.pcm_out(handle(self), bitDepth, DATATYPE[dataType], interleaved)
end

--- Returns a function that can be called to encode PCM samples in chunks, and the total length in seconds. (A:921)
function Audio:stream(chunkSize, bitDepth, dataType)
    chunkSize = expect(1, chunkSize, "number", "nil") or 131072
    bitDepth = expect(2, bitDepth, "number", "nil") or 8
    dataType = expect(3, dataType, "string", "nil") or "signed"
    if bitDepth ~= 8 and bitDepth ~= 16 and bitDepth ~= 24 and bitDepth ~= 32 then error("bad argument #2 (invalid bit depth)", 2) end
    if dataType ~= "signed" and dataType ~= "unsigned" and dataType ~= "float" then error("bad argument #3 (invalid data type)", 2) end
    if dataType == "float" and bitDepth ~= 32 then error("bad argument #2 (float audio must have 32-bit depth)", 2) end
    local pos, done = 1, false
    return function()
        if done then return nil end
        local p = pos / self.sampleRate
        local v = cu.stream_chunk(handle(self), bitDepth, DATATYPE[dataType], pos, chunkSize)   -- nil past the end (A:878)
        if v == nil then done = true return nil end
        pos = pos + chunkSize
        return v, p
    end, #self.data[1] / self.sampleRate
end

--- Converts the audio data to a WAV file (PCM depths; DFPWM is outside this module). (A:942)
-- aukit.packRounding: how the host's string.pack narrows the un-rounded sample values of Audio:pcm
-- ("floor", "truncate" or "nearest"); the reference leaves this to the Lua it runs on.
function Audio:wav(bitDepth)
    bitDepth = expect(1, bitDepth, "number", "nil") or 16
    if bitDepth == 1 then error("aukit_cuda: DFPWM output is outside the accelerated path", 2)
    elseif bitDepth ~= 8 and bitDepth ~= 16 and bitDepth ~= 24 and bitDepth ~= 32 then error("bad argument #2 (invalid bit depth)", 2) end
    local rounding = ({truncate = 0, floor = 1, nearest = 2})[aukit.packRounding or "floor"] or 1
    local str = cu.pcm_bytes(handle(self), bitDepth, bitDepth == 8 and 1 or 0, true, rounding)
    local nc = #self.data
    if self.metadata and next(self.metadata) then
        local info = {}
        for k, v in pairs(self.metadata) do
            for l, w in pairs(wavMetadata) do
                if w == k then
                    info[#info+1] = l
                    info[#info+1] = tostring(v)
                    break
                end
            end
        end
        local list = string.pack("!2<c4" .. ("c4s4Xh"):rep(#info / 2), "INFO", table.unpack(info))
        return string.pack("<c4Ic4c4IHHIIHHc4s4c4I", "RIFF", #str + 36, "WAVE", "fmt ", 16, 1, nc, self.sampleRate,
            self.sampleRate * nc * bitDepth / 8, nc * bitDepth / 8, bitDepth, "LIST", list, "data", #str) .. str
    end
    return string.pack("<c4Ic4c4IHHIIHHc4I", "RIFF", #str + 36, "WAVE", "fmt ", 16, 1, nc, self.sampleRate,
        self.sampleRate * nc * bitDepth / 8, nc * bitDepth / 8, bitDepth, "data", #str) .. str
end

--- Concatenates this audio object with others, resampling where rates differ. (A:696)
function Audio:concat(...)
    local audios = {self, ...}
    local hs = {handle(self)}
    for i = 2, #audios do
        expectAudio(i - 1, audios[i])
        if audios[i].sampleRate ~= self.sampleRate then audios[i] = audios[i]:resample(self.sampleRate) end
        hs[i] = handle(audios[i])
    end
    local out = wrap(cu.concat(table.unpack(hs)), copy(self.metadata), copy(self.info))
    out.sampleRate = self.sampleRate
    return out
end

Audio_mt = {__index = Audio, __concat = Audio.concat, __len = Audio.len,
            __tostring = function(self) return "Audio: " .. #self.data .. " channels, " .. self:len() .. " seconds (device)" end}

-- ---------------------------------------------------------------------------------------------
-- loaders

--- Creates a new audio object from the specified raw PCM data. (A:1049)
function aukit.pcm(data, bitDepth, dataType, channels, sampleRate, interleaved, bigEndian)
    expect(1, data, "string", "table")
    if type(data) == "table" then error("table PCM input is host-side only on the accelerated path; pass the packed string", 2) end
    bitDepth = expect(2, bitDepth, "number", "nil") or 8
    dataType = expect(3, dataType, "string", "nil") or "signed"
    channels = expect(4, channels, "number", "nil") or 1
    sampleRate = expect(5, sampleRate, "number", "nil") or 48000
    expect(6, interleaved, "boolean", "nil")
    if interleaved == nil then interleaved = true end
    expect(7, bigEndian, "boolean", "nil")
    if bitDepth ~= 8 and bitDepth ~= 16 and bitDepth ~= 24 and bitDepth ~= 32 then error("bad argument #2 (invalid bit depth)", 2) end
    if dataType ~= "signed" and dataType ~= "unsigned" and dataType ~= "float" then error("bad argument #3 (invalid data type)", 2) end
    if dataType == "float" and bitDepth ~= 32 then error("bad argument #2 (float audio must have 32-bit depth)", 2) end
    expect.range(channels, 1)
    expect.range(sampleRate, 1)
    if (#data / (bitDepth / 8)) % channels ~= 0 then error("bad argument #1 (uneven amount of data per channel)", 2) end
    local out = wrap(cu.pcm(data, bitDepth, DATATYPE[dataType], channels, sampleRate, interleaved, bigEndian or false),
                     {}, {bitDepth = bitDepth, dataType = dataType})
    out.sampleRate = sampleRate
    return out
end

--- Creates a new audio object from IMA ADPCM data (nibble string). (A:1183)
function aukit.adpcm(data, channels, sampleRate, topFirst, interleaved, predictor, step_index)
    expect(1, data, "string", "table")
    if type(data) == "table" then error("table ADPCM input is host-side only on the accelerated path", 2) end
    channels = expect(2, channels, "number", "nil") or 1
    sampleRate = expect(3, sampleRate, "number", "nil") or 48000
    expect(4, topFirst, "boolean", "nil")
    if topFirst == nil then topFirst = true end
    expect(5, interleaved, "boolean", "nil")
    if interleaved == nil then interleaved = true end
    predictor = expect(6, predictor, "number", "table", "nil")
    step_index = expect(7, step_index, "number", "table", "nil")
    expect.range(channels, 1)
    expect.range(sampleRate, 1)
    local function state(v, argn, lo, hi)
        if v == nil then return nil end
        if type(v) == "number" then
            if channels ~= 1 then error("bad argument #" .. argn .. " (table too short)", 3) end
            return {expect.range(v, lo, hi)}
        end
        if channels > #v then error("bad argument #" .. argn .. " (table too short)", 3) end
        for i = 1, channels do expect.range(v[i], lo, hi) end
        return v
    end
    predictor, step_index = state(predictor, 6, -32768, 32767), state(step_index, 7, 0, 88)
    local out = wrap(cu.adpcm(data, channels, sampleRate, topFirst, interleaved, predictor, step_index),
                     {}, {bitDepth = 16, dataType = "signed"})
    out.sampleRate = sampleRate
    return out
end

--- NOT in the reference: auplay.lua:12-27 -- aukit.pcm(data, ...) -> :resample(targetRate, interpolation) -> :mono()
--- -> effects.normalize(peakAmplitude) -- as ONE call on the packed bytes (two fused GPU passes, no intermediate Audio;
--- on a box with several GPUs the buffer is time-sharded over all of them).  Same result, within fp32 rounding, as the
--- four reference calls; arguments and defaults are theirs (mono = true and peakAmplitude = 0.8 as auplay uses them).
function aukit.preload(data, bitDepth, dataType, channels, sampleRate, targetRate, interpolation, mono, peakAmplitude, bigEndian)
    expect(1, data, "string")
    bitDepth = expect(2, bitDepth, "number", "nil") or 8
    dataType = expect(3, dataType, "string", "nil") or "signed"
    channels = expect(4, channels, "number", "nil") or 1
    sampleRate = expect(5, sampleRate, "number", "nil") or 48000
    targetRate = expect(6, targetRate, "number", "nil") or 48000
    interpolation = expect(7, interpolation, "string", "nil") or aukit.defaultInterpolation
    expect(8, mono, "boolean", "nil")
    if mono == nil then mono = true end
    peakAmplitude = expect(9, peakAmplitude, "number", "nil") or 0.8
    expect(10, bigEndian, "boolean", "nil")
    if bitDepth ~= 8 and bitDepth ~= 16 and bitDepth ~= 24 and bitDepth ~= 32 then error("bad argument #2 (invalid bit depth)", 2) end
    if not DATATYPE[dataType] then error("bad argument #3 (invalid data type)", 2) end
    if dataType == "float" and bitDepth ~= 32 then error("bad argument #2 (float audio must have 32-bit depth)", 2) end
    if not INTERP[interpolation] then error("bad argument #7 (invalid interpolation type)", 2) end
    expect.range(channels, 1)
    expect.range(sampleRate, 1)
    expect.range(targetRate, 1)
    if #data % (channels * bitDepth / 8) ~= 0 then error("bad argument #1 (uneven amount of data per channel)", 2) end
    local out = wrap(cu.preload(data, bitDepth, DATATYPE[dataType], channels, sampleRate, targetRate, INTERP[interpolation], mono,
                                peakAmplitude, bigEndian or false), {}, {bitDepth = bitDepth, dataType = dataType})
    out.sampleRate = targetRate
    return out
end

--- Number of GPUs aukit.preload spreads a buffer over.
function aukit.deviceCount() return cu.device_count() end

--- Creates a new audio object from Microsoft ADPCM data. (A:1283)
function aukit.msadpcm(data, blockAlign, channels, sampleRate, coefficients)
    expect(1, data, "string")
    expect(2, blockAlign, "number")
    channels = expect(3, channels, "number", "nil") or 1
    sampleRate = expect(4, sampleRate, "number", "nil") or 48000
    expect(5, coefficients, "table", "nil")
    expect.range(sampleRate, 1)
    local c1, c2
    if coefficients then
        if type(coefficients[1]) ~= "table" then error("bad argument #5 (first entry is not a table)", 2) end
        if type(coefficients[2]) ~= "table" then error("bad argument #5 (second entry is not a table)", 2) end
        if #coefficients[1] ~= #coefficients[2] then error("bad argument #5 (lists are not the same length)", 2) end
        c1, c2 = coefficients[1], coefficients[2]
    end
    local out = wrap(cu.msadpcm(data, blockAlign, channels, sampleRate, c1, c2, aukit.adpcmDialect),
                     {}, {bitDepth = 16, dataType = "signed"})
    out.sampleRate = sampleRate
    return out
end

--- Creates a new audio object from G.711 u-law/A-law data. (A:1361)
function aukit.g711(data, ulaw, channels, sampleRate)
    expect(1, data, "string")
    expect(2, ulaw, "boolean")
    channels = expect(3, channels, "number", "nil") or 1
    sampleRate = expect(4, sampleRate, "number", "nil") or 8000
    -- A:1383: bitDepth/dataType land in `metadata`, `info` stays empty
    local out = wrap(cu.g711(data, ulaw, channels, sampleRate), {bitDepth = ulaw and 14 or 13, dataType = "signed"}, {})
    out.sampleRate = sampleRate
    return out
end

--- Creates a new empty audio object. (A:1784)
function aukit.new(duration, channels, sampleRate)
    expect(1, duration, "number")
    channels = expect(2, channels, "number", "nil") or 1
    sampleRate = expect(3, sampleRate, "number", "nil") or 48000
    expect.range(channels, 1)
    expect.range(sampleRate, 1)
    local out = wrap(cu.new(channels, math.floor(duration * sampleRate), sampleRate), {}, {})
    out.sampleRate = sampleRate
    return out
end

--- Creates a new audio object from a WAV file: RIFF walk on the host, decode on the device. (A:1456)
function aukit.wav(data, head)
    expect(1, data, "string")
    local h, info = cu.wav(data, head and true or false, aukit.adpcmDialect)
    local meta = {}
    for _, tag in ipairs(info.tags) do                                                     -- A:1559-1568
        local key = wavMetadata[tag[1]]
        if key then meta[key] = tonumber(tag[2]) or tag[2] end
    end
    local out = wrap(h, meta, {dataType = info.dataType, bitDepth = info.bitDepth})         -- A:1553-1554
    return out
end

local function container(h, info, aiffMeta)
    local metadata, inf
    if info.codec == "g711" then metadata, inf = {bitDepth = info.ulaw and 14 or 13, dataType = "signed"}, {}    -- A:1383
    else metadata, inf = {}, {bitDepth = info.bitDepth, dataType = info.dataType} end                          -- A:1171
    if aiffMeta then
        metadata = {}
        for _, kv in ipairs(info.meta) do metadata[kv[1]] = kv[2] end                                           -- A:1617-1630
    end
    return wrap(h, metadata, inf)
end

--- Creates a new audio object from an AU file. (A:1634)
function aukit.au(data)
    expect(1, data, "string")
    local h, info = cu.au(data)
    return container(h, info, false)
end

--- Creates a new audio object from an AIFF or AIFC file. (A:1580)
function aukit.aiff(data, head)
    expect(1, data, "string")
    local h, info = cu.aiff(data, head and true or false)
    local out = container(h, info, true)
    if head then out.info = {} end
    return out
end

-- ---------------------------------------------------------------------------------------------
-- effects: mutate the argument and return it (A:3368, A:3458; auplay.lua:27 relies on it)

--- Amplifies the audio by the multiplier specified. (A:3356)
function aukit.effects.amplify(audio, multiplier)
    expectAudio(1, audio)
    expect(2, multiplier, "number")
    if multiplier == 1 then return audio end
    cu.amplify(handle(audio), multiplier)
    invalidate(audio)
    return audio
end

--- Inverts all channels in the specified audio. (A:3412)
function aukit.effects.invert(audio)
    expectAudio(1, audio)
    cu.invert(handle(audio))
    invalidate(audio)
    return audio
end

--- Fades a period of music from one amplitude to another. (A:3392)
function aukit.effects.fade(audio, startTime, startAmplitude, endTime, endAmplitude)
    expectAudio(1, audio)
    expect(2, startTime, "number")
    expect(3, startAmplitude, "number")
    expect(4, endTime, "number")
    expect(5, endAmplitude, "number")
    if startAmplitude == 1 and endAmplitude == 1 then return audio end
    cu.fade(handle(audio), startTime, startAmplitude, endTime, endAmplitude)
    invalidate(audio)
    return audio
end

--- Adds a delay to the specified audio. (A:3500)
function aukit.effects.delay(audio, delay, multiplier)
    expectAudio(1, audio)
    expect(2, delay, "number")
    multiplier = expect(3, multiplier, "number", "nil") or 0.5
    cu.delay(handle(audio), delay, multiplier)
    invalidate(audio)
    return audio
end

--- Centers the DC offset of each channel. (A:3465)
function aukit.effects.center(audio)
    expectAudio(1, audio)
    cu.center(handle(audio))
    invalidate(audio)
    return audio
end

--- Applies a low-pass filter to the specified audio. (A:3586; auplay.lua:30)
function aukit.effects.lowpass(audio, frequency)
    expectAudio(1, audio)
    expect(2, frequency, "number")
    cu.lowpass(handle(audio), frequency)
    invalidate(audio)
    return audio
end

--- Applies a high-pass filter to the specified audio. (A:3605)
function aukit.effects.highpass(audio, frequency)
    expectAudio(1, audio)
    expect(2, frequency, "number")
    cu.highpass(handle(audio), frequency)
    invalidate(audio)
    return audio
end

--- Normalizes audio to the specified peak amplitude. (A:3431)
function aukit.effects.normalize(audio, peakAmplitude, independent)
    expectAudio(1, audio)
    peakAmplitude = expect(2, peakAmplitude, "number", "nil") or 1
    expect(3, independent, "boolean", "nil")
    cu.normalize(handle(audio), peakAmplitude, independent or false)
    invalidate(audio)
    return audio
end

return aukit
