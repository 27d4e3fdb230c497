// containers.cu -- aukit.au (A:1634-1647) and aukit.aiff (A:1580-1631): host-side header walks that end in
// the K1 / K2 kernels (aukit.pcm with bigEndian = true, aukit.g711).  SURVEY 8(f) rank 3.
// The walks keep the reference's behaviour where it differs from the file formats' specifications:
//   * AU: the header's data offset is 0-based but the reference hands it to str_sub as a 1-based index, so
//     the payload starts one byte early (and, with a size field, ends one byte early);
//   * AIFF: the FORM size is skipped, a COMM chunk advances by its parsed fields rather than by its size
//     field, text chunks are not padded to even lengths, the 80-bit sample rate keeps 56 mantissa bits in a
//     double, and the payload length comes from COMM (frames * channels * floor(bitDepth / 8)).
#include "common.cuh"

#include <math.h>
#include <string.h>

namespace {

struct reader {
    const uint8_t *p;
    size_t n;
    uint32_t be32(size_t o) const { return ((uint32_t)p[o] << 24) | ((uint32_t)p[o + 1] << 16) | ((uint32_t)p[o + 2] << 8) | p[o + 3]; }
    int be16s(size_t o) const { return (int)(int16_t)(((unsigned)p[o] << 8) | p[o + 1]); }
    bool tag(size_t o, const char *t) const { return memcmp(p + o, t, 4) == 0; }
};

// Lua's string.sub(s, i, j) for i >= 0 and signed j: selected bytes as a 0-based (offset, length)
void sub_range(size_t n, long long i, long long j, size_t *off, size_t *len) {
    if (i < 1) i = 1;
    if (j < 0) j += (long long)n + 1;
    if (j > (long long)n) j = (long long)n;
    *off = i > j ? 0 : (size_t)(i - 1);
    *len = i > j ? 0 : (size_t)(j - i + 1);
}

int short_data() { return aukit_fail("bad argument #2 to 'unpack' (data string too short)"); }   // string.unpack's own message

}  // namespace

extern "C" int aukit_cuda_au_parse(const void *h_data, size_t nbytes, aukit_container_info *info) {
    if (!h_data || !info) return aukit_fail("aukit_cuda: null argument");
    memset(info, 0, sizeof *info);
    const reader r{static_cast<const uint8_t *>(h_data), nbytes};
    if (nbytes < 24) return short_data();
    if (!r.tag(0, ".snd")) return aukit_fail("invalid AU file");
    const long long offset = r.be32(4), size = r.be32(8);
    const uint32_t enc = r.be32(12);
    info->sampleRate = (double)r.be32(16);
    info->channels = (int)r.be32(20);
    sub_range(nbytes, offset, size != 0xFFFFFFFFll ? offset + size - 1 : -1, &info->data_off, &info->data_len);
    info->bigEndian = 1;
    info->dataType = AUKIT_SIGNED;
    switch (enc) {
    case 1: info->codec = AUKIT_CODEC_G711; info->ulaw = 1; return 0;
    case 27: info->codec = AUKIT_CODEC_G711; info->ulaw = 0; return 0;
    case 2: info->bitDepth = 8; return 0;
    case 3: info->bitDepth = 16; return 0;
    case 4: info->bitDepth = 24; return 0;
    case 5: info->bitDepth = 32; return 0;
    case 6: info->bitDepth = 32; info->dataType = AUKIT_FLOAT; return 0;
    }
    return aukit_fail("unsupported encoding type %u", (unsigned)enc);
}

extern "C" int aukit_cuda_aiff_parse(const void *h_data, size_t nbytes, aukit_container_info *info) {
    if (!h_data || !info) return aukit_fail("aukit_cuda: null argument");
    memset(info, 0, sizeof *info);
    const reader r{static_cast<const uint8_t *>(h_data), nbytes};
    if (nbytes < 4) return short_data();
    if (!r.tag(0, "FORM")) return aukit_fail("bad argument #1 (not an AIFF file)");
    if (nbytes < 12) return short_data();
    const bool aifc = r.tag(8, "AIFC");
    if (!aifc && !r.tag(8, "AIFF")) return aukit_fail("bad argument #1 (not an AIFF file)");
    size_t pos = 12;
    bool comm = false;
    char comp[5] = "";
    double payload_len = 0.0;
    static const struct { const char *tag, *key; } texts[] = {{"NAME", "title"}, {"AUTH", "artist"}, {"(c) ", "copyright"}, {"ANNO", "comment"}};
    while (pos < nbytes) {
        if (pos + 8 > nbytes) return short_data();
        const size_t id = pos;
        const uint32_t size = r.be32(pos + 4);
        pos += 8;
        if (r.tag(id, "COMM")) {
            if (pos + 18 > nbytes) return short_data();
            info->channels = r.be16s(pos);
            const double frames = (double)r.be32(pos + 2);
            info->bitDepth = r.be16s(pos + 6);
            const int e = ((int)r.p[pos + 8] << 8) | r.p[pos + 9];
            unsigned long long m = 0;
            for (int k = 0; k < 7; k++) m = (m << 8) | r.p[pos + 10 + k];
            pos += 18;
            if (aifc) {
                if (pos + 5 > nbytes) return short_data();
                memcpy(comp, r.p + pos, 4);
                const size_t sl = r.p[pos + 4];
                if (pos + 5 + sl > nbytes) return short_data();
                pos += 5 + sl + (sl % 2 == 0 ? 1 : 0);
            }
            payload_len = frames * (double)info->channels * floor((double)info->bitDepth / 8.0);
            int ex = ((e & 0x7FFF) - 0x3FFE) % 0x800;
            if (ex < 0) ex += 0x800;                                    // Lua's % is floored
            info->sampleRate = ldexp(((e & 0x8000) ? -1.0 : 1.0) * (double)m / 72057594037927936.0, ex);
            comm = true;
        } else if (r.tag(id, "SSND")) {
            if (pos + 8 > nbytes) return short_data();
            const double first = (double)(pos + 8 + 1) + (double)r.be32(pos);                  // 1-based pos + offset
            if (!comm) return aukit_fail("attempt to perform arithmetic on a nil value (local 'length')");
            sub_range(nbytes, (long long)first, (long long)(first + payload_len - 1.0), &info->data_off, &info->data_len);
            info->bigEndian = 1;
            info->dataType = AUKIT_SIGNED;
            if (!aifc || !strcmp(comp, "NONE")) return 0;
            if (!strcmp(comp, "sowt")) { info->bigEndian = 0; return 0; }
            if (!strcmp(comp, "fl32") || !strcmp(comp, "FL32")) { info->bitDepth = 32; info->dataType = AUKIT_FLOAT; return 0; }
            if (!strcmp(comp, "alaw") || !strcmp(comp, "ALAW")) { info->codec = AUKIT_CODEC_G711; info->ulaw = 0; return 0; }
            if (!strcmp(comp, "ulaw") || !strcmp(comp, "ULAW")) { info->codec = AUKIT_CODEC_G711; info->ulaw = 1; return 0; }
            return aukit_fail("Unsupported compression scheme %s", comp);
        } else {
            for (const auto &t : texts)
                if (r.tag(id, t.tag) && info->nmeta < 16) {
                    auto &m = info->meta[info->nmeta++];
                    strcpy(m.key, t.key);
                    sub_range(nbytes, (long long)pos + 1, (long long)pos + (long long)size, &m.off, &m.len);
                }
            pos += size;
        }
    }
    return aukit_fail("invalid AIFF file");
}

static int container_load(aukit_ctx *ctx, const void *h_data, const aukit_container_info *ci, int head_only, aukit_audio **out) {
    if (head_only) {                                                    // aukit.new(0, channels, sampleRate), A:1610
        if (ci->channels < 1) return aukit_fail("number outside of range (expected %d to be at least 1)", ci->channels);
        if (ci->sampleRate < 1) return aukit_fail("number outside of range (expected %g to be at least 1)", ci->sampleRate);
        return aukit_cuda_audio_new(ctx, ci->channels, 0, ci->sampleRate, out);
    }
    const uint8_t *payload = static_cast<const uint8_t *>(h_data) + ci->data_off;
    if (ci->codec == AUKIT_CODEC_G711) return aukit_cuda_g711(ctx, payload, ci->data_len, ci->ulaw, ci->channels, ci->sampleRate, out);
    return aukit_cuda_pcm(ctx, payload, ci->data_len, ci->bitDepth, ci->dataType, ci->channels, ci->sampleRate, 1, ci->bigEndian, out);
}

extern "C" int aukit_cuda_au(aukit_ctx *ctx, const void *h_data, size_t nbytes, aukit_container_info *info_out, aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    aukit_container_info local;
    aukit_container_info *ci = info_out ? info_out : &local;
    if (aukit_cuda_au_parse(h_data, nbytes, ci)) return -1;
    return container_load(ctx, h_data, ci, 0, out);
}

extern "C" int aukit_cuda_aiff(aukit_ctx *ctx, const void *h_data, size_t nbytes, int head_only, aukit_container_info *info_out,
                               aukit_audio **out) {
    if (!ctx || !out) return aukit_fail("aukit_cuda: null argument");
    aukit_container_info local;
    aukit_container_info *ci = info_out ? info_out : &local;
    if (aukit_cuda_aiff_parse(h_data, nbytes, ci)) return -1;
    return container_load(ctx, h_data, ci, head_only, out);
}
