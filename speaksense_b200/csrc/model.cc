// model.cc - ggml legacy .bin parser + arena packer (see model.h).
#include "model.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <map>

namespace ss {

static const char *g_lang[kNumLangTable] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"};

int lang_id(const char *code) {
    for (int i = 0; i < kNumLangTable; i++)
        if (!strcmp(code, g_lang[i])) return i;
    return -1;
}

namespace {

constexpr uint32_t kGgmlMagic = 0x67676d6c;
constexpr char kArenaMagic[8] = {'S', 'S', 'A', 'R', 'E', 'N', 'A', '1'};

struct FileTensor {
    int n_dims = 0;
    int ne[4] = {1, 1, 1, 1};
    int ttype = 0;   // GGML_TYPE: 0 f32, 1 f16, 2 q4_0, 3 q4_1, 6 q5_0, 7 q5_1, 8 q8_0 (blocks of 32, dequantised at load)
    const unsigned char *data = nullptr;
    size_t count() const { return (size_t)ne[0] * ne[1] * ne[2] * ne[3]; }
};

struct ParsedFile {
    std::vector<unsigned char> buf;
    HParams hp{};
    size_t meta_len = 0;             // bytes up to the first tensor
    std::map<std::string, FileTensor> tensors;
};

size_t parse_meta(const unsigned char *b, size_t sz, HParams &hp, const float **filters, std::vector<std::string> *toks) {
    size_t o = 0;
    auto need = [&](size_t n) { if (o + n > sz) SS_THROW(-2, "model file truncated at byte %zu", o); };
    need(4 + 44 + 8);
    uint32_t magic; memcpy(&magic, b, 4); o = 4;
    if (magic != kGgmlMagic) SS_THROW(-2, "bad ggml magic 0x%08x", magic);
    memcpy(&hp, b + o, 44); o += 44;
    int n_mel, n_fft; memcpy(&n_mel, b + o, 4); memcpy(&n_fft, b + o + 4, 4); o += 8;
    if (n_mel != hp.n_mels || n_fft != kNBins) SS_THROW(-2, "unexpected mel filterbank %dx%d", n_mel, n_fft);
    // (every count is validated BEFORE it is used as a divisor or a size: a malformed file must come back as an error)
    if (hp.n_audio_head <= 0 || hp.n_text_head <= 0 || hp.n_audio_state <= 0 || hp.n_text_state <= 0 || hp.n_audio_layer <= 0 ||
        hp.n_text_layer <= 0 || hp.n_audio_ctx <= 0 || hp.n_text_ctx <= 0 || hp.n_mels <= 0 || hp.n_audio_layer > 256 || hp.n_text_layer > 256 ||
        hp.n_audio_state > 16384 || hp.n_text_state > 16384 || hp.n_audio_ctx > 65536 || hp.n_text_ctx > 65536 || hp.n_mels > 1024)
        SS_THROW(-2, "model header holds a non-positive / absurd hyper-parameter");
    if (hp.n_vocab < 50000 || hp.n_vocab > 100000 || hp.n_audio_state % 128 || hp.n_text_state % 128 ||
        hp.n_audio_state / hp.n_audio_head != 64 || hp.n_text_state / hp.n_text_head != 64)
        SS_THROW(-2, "unsupported hyper-parameters (need d %% 128 == 0, head dim 64)");
    need((size_t)n_mel * n_fft * 4);
    if (filters) *filters = reinterpret_cast<const float *>(b + o);
    o += (size_t)n_mel * n_fft * 4;
    need(4);
    int n_tok; memcpy(&n_tok, b + o, 4); o += 4;
    if (n_tok < 0 || n_tok > hp.n_vocab) SS_THROW(-2, "bad vocab size %d", n_tok);
    for (int i = 0; i < n_tok; i++) {
        uint32_t len; need(4); memcpy(&len, b + o, 4); o += 4; need(len);
        if (toks) toks->emplace_back(reinterpret_cast<const char *>(b + o), len);
        o += len;
    }
    return o;
}

// ---- ggml block-quantised types (QK = 32): the files whisper.cpp's quantize tool writes and the reference's
// script/download-ggml-model.sh:28-51 fetches (e.g. large-v3-q5_0).  Weights are dequantised ONCE at load into the f16
// arena (ggml dequantize_row_*: f16 scale [, f16 min], 4 / 5 / 8-bit codes), then the f16 kernels run unchanged.
// whisper.cpp itself multiplies such weights against q8_0-quantised activations; that integer path is not restated.
inline size_t quant_block_bytes(int ttype) { return ttype == 2 ? 18 : ttype == 3 ? 20 : ttype == 6 ? 22 : ttype == 7 ? 24 : ttype == 8 ? 34 : 0; }
inline float f16_bits_to_f32_host(uint16_t u);
__attribute__((optimize("fp-contract=off")))      // x * d + m in two roundings, as ggml's scalar dequantize_row_*
void dequantize_block(const unsigned char *b, int ttype, float *y) {
    uint16_t dh; memcpy(&dh, b, 2);
    const float d = f16_bits_to_f32_host(dh);
    if (ttype == 8) { const int8_t *q = reinterpret_cast<const int8_t *>(b + 2); for (int j = 0; j < 32; j++) y[j] = (float)q[j] * d; return; }
    float m = 0.f; size_t o = 2;
    if (ttype == 3 || ttype == 7) { uint16_t mh; memcpy(&mh, b + 2, 2); m = f16_bits_to_f32_host(mh); o = 4; }
    uint32_t qh = 0;
    if (ttype == 6 || ttype == 7) { memcpy(&qh, b + o, 4); o += 4; }
    const unsigned char *qs = b + o;
    for (int j = 0; j < 16; j++) {
        int x0 = qs[j] & 0x0F, x1 = qs[j] >> 4;
        if (ttype == 6 || ttype == 7) { x0 |= (int)((qh >> j) & 1u) << 4; x1 |= (int)((qh >> (j + 16)) & 1u) << 4; }
        if (ttype == 2) { y[j] = (float)(x0 - 8) * d; y[j + 16] = (float)(x1 - 8) * d; }
        else if (ttype == 6) { y[j] = (float)(x0 - 16) * d; y[j + 16] = (float)(x1 - 16) * d; }
        else { y[j] = (float)x0 * d + m; y[j + 16] = (float)x1 * d + m; }
    }
}

void parse_file(const std::string &path, ParsedFile &pf) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) SS_THROW(-2, "cannot open model file '%s'", path.c_str());
    size_t sz = (size_t)f.tellg();
    f.seekg(0);
    pf.buf.resize(sz);
    if (!f.read(reinterpret_cast<char *>(pf.buf.data()), (std::streamsize)sz)) SS_THROW(-2, "short read on '%s'", path.c_str());
    const unsigned char *b = pf.buf.data();
    size_t o = parse_meta(b, sz, pf.hp, nullptr, nullptr);
    pf.meta_len = o;
    while (o < sz) {
        if (o + 12 > sz) SS_THROW(-2, "truncated tensor header");
        int32_t hdr[3]; memcpy(hdr, b + o, 12); o += 12;
        FileTensor t; t.n_dims = hdr[0]; t.ttype = hdr[2];
        int nlen = hdr[1];
        if (t.n_dims < 1 || t.n_dims > 4 || nlen <= 0 || nlen > 256) SS_THROW(-2, "bad tensor header at %zu", o);
        if (t.ttype != 0 && t.ttype != 1 && quant_block_bytes(t.ttype) == 0) SS_THROW(-2, "unsupported ggml tensor type %d (f32, f16, q4_0, q4_1, q5_0, q5_1, q8_0 are)", t.ttype);
        if (o + 4 * (size_t)t.n_dims + (size_t)nlen > sz) SS_THROW(-2, "truncated tensor header");
        if (o + 4 * (size_t)t.n_dims + (size_t)nlen > sz) SS_THROW(-2, "truncated tensor header");
        for (int d = 0; d < t.n_dims; d++) {
            memcpy(&t.ne[d], b + o, 4); o += 4;
            if (t.ne[d] <= 0 || t.ne[d] > (1 << 28)) SS_THROW(-2, "tensor dimension %d out of range at byte %zu", t.ne[d], o);
        }
        std::string name(reinterpret_cast<const char *>(b + o), (size_t)nlen); o += (size_t)nlen;
        if (t.ttype != 0 && t.ttype != 1 && !quant_block_bytes(t.ttype)) SS_THROW(-2, "tensor '%s': unsupported ggml type %d", name.c_str(), t.ttype);
        if ((double)t.ne[0] * t.ne[1] * t.ne[2] * t.ne[3] > 4e9) SS_THROW(-2, "tensor '%s' is absurdly large", name.c_str());
        size_t nb = t.count() * (t.ttype == 1 ? 2 : 4);
        if (quant_block_bytes(t.ttype)) {
            if (t.ne[0] % 32) SS_THROW(-2, "quantised tensor '%s': row length %d is not a multiple of 32", name.c_str(), t.ne[0]);
            nb = t.count() / 32 * quant_block_bytes(t.ttype);
        }
        if (nb > sz - o) SS_THROW(-2, "tensor '%s' truncated", name.c_str());
        t.data = b + o; o += nb;
        pf.tensors[name] = t;
    }
}

inline uint16_t f32_to_f16_bits(float f) { __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; }
inline float f16_bits_to_f32(uint16_t u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }
inline float f16_bits_to_f32_host(uint16_t u) { return f16_bits_to_f32(u); }

// Walks the arena layout.  With base != nullptr and pf != nullptr it also fills the image.
struct Walker {
    unsigned char *base;          // host image (fill mode) or device arena (bind mode)
    const ParsedFile *pf;         // non-null in fill mode
    size_t off;

    size_t take(size_t bytes) { off = (off + 255) & ~(size_t)255; size_t o = off; off += bytes; return o; }
    const FileTensor &get(const std::string &name) const {
        auto it = pf->tensors.find(name);
        if (it == pf->tensors.end()) SS_THROW(-2, "model file lacks tensor '%s'", name.c_str());
        return it->second;
    }
    // copy `n` elements of tensor `name` as f32 into dst
    void put_f32(float *dst, const std::string &name, size_t n) const {
        const FileTensor &t = get(name);
        if (t.count() != n) SS_THROW(-2, "tensor '%s' has %zu elements, expected %zu", name.c_str(), t.count(), n);
        if (t.ttype == 0) memcpy(dst, t.data, n * 4);
        else if (const size_t bb = quant_block_bytes(t.ttype)) { for (size_t i = 0; i < n / 32; i++) dequantize_block(t.data + i * bb, t.ttype, dst + 32 * i); }
        else { const uint16_t *s = reinterpret_cast<const uint16_t *>(t.data); uint16_t v; for (size_t i = 0; i < n; i++) { memcpy(&v, s + i, 2); dst[i] = f16_bits_to_f32(v); } }
    }
    void put_f16(uint16_t *dst, const std::string &name, size_t n) const {
        const FileTensor &t = get(name);
        if (t.count() != n) SS_THROW(-2, "tensor '%s' has %zu elements, expected %zu", name.c_str(), t.count(), n);
        if (t.ttype == 1) memcpy(dst, t.data, n * 2);
        else if (const size_t bb = quant_block_bytes(t.ttype)) {
            float y[32];
            for (size_t i = 0; i < n / 32; i++) { dequantize_block(t.data + i * bb, t.ttype, y); for (int j = 0; j < 32; j++) dst[32 * i + j] = f32_to_f16_bits(y[j]); }
        }
        else { const unsigned char *s = t.data; float v; for (size_t i = 0; i < n; i++) { memcpy(&v, s + 4 * i, 4); dst[i] = f32_to_f16_bits(v); } }
    }
    const float *f32(const std::string &name, size_t n) {
        size_t o = take(n * 4);
        if (pf) put_f32(reinterpret_cast<float *>(base + o), name, n);
        return reinterpret_cast<const float *>(base + o);
    }
    const __half *f16(const std::string &name, size_t n) {
        size_t o = take(n * 2);
        if (pf) put_f16(reinterpret_cast<uint16_t *>(base + o), name, n);
        return reinterpret_cast<const __half *>(base + o);
    }
    LNp ln(const std::string &p, int d) { LNp r; r.w = f32(p + ".weight", d); r.b = f32(p + ".bias", d); return r; }
    Lin lin(const std::string &p, int n_out, int n_in, bool bias = true) {
        Lin r; r.n_out = n_out; r.n_in = n_in;
        r.w = f16(p + ".weight", (size_t)n_out * n_in);
        if (bias) r.b = f32(p + ".bias", n_out);
        return r;
    }
    // rows of several [d_out_i][n_in] matrices stacked; missing bias => zeros
    Lin fused(const std::vector<std::string> &ps, const std::vector<bool> &has_bias, int d_each, int n_in) {
        Lin r; r.n_out = d_each * (int)ps.size(); r.n_in = n_in;
        size_t ow = take((size_t)r.n_out * n_in * 2);
        size_t ob = take((size_t)r.n_out * 4);
        if (pf) {
            for (size_t i = 0; i < ps.size(); i++) {
                put_f16(reinterpret_cast<uint16_t *>(base + ow) + i * (size_t)d_each * n_in, ps[i] + ".weight", (size_t)d_each * n_in);
                float *bd = reinterpret_cast<float *>(base + ob) + i * (size_t)d_each;
                if (has_bias[i]) put_f32(bd, ps[i] + ".bias", d_each); else memset(bd, 0, (size_t)d_each * 4);
            }
        }
        r.w = reinterpret_cast<const __half *>(base + ow);
        r.b = reinterpret_cast<const float *>(base + ob);
        return r;
    }
    // conv1d weight [out][c][3] -> [out][k][c]
    Lin conv(const std::string &p, int n_out, int c) {
        Lin r; r.n_out = n_out; r.n_in = 3 * c;
        size_t ow = take((size_t)n_out * 3 * c * 2);
        if (pf) {
            std::vector<uint16_t> tmp((size_t)n_out * 3 * c);
            put_f16(tmp.data(), p + ".weight", tmp.size());
            uint16_t *d = reinterpret_cast<uint16_t *>(base + ow);
            for (int o = 0; o < n_out; o++)
                for (int ci = 0; ci < c; ci++)
                    for (int k = 0; k < 3; k++) d[((size_t)o * 3 + k) * c + ci] = tmp[((size_t)o * c + ci) * 3 + k];
        }
        r.w = reinterpret_cast<const __half *>(base + ow);
        r.b = f32(p + ".bias", n_out);
        return r;
    }
};

void walk(Walker &w, Model &m) {
    const HParams &hp = m.hp;
    const int d = hp.n_audio_state, dd = hp.n_text_state, T = hp.n_audio_ctx;
    // mel filterbank + per-filter nonzero ranges (from the meta prefix; not a named tensor)
    {
        size_t of = w.take((size_t)hp.n_mels * kNBins * 4);
        size_t orr = w.take((size_t)hp.n_mels * sizeof(int2));
        if (w.pf) {
            const float *fl = nullptr; HParams tmp;
            parse_meta(w.pf->buf.data(), w.pf->buf.size(), tmp, &fl, nullptr);
            memcpy(w.base + of, fl, (size_t)hp.n_mels * kNBins * 4);
            int2 *rg = reinterpret_cast<int2 *>(w.base + orr);
            for (int j = 0; j < hp.n_mels; j++) {
                int lo = kNBins, hi = 0;
                for (int k = 0; k < kNBins; k++) { float v; memcpy(&v, reinterpret_cast<const unsigned char *>(fl) + ((size_t)j * kNBins + k) * 4, 4); if (v != 0.f) { if (k < lo) lo = k; hi = k + 1; } }
                if (lo >= hi) { lo = 0; hi = 0; }
                rg[j] = make_int2(lo, hi);
            }
        }
        m.filters = reinterpret_cast<const float *>(w.base + of);
        m.filt_range = reinterpret_cast<const int2 *>(w.base + orr);
    }
    m.e_pos = w.f32("encoder.positional_embedding", (size_t)T * d);
    m.conv1 = w.conv("encoder.conv1", d, hp.n_mels);
    m.conv2 = w.conv("encoder.conv2", d, d);
    m.enc.resize(hp.n_audio_layer);
    for (int i = 0; i < hp.n_audio_layer; i++) {
        std::string b = "encoder.blocks." + std::to_string(i);
        EncLayer &L = m.enc[i];
        L.attn_ln = w.ln(b + ".attn_ln", d);
        L.qkv = w.fused({b + ".attn.query", b + ".attn.key", b + ".attn.value"}, {true, false, true}, d, d);
        L.o = w.lin(b + ".attn.out", d, d);
        L.mlp_ln = w.ln(b + ".mlp_ln", d);
        L.fc1 = w.lin(b + ".mlp.0", 4 * d, d);
        L.fc2 = w.lin(b + ".mlp.2", d, 4 * d);
    }
    m.ln_post = w.ln("encoder.ln_post", d);
    m.d_pos = w.f32("decoder.positional_embedding", (size_t)hp.n_text_ctx * dd);
    m.tok_emb = w.f16("decoder.token_embedding.weight", (size_t)hp.n_vocab * dd);
    m.dec.resize(hp.n_text_layer);
    for (int i = 0; i < hp.n_text_layer; i++) {
        std::string b = "decoder.blocks." + std::to_string(i);
        DecLayer &L = m.dec[i];
        L.attn_ln = w.ln(b + ".attn_ln", dd);
        L.qkv = w.fused({b + ".attn.query", b + ".attn.key", b + ".attn.value"}, {true, false, true}, dd, dd);
        L.o = w.lin(b + ".attn.out", dd, dd);
        L.cross_ln = w.ln(b + ".cross_attn_ln", dd);
        L.cq = w.lin(b + ".cross_attn.query", dd, dd);
        L.ckv = w.fused({b + ".cross_attn.key", b + ".cross_attn.value"}, {false, true}, dd, d);
        L.co = w.lin(b + ".cross_attn.out", dd, dd);
        L.mlp_ln = w.ln(b + ".mlp_ln", dd);
        L.fc1 = w.lin(b + ".mlp.0", 4 * dd, dd);
        L.fc2 = w.lin(b + ".mlp.2", dd, 4 * dd);
    }
    m.d_ln = w.ln("decoder.ln", dd);
}

void fill_vocab(Model &m, std::vector<std::string> &&toks) {
    Vocab &v = m.vocab;
    const HParams &hp = m.hp;
    v.eot = 50256; v.sot = 50257; v.translate = 50357; v.transcribe = 50358; v.solm = 50359;
    v.prev = 50360; v.nosp = 50361; v.not_ = 50362; v.beg = 50363;
    v.multilingual = hp.n_vocab >= 51865;
    v.n_lang = hp.n_vocab - 51765 - (v.multilingual ? 1 : 0);
    if (v.multilingual) {
        v.eot++; v.sot++;
        const int dt = v.n_lang - 98;
        v.translate += dt; v.transcribe += dt; v.solm += dt; v.prev += dt; v.nosp += dt; v.not_ += dt; v.beg += dt;
    }
    v.id_to_token = std::move(toks);
    for (int i = (int)v.id_to_token.size(); i < hp.n_vocab; i++) {
        std::string w;
        if (i > v.beg) w = "[_TT_" + std::to_string(i - v.beg) + "]";
        else if (i == v.eot) w = "[_EOT_]";
        else if (i == v.sot) w = "[_SOT_]";
        else if (i == v.translate) w = "[_TRANSLATE_]";
        else if (i == v.transcribe) w = "[_TRANSCRIBE_]";
        else if (i == v.solm) w = "[_SOLM_]";
        else if (i == v.prev) w = "[_PREV_]";
        else if (i == v.nosp) w = "[_NOSP_]";
        else if (i == v.not_) w = "[_NOT_]";
        else if (i == v.beg) w = "[_BEG_]";
        else if (i > v.sot && i <= v.sot + v.n_lang && i - v.sot - 1 < kNumLangTable) w = std::string("[_LANG_") + g_lang[i - v.sot - 1] + "]";
        else w = "[_extra_token_" + std::to_string(i) + "]";
        v.id_to_token.push_back(w);
    }
    v.blank = -1;
    for (int i = 0; i < hp.n_vocab; i++) if (v.id_to_token[i] == " ") v.blank = i;   // token_to_id.at(" "): last wins
}

}  // namespace

std::vector<unsigned char> build_arena_image(const std::string &path) {
    ParsedFile pf;
    parse_file(path, pf);
    Model tmp; tmp.hp = pf.hp;
    // pass 1: size
    const size_t hdr = 16 + pf.meta_len;
    Walker sz{nullptr, nullptr, hdr};
    walk(sz, tmp);
    std::vector<unsigned char> img(((sz.off + 255) & ~(size_t)255), 0);
    memcpy(img.data(), kArenaMagic, 8);
    uint64_t ml = pf.meta_len; memcpy(img.data() + 8, &ml, 8);
    memcpy(img.data() + 16, pf.buf.data(), pf.meta_len);
    Walker fw{img.data(), &pf, hdr};
    walk(fw, tmp);
    tmp.arena = nullptr;
    return img;
}

void load_model_meta(const std::string &path, Model &m) {
    ParsedFile pf;
    parse_file(path, pf);
    std::vector<std::string> toks;
    parse_meta(pf.buf.data(), pf.meta_len, m.hp, nullptr, &toks);
    fill_vocab(m, std::move(toks));
}

ModelProbe probe_model(const std::string &path) {
    std::vector<unsigned char> img = build_arena_image(path);
    ModelProbe pr{};
    uint64_t ml; memcpy(&ml, img.data() + 8, 8);
    std::vector<std::string> toks;
    parse_meta(img.data() + 16, ml, pr.hp, nullptr, &toks);
    Model tmp; tmp.hp = pr.hp;
    pr.n_vocab_strings = (int)toks.size();
    fill_vocab(tmp, std::move(toks));
    pr.eot = tmp.vocab.eot; pr.beg = tmp.vocab.beg;
    pr.arena_bytes = img.size();
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : img) { h ^= c; h *= 1099511628211ull; }
    pr.fnv1a = h;
    return pr;
}

void bind_model(Model &m, unsigned char *d_arena, size_t bytes, int device) {
    m.device = device; m.arena = d_arena; m.arena_bytes = bytes;
    unsigned char head[16];
    CUDA_CHECK(cudaMemcpy(head, d_arena, 16, cudaMemcpyDeviceToHost));
    if (memcmp(head, kArenaMagic, 8)) SS_THROW(-9, "weight arena magic mismatch");
    uint64_t ml; memcpy(&ml, head + 8, 8);
    if (16 + ml > bytes) SS_THROW(-9, "weight arena meta length out of range");
    std::vector<unsigned char> meta(ml);
    CUDA_CHECK(cudaMemcpy(meta.data(), d_arena + 16, ml, cudaMemcpyDeviceToHost));
    std::vector<std::string> toks;
    parse_meta(meta.data(), meta.size(), m.hp, nullptr, &toks);
    m.meta_bytes = ml;
    fill_vocab(m, std::move(toks));
    Walker bw{d_arena, nullptr, 16 + (size_t)ml};
    walk(bw, m);
    if (((bw.off + 255) & ~(size_t)255) != bytes) SS_THROW(-9, "weight arena size mismatch (%zu vs %zu)", bw.off, bytes);
}

Model::~Model() {
    if (arena) { cudaSetDevice(device); cudaFree(arena); }
}

}  // namespace ss
