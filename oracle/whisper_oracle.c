/*
 * whisper_oracle.c - CPU restatement of the transcribe hot path.  TEST INFRASTRUCTURE ONLY
 * (see whisper_oracle.h).  PARITY UNPINNED against whisper.cpp itself (not in the reference
 * tree); pinned against HF transformers Whisper via tests/golden/.
 *
 * Follows, stage by stage (SURVEY.md Appendix A = [wcpp-knowledge] restatement of whisper.cpp
 * ~v1.5.x, the library behind /root/reference/src/asr/whisper.rs:75 `state.full`):
 *   A.1 ggml legacy .bin loader      <- WhisperContext::new_with_params   whisper.rs:23
 *   A.2 log-mel                      <- whisper_pcm_to_mel                (inside full, whisper.rs:75)
 *   A.3 encoder graph, A.4 cross-KV + decoder graph
 *   A.5 whisper_full loop: prompt, logits filter, greedy / sampled decoders, temperature fallback,
 *       segmentation.  Parameters are the ones build_params fixes (whisper.rs:131-173).
 * Arithmetic choices mirror ggml's CPU backend: f16 weights, activations rounded to f16 in front
 * of every mul_mat, f32 accumulation, tanh-GELU and exp through f16 rounding
 * (resources/ggml-metal.metal:262-277 GELU form, :351-435 softmax, :571-621 norm, :1778-1809 im2col f16).
 */
#define _GNU_SOURCE
#include "whisper_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WO_SAMPLE_RATE 16000
#define WO_N_FFT 400
#define WO_HOP 160
#define WO_CHUNK 30
#define WO_MAX_DECODERS 8

typedef _Float16 f16;

static __thread char g_err[512];
static int g_threads = 0;
static void set_err(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
const char *wo_last_error(void) { return g_err; }
void wo_set_threads(int n) {
    g_threads = n;
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

static inline float r16(float x) { return (float)(f16)x; }   /* round-trip through f16 (RNE) */

/* ------------------------------------------------------------------------------------------ */
/* model                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const f16 *w; const float *b; int n_out, n_in; } lin_t;
typedef struct { const float *w, *b; } ln_t;
typedef struct { ln_t attn_ln; lin_t q, k, v, o; ln_t mlp_ln; lin_t fc1, fc2; } enc_layer_t;
typedef struct { ln_t attn_ln; lin_t q, k, v, o; ln_t cross_ln; lin_t cq, ck, cv, co;
                 ln_t mlp_ln; lin_t fc1, fc2; } dec_layer_t;

typedef struct { char *s; int len; } tok_t;

struct wo_model {
    wo_hparams hp;
    unsigned char *blob; size_t blob_size;
    int filt_n_mel, filt_n_fft; const float *filters;
    int n_vocab_file; tok_t *tok;           /* n_vocab entries (extra ones synthesised) */
    /* special tokens */
    int eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, blank, multilingual, n_lang;
    /* encoder */
    const float *e_pos; lin_t conv1, conv2; enc_layer_t *enc; ln_t ln_post;
    /* decoder */
    const float *d_pos; const f16 *tok_emb; dec_layer_t *dec; ln_t d_ln;
};

static const char *g_lang[] = {
 "en","zh","de","es","ru","ko","fr","ja","pt","tr","pl","ca","nl","ar","sv","it","id","hi","fi","vi",
 "he","uk","el","ms","cs","ro","da","hu","ta","no","th","ur","hr","bg","lt","la","mi","ml","cy","sk",
 "te","fa","lv","bn","sr","az","sl","kn","et","mk","br","eu","is","hy","ne","mn","bs","kk","sq","sw",
 "gl","mr","pa","si","km","sn","yo","so","af","oc","ka","be","tg","sd","gu","am","yi","lo","uz","fo",
 "ht","ps","tk","nn","mt","sa","lb","my","bo","tl","mg","as","tt","haw","ln","ha","ba","jw","su","yue" };
#define N_LANG_TABLE 100

int wo_lang_id(const char *lang) {
    for (int i = 0; i < N_LANG_TABLE; i++) if (!strcmp(lang, g_lang[i])) return i;
    return -1;
}

/* ggml block-quantised tensor types, QK = 32 (GGML_TYPE 2 q4_0, 3 q4_1, 6 q5_0, 7 q5_1, 8 q8_0): ggml's dequantize_row_*.
 * The reference's model script fetches such files (script/download-ggml-model.sh:28-51).  NOTE: whisper.cpp multiplies
 * quantised weights against q8_0-quantised activations; this oracle (like the GPU loader) dequantises to f16 instead. */
static size_t quant_block_bytes(int tt) { return tt == 2 ? 18 : tt == 3 ? 20 : tt == 6 ? 22 : tt == 7 ? 24 : tt == 8 ? 34 : 0; }
__attribute__((optimize("fp-contract=off")))      /* x * d + m in two roundings, as ggml's scalar dequantize_row_* and numpy */
static void dequantize_block(const unsigned char *b, int tt, float *y) {
    f16 dh; memcpy(&dh, b, 2);
    const float d = (float)dh;
    if (tt == 8) { const int8_t *q = (const int8_t *)(b + 2); for (int j = 0; j < 32; j++) y[j] = (float)q[j] * d; return; }
    float m = 0.f; size_t o = 2;
    if (tt == 3 || tt == 7) { f16 mh; memcpy(&mh, b + 2, 2); m = (float)mh; o = 4; }
    uint32_t qh = 0;
    if (tt == 6 || tt == 7) { memcpy(&qh, b + o, 4); o += 4; }
    const unsigned char *qs = b + o;
    for (int j = 0; j < 16; j++) {
        int x0 = qs[j] & 0x0F, x1 = qs[j] >> 4;
        if (tt == 6 || tt == 7) { x0 |= (int)((qh >> j) & 1u) << 4; x1 |= (int)((qh >> (j + 16)) & 1u) << 4; }
        if (tt == 2) { y[j] = (float)(x0 - 8) * d; y[j + 16] = (float)(x1 - 8) * d; }
        else if (tt == 6) { y[j] = (float)(x0 - 16) * d; y[j + 16] = (float)(x1 - 16) * d; }
        else { y[j] = (float)x0 * d + m; y[j + 16] = (float)x1 * d + m; }
    }
}
typedef struct { const char *name; int n_dims; int ne[4]; int ttype; const void *data; } tensor_t;

static const tensor_t *find_tensor(const tensor_t *ts, int n, const char *name) {
    for (int i = 0; i < n; i++) if (!strcmp(ts[i].name, name)) return &ts[i];
    return NULL;
}

static int get_lin(const tensor_t *ts, int n, const char *prefix, int has_bias, lin_t *out) {
    char nm[256];
    snprintf(nm, sizeof nm, "%s.weight", prefix);
    const tensor_t *w = find_tensor(ts, n, nm);
    if (!w) { set_err("missing tensor %s", nm); return -1; }
    if (w->ttype != 1) { set_err("tensor %s: only f16 matrices are supported by the oracle", nm); return -1; }
    out->w = (const f16 *)w->data;
    if (w->n_dims == 3) { out->n_in = w->ne[0] * w->ne[1]; out->n_out = w->ne[2]; }
    else { out->n_in = w->ne[0]; out->n_out = w->ne[1]; }
    out->b = NULL;
    if (has_bias) {
        snprintf(nm, sizeof nm, "%s.bias", prefix);
        const tensor_t *b = find_tensor(ts, n, nm);
        if (!b || b->ttype != 0) { set_err("missing/non-f32 tensor %s", nm); return -1; }
        out->b = (const float *)b->data;
    }
    return 0;
}
static int get_ln(const tensor_t *ts, int n, const char *prefix, ln_t *out) {
    char nm[256];
    snprintf(nm, sizeof nm, "%s.weight", prefix);
    const tensor_t *w = find_tensor(ts, n, nm);
    snprintf(nm, sizeof nm, "%s.bias", prefix);
    const tensor_t *b = find_tensor(ts, n, nm);
    if (!w || !b || w->ttype != 0 || b->ttype != 0) { set_err("missing LN %s", prefix); return -1; }
    out->w = (const float *)w->data; out->b = (const float *)b->data;
    return 0;
}

wo_model *wo_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { set_err("cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END); size_t sz = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    unsigned char *blob = NULL;
    if (posix_memalign((void **)&blob, 64, sz + 64)) { fclose(f); set_err("oom"); return NULL; }
    if (fread(blob, 1, sz, f) != sz) { fclose(f); free(blob); set_err("short read"); return NULL; }
    fclose(f);
    wo_model *m = (wo_model *)calloc(1, sizeof *m);
    m->blob = blob; m->blob_size = sz;
    size_t o = 0;
#define NEED(nb) do { if (o + (size_t)(nb) > sz) { set_err("truncated model file at %zu", o); goto fail; } } while (0)
    NEED(4 + 44 + 8);
    uint32_t magic; memcpy(&magic, blob, 4); o = 4;
    if (magic != 0x67676d6c) { set_err("bad magic %08x", magic); goto fail; }
    memcpy(&m->hp, blob + o, 44); o += 44;
    memcpy(&m->filt_n_mel, blob + o, 4); memcpy(&m->filt_n_fft, blob + o + 4, 4); o += 8;
    /* filters may be unaligned in the blob: copy */
    {
        size_t nb = (size_t)m->filt_n_mel * m->filt_n_fft * 4; NEED(nb);
        float *fl = (float *)malloc(nb); memcpy(fl, blob + o, nb); m->filters = fl; o += nb;
    }
    NEED(4); memcpy(&m->n_vocab_file, blob + o, 4); o += 4;
    const wo_hparams *hp = &m->hp;
    if (hp->n_vocab < m->n_vocab_file || hp->n_vocab > 100000 || hp->n_mels != m->filt_n_mel ||
        m->filt_n_fft != 1 + WO_N_FFT / 2) { set_err("inconsistent header"); goto fail; }
    m->tok = (tok_t *)calloc((size_t)hp->n_vocab, sizeof(tok_t));
    for (int i = 0; i < m->n_vocab_file; i++) {
        uint32_t len; NEED(4); memcpy(&len, blob + o, 4); o += 4; NEED(len);
        m->tok[i].s = (char *)malloc(len + 1); memcpy(m->tok[i].s, blob + o, len); m->tok[i].s[len] = 0;
        m->tok[i].len = (int)len; o += len;
    }
    /* special tokens (Appendix A.5) */
    m->eot = 50256; m->sot = 50257; m->translate = 50357; m->transcribe = 50358; m->solm = 50359;
    m->prev = 50360; m->nosp = 50361; m->not_ = 50362; m->beg = 50363;
    m->multilingual = hp->n_vocab >= 51865;
    m->n_lang = hp->n_vocab - 51765 - (m->multilingual ? 1 : 0);
    if (m->multilingual) {
        m->eot++; m->sot++;
        int dt = m->n_lang - 98;
        m->translate += dt; m->transcribe += dt; m->solm += dt; m->prev += dt; m->nosp += dt;
        m->not_ += dt; m->beg += dt;
    }
    for (int i = m->n_vocab_file; i < hp->n_vocab; i++) {
        char w[64];
        if (i > m->beg) snprintf(w, sizeof w, "[_TT_%d]", i - m->beg);
        else if (i == m->eot) snprintf(w, sizeof w, "[_EOT_]");
        else if (i == m->sot) snprintf(w, sizeof w, "[_SOT_]");
        else if (i == m->translate) snprintf(w, sizeof w, "[_TRANSLATE_]");
        else if (i == m->transcribe) snprintf(w, sizeof w, "[_TRANSCRIBE_]");
        else if (i == m->solm) snprintf(w, sizeof w, "[_SOLM_]");
        else if (i == m->prev) snprintf(w, sizeof w, "[_PREV_]");
        else if (i == m->nosp) snprintf(w, sizeof w, "[_NOSP_]");
        else if (i == m->not_) snprintf(w, sizeof w, "[_NOT_]");
        else if (i == m->beg) snprintf(w, sizeof w, "[_BEG_]");
        else if (i > m->sot && i <= m->sot + m->n_lang && i - m->sot - 1 < N_LANG_TABLE)
            snprintf(w, sizeof w, "[_LANG_%s]", g_lang[i - m->sot - 1]);
        else snprintf(w, sizeof w, "[_extra_token_%d]", i);
        m->tok[i].s = strdup(w); m->tok[i].len = (int)strlen(w);
    }
    m->blank = -1;
    for (int i = 0; i < hp->n_vocab; i++) if (m->tok[i].len == 1 && m->tok[i].s[0] == ' ') m->blank = i;

    /* tensors */
    int cap = 2048, nt = 0;
    tensor_t *ts = (tensor_t *)calloc((size_t)cap, sizeof *ts);
    char **names = (char **)calloc((size_t)cap, sizeof(char *));
    while (o < sz) {
        int32_t hdr[3]; NEED(12); memcpy(hdr, blob + o, 12); o += 12;
        int n_dims = hdr[0], nlen = hdr[1], tt = hdr[2];
        if (n_dims < 1 || n_dims > 4 || nlen <= 0 || nlen > 200 || (tt != 0 && tt != 1 && !quant_block_bytes(tt))) {
            set_err("bad tensor header at %zu (n_dims %d, len %d, type %d)", o, n_dims, nlen, tt); free(ts); goto fail; }
        tensor_t *t = &ts[nt];
        size_t ne = 1;
        NEED(4 * n_dims + nlen);
        for (int d = 0; d < n_dims; d++) { memcpy(&t->ne[d], blob + o, 4); o += 4; ne *= (size_t)t->ne[d]; }
        names[nt] = (char *)malloc((size_t)nlen + 1); memcpy(names[nt], blob + o, (size_t)nlen); names[nt][nlen] = 0; o += (size_t)nlen;
        t->name = names[nt]; t->n_dims = n_dims; t->ttype = tt;
        size_t nb = ne * (tt == 1 ? 2 : 4);
        if (quant_block_bytes(tt)) {      /* block-quantised (whisper.cpp `quantize` output): dequantise once into f16, as the GPU loader does */
            if (t->ne[0] % 32) { set_err("quantised tensor %s: row length not a multiple of 32", t->name); free(ts); goto fail; }
            nb = ne / 32 * quant_block_bytes(tt); NEED(nb);
            f16 *c = (f16 *)malloc(ne * sizeof(f16));   /* leaked with the model */
            float y[32];
            for (size_t i = 0; i < ne / 32; i++) { dequantize_block(blob + o + i * quant_block_bytes(tt), tt, y); for (int j = 0; j < 32; j++) c[32 * i + j] = (f16)y[j]; }
            t->data = c; t->ttype = 1; o += nb;
            if (++nt == cap) { set_err("too many tensors"); free(ts); goto fail; }
            continue;
        }
        NEED(nb);
        /* the legacy container does not align tensor data; realign into an owned buffer when needed */
        if (((uintptr_t)(blob + o)) % (tt == 1 ? 2 : 4)) {
            void *c = malloc(nb); memcpy(c, blob + o, nb); t->data = c;   /* leaked with the model */
        } else t->data = blob + o;
        o += nb;
        if (++nt == cap) { set_err("too many tensors"); free(ts); goto fail; }
    }
    {
        char nm[256]; const tensor_t *t;
        int rc = 0;
        t = find_tensor(ts, nt, "encoder.positional_embedding"); if (!t) { set_err("no encoder pos"); rc = -1; } else m->e_pos = (const float *)t->data;
        t = find_tensor(ts, nt, "decoder.positional_embedding"); if (!t) { set_err("no decoder pos"); rc = -1; } else m->d_pos = (const float *)t->data;
        t = find_tensor(ts, nt, "decoder.token_embedding.weight"); if (!t || t->ttype != 1) { set_err("no f16 token embedding"); rc = -1; } else m->tok_emb = (const f16 *)t->data;
        rc |= get_lin(ts, nt, "encoder.conv1", 1, &m->conv1);
        rc |= get_lin(ts, nt, "encoder.conv2", 1, &m->conv2);
        rc |= get_ln(ts, nt, "encoder.ln_post", &m->ln_post);
        rc |= get_ln(ts, nt, "decoder.ln", &m->d_ln);
        m->enc = (enc_layer_t *)calloc((size_t)hp->n_audio_layer, sizeof(enc_layer_t));
        m->dec = (dec_layer_t *)calloc((size_t)hp->n_text_layer, sizeof(dec_layer_t));
        for (int i = 0; i < hp->n_audio_layer && !rc; i++) {
            enc_layer_t *L = &m->enc[i];
#define P(suffix) (snprintf(nm, sizeof nm, "encoder.blocks.%d." suffix, i), nm)
            rc |= get_ln(ts, nt, P("attn_ln"), &L->attn_ln);
            rc |= get_lin(ts, nt, P("attn.query"), 1, &L->q);
            rc |= get_lin(ts, nt, P("attn.key"), 0, &L->k);
            rc |= get_lin(ts, nt, P("attn.value"), 1, &L->v);
            rc |= get_lin(ts, nt, P("attn.out"), 1, &L->o);
            rc |= get_ln(ts, nt, P("mlp_ln"), &L->mlp_ln);
            rc |= get_lin(ts, nt, P("mlp.0"), 1, &L->fc1);
            rc |= get_lin(ts, nt, P("mlp.2"), 1, &L->fc2);
#undef P
        }
        for (int i = 0; i < hp->n_text_layer && !rc; i++) {
            dec_layer_t *L = &m->dec[i];
#define P(suffix) (snprintf(nm, sizeof nm, "decoder.blocks.%d." suffix, i), nm)
            rc |= get_ln(ts, nt, P("attn_ln"), &L->attn_ln);
            rc |= get_lin(ts, nt, P("attn.query"), 1, &L->q);
            rc |= get_lin(ts, nt, P("attn.key"), 0, &L->k);
            rc |= get_lin(ts, nt, P("attn.value"), 1, &L->v);
            rc |= get_lin(ts, nt, P("attn.out"), 1, &L->o);
            rc |= get_ln(ts, nt, P("cross_attn_ln"), &L->cross_ln);
            rc |= get_lin(ts, nt, P("cross_attn.query"), 1, &L->cq);
            rc |= get_lin(ts, nt, P("cross_attn.key"), 0, &L->ck);
            rc |= get_lin(ts, nt, P("cross_attn.value"), 1, &L->cv);
            rc |= get_lin(ts, nt, P("cross_attn.out"), 1, &L->co);
            rc |= get_ln(ts, nt, P("mlp_ln"), &L->mlp_ln);
            rc |= get_lin(ts, nt, P("mlp.0"), 1, &L->fc1);
            rc |= get_lin(ts, nt, P("mlp.2"), 1, &L->fc2);
#undef P
        }
        for (int i = 0; i < nt; i++) free(names[i]);
        free(names); free(ts);
        if (rc) goto fail;
    }
    return m;
fail:
    wo_free(m);
    return NULL;
#undef NEED
}

void wo_free(wo_model *m) {
    if (!m) return;
    if (m->tok) { for (int i = 0; i < m->hp.n_vocab; i++) free(m->tok[i].s); free(m->tok); }
    free((void *)m->filters); free(m->enc); free(m->dec); free(m->blob); free(m);
}
const wo_hparams *wo_get_hparams(const wo_model *m) { return &m->hp; }
int wo_token_id(const wo_model *m, const char *n) {
    if (!strcmp(n, "eot")) return m->eot;
    if (!strcmp(n, "sot")) return m->sot;
    if (!strcmp(n, "translate")) return m->translate;
    if (!strcmp(n, "transcribe")) return m->transcribe;
    if (!strcmp(n, "solm")) return m->solm;
    if (!strcmp(n, "prev")) return m->prev;
    if (!strcmp(n, "nosp")) return m->nosp;
    if (!strcmp(n, "not")) return m->not_;
    if (!strcmp(n, "beg")) return m->beg;
    if (!strcmp(n, "blank")) return m->blank;
    return -1;
}
int wo_token_bytes(const wo_model *m, int id, const char **b) {
    if (id < 0 || id >= m->hp.n_vocab) return -1;
    *b = m->tok[id].s; return m->tok[id].len;
}

/* ------------------------------------------------------------------------------------------ */
/* dense kernels: C[M][N] = A[M][K] (f32, already f16-rounded) x W[N][K]^T (f16), f32 accumulate */
/* ------------------------------------------------------------------------------------------ */
typedef float v16f __attribute__((vector_size(64), aligned(4)));
typedef f16 v16h __attribute__((vector_size(32), aligned(2)));

static inline float hsum16(v16f v) {
    float s = 0.f; for (int i = 0; i < 16; i++) s += v[i]; return s;
}

/* 4 rows of A x 4 rows of W (both f32, K-contiguous) */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void micro_4x4(const float *a, size_t lda, const float *w, size_t ldw, int K, float *c, size_t ldc,
                      int mr, int nr) {
    v16f acc[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i][j] = (v16f){0};
    const float *a0 = a, *a1 = a + (mr > 1 ? lda : 0), *a2 = a + (mr > 2 ? 2 * lda : 0), *a3 = a + (mr > 3 ? 3 * lda : 0);
    const float *w0 = w, *w1 = w + (nr > 1 ? ldw : 0), *w2 = w + (nr > 2 ? 2 * ldw : 0), *w3 = w + (nr > 3 ? 3 * ldw : 0);
    int k = 0;
    for (; k + 16 <= K; k += 16) {
        v16f x0 = *(const v16f *)(a0 + k), x1 = *(const v16f *)(a1 + k), x2 = *(const v16f *)(a2 + k), x3 = *(const v16f *)(a3 + k);
        v16f y0 = *(const v16f *)(w0 + k), y1 = *(const v16f *)(w1 + k), y2 = *(const v16f *)(w2 + k), y3 = *(const v16f *)(w3 + k);
        acc[0][0] += x0 * y0; acc[0][1] += x0 * y1; acc[0][2] += x0 * y2; acc[0][3] += x0 * y3;
        acc[1][0] += x1 * y0; acc[1][1] += x1 * y1; acc[1][2] += x1 * y2; acc[1][3] += x1 * y3;
        acc[2][0] += x2 * y0; acc[2][1] += x2 * y1; acc[2][2] += x2 * y2; acc[2][3] += x2 * y3;
        acc[3][0] += x3 * y0; acc[3][1] += x3 * y1; acc[3][2] += x3 * y2; acc[3][3] += x3 * y3;
    }
    float tail[4][4] = {{0}};
    for (; k < K; k++) {
        float xa[4] = {a0[k], a1[k], a2[k], a3[k]}, yw[4] = {w0[k], w1[k], w2[k], w3[k]};
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) tail[i][j] += xa[i] * yw[j];
    }
    for (int i = 0; i < mr; i++) for (int j = 0; j < nr; j++) c[i * ldc + j] = hsum16(acc[i][j]) + tail[i][j];
}

__attribute__((target_clones("avx512f", "avx2", "default")))
static void cvt_rows_f16(const f16 *w, size_t ldw, int rows, int K, float *out) {
    for (int r = 0; r < rows; r++) {
        const f16 *s = w + r * ldw; float *d = out + (size_t)r * K;
        int k = 0;
        for (; k + 16 <= K; k += 16) { v16h h = *(const v16h *)(s + k); *(v16f *)(d + k) = __builtin_convertvector(h, v16f); }
        for (; k < K; k++) d[k] = (float)s[k];
    }
}

/* one activation row against a block of f16 weight rows, converting on the fly (decode GEMV) */
__attribute__((target_clones("avx512f", "avx2", "default")))
static void gemv_rows_f16(const float *a, const f16 *W, size_t ldw, int n0, int n1, int K, float *c) {
    for (int n = n0; n < n1; n++) {
        const f16 *w = W + (size_t)n * ldw;
        v16f acc0 = (v16f){0}, acc1 = (v16f){0};
        int k = 0;
        for (; k + 32 <= K; k += 32) {
            acc0 += *(const v16f *)(a + k) * __builtin_convertvector(*(const v16h *)(w + k), v16f);
            acc1 += *(const v16f *)(a + k + 16) * __builtin_convertvector(*(const v16h *)(w + k + 16), v16f);
        }
        float t = 0.f;
        for (; k < K; k++) t += a[k] * (float)w[k];
        c[n] = hsum16(acc0 + acc1) + t;
    }
}

/* C = A W^T.  A: [M][K] f32 lda; W: [N][K] f16 ldw; C: [M][N] f32 ldc.  bias[N] optional. */
static void gemm_f16w(const float *A, size_t lda, const f16 *W, size_t ldw, float *C, size_t ldc,
                      int M, int N, int K, const float *bias) {
    const int MC = 48;
    if (M == 1) {
        const int NB = 64; int nblk = (N + NB - 1) / NB;
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nblk; b++) gemv_rows_f16(A, W, ldw, b * NB, (b + 1) * NB < N ? (b + 1) * NB : N, K, C);
    } else if (M >= 32) {
        int nchunk = (M + MC - 1) / MC;
#pragma omp parallel
        {
            float *wf = (float *)aligned_alloc(64, (((size_t)4 * K * sizeof(float)) + 127) & ~(size_t)63);
#pragma omp for schedule(dynamic, 1)
            for (int ch = 0; ch < nchunk; ch++) {
                int m0 = ch * MC, m1 = m0 + MC > M ? M : m0 + MC;
                for (int n0 = 0; n0 < N; n0 += 4) {
                    int nr = N - n0 < 4 ? N - n0 : 4;
                    cvt_rows_f16(W + (size_t)n0 * ldw, ldw, nr, K, wf);
                    for (int i = m0; i < m1; i += 4) {
                        int mr = m1 - i < 4 ? m1 - i : 4;
                        micro_4x4(A + (size_t)i * lda, lda, wf, (size_t)K, K, C + (size_t)i * ldc + n0, ldc, mr, nr);
                    }
                }
            }
            free(wf);
        }
    } else {
        int nblk = (N + 3) / 4;
#pragma omp parallel
        {
            float *wf = (float *)aligned_alloc(64, (((size_t)4 * K * sizeof(float)) + 127) & ~(size_t)63);
#pragma omp for schedule(static)
            for (int b = 0; b < nblk; b++) {
                int n0 = b * 4, nr = N - n0 < 4 ? N - n0 : 4;
                cvt_rows_f16(W + (size_t)n0 * ldw, ldw, nr, K, wf);
                for (int i = 0; i < M; i += 4) {
                    int mr = M - i < 4 ? M - i : 4;
                    micro_4x4(A + (size_t)i * lda, lda, wf, (size_t)K, K, C + (size_t)i * ldc + n0, ldc, mr, nr);
                }
            }
            free(wf);
        }
    }
    if (bias) {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < M; i++) { float *c = C + (size_t)i * ldc; for (int j = 0; j < N; j++) c[j] += bias[j]; }
    }
}

/* C = A B^T with both operands f32 (used by attention, operands already f16-rounded) */
static void gemm_f32(const float *A, size_t lda, const float *B, size_t ldb, float *C, size_t ldc, int M, int N, int K) {
    for (int i = 0; i < M; i += 4) {
        int mr = M - i < 4 ? M - i : 4;
        for (int j = 0; j < N; j += 4) {
            int nr = N - j < 4 ? N - j : 4;
            micro_4x4(A + (size_t)i * lda, lda, B + (size_t)j * ldb, ldb, K, C + (size_t)i * ldc + j, ldc, mr, nr);
        }
    }
}

static void round_f16_inplace(float *x, size_t n) {
#pragma omp parallel for schedule(static) if (n > 65536)
    for (size_t i = 0; i < n; i++) x[i] = r16(x[i]);
}

/* ggml_norm + mul + add.  mean / variance in double like ggml_float.  out may alias nothing. */
static void layer_norm(const float *x, float *y, int rows, int d, const ln_t *ln, int round_out) {
#pragma omp parallel for schedule(static) if (rows > 8)
    for (int r = 0; r < rows; r++) {
        const float *xi = x + (size_t)r * d; float *yi = y + (size_t)r * d;
        double sum = 0.0; for (int i = 0; i < d; i++) sum += (double)xi[i];
        float mean = (float)(sum / d);
        double sum2 = 0.0;
        for (int i = 0; i < d; i++) { float v = xi[i] - mean; yi[i] = v; sum2 += (double)(v * v); }
        float var = (float)(sum2 / d);
        const float scale = 1.0f / sqrtf(var + 1e-5f);
        for (int i = 0; i < d; i++) { float v = yi[i] * scale * ln->w[i] + ln->b[i]; yi[i] = round_out ? r16(v) : v; }
    }
}

/* ggml CPU GELU: f16 lookup table == f16(gelu_f32(f16(x))) */
static inline float gelu_ggml(float x) {
    const float xh = r16(x);
    const float g = 0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh)));
    return r16(g);
}
/* ggml CPU soft_max of this era: exp through the f16 table == f16(expf(f16(x - max))), sum in double */
static void softmax_row(float *x, int n) {
    float mx = -INFINITY; for (int i = 0; i < n; i++) if (x[i] > mx) mx = x[i];
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        if (x[i] == -INFINITY) { x[i] = 0.f; continue; }
        float e = r16(expf(r16(x[i] - mx))); x[i] = e; sum += (double)e;
    }
    const float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) x[i] *= inv;
}

/* ------------------------------------------------------------------------------------------ */
/* log-mel (Appendix A.2)                                                                       */
/* ------------------------------------------------------------------------------------------ */
static float g_sin[WO_N_FFT], g_cos[WO_N_FFT], g_hann[WO_N_FFT];
static int g_tab_init = 0;
static void init_tables(void) {
    if (g_tab_init) return;
    for (int i = 0; i < WO_N_FFT; i++) {
        double t = (2.0 * M_PI * i) / WO_N_FFT;
        g_sin[i] = sinf((float)t); g_cos[i] = cosf((float)t);
        g_hann[i] = 0.5f * (1.0f - cosf((float)((2.0 * M_PI * i) / WO_N_FFT)));
    }
    g_tab_init = 1;
}
static void dft_naive(const float *in, int N, float *out) {
    const int step = WO_N_FFT / N;
    for (int k = 0; k < N; k++) {
        float re = 0, im = 0;
        for (int n = 0; n < N; n++) {
            int idx = (k * n * step) % WO_N_FFT;
            re += in[n] * g_cos[idx]; im -= in[n] * g_sin[idx];
        }
        out[2 * k] = re; out[2 * k + 1] = im;
    }
}
/* recursive radix-2 down to odd sizes, all f32, same operation order as whisper.cpp's fft() */
static void fft_rec(const float *in, int N, float *out, float *scratch) {
    if (N == 1) { out[0] = in[0]; out[1] = 0; return; }
    if (N % 2 == 1) { dft_naive(in, N, out); return; }
    int h = N / 2;
    float *even = scratch, *odd = scratch + h, *ef = scratch + N, *of = scratch + N + 2 * h, *next = scratch + N + 4 * h;
    for (int i = 0; i < h; i++) { even[i] = in[2 * i]; odd[i] = in[2 * i + 1]; }
    fft_rec(even, h, ef, next); fft_rec(odd, h, of, next);
    const int step = WO_N_FFT / N;
    for (int k = 0; k < h; k++) {
        int idx = k * step; float re = g_cos[idx], im = -g_sin[idx];
        float ro = of[2 * k], io = of[2 * k + 1];
        out[2 * k] = ef[2 * k] + re * ro - im * io;
        out[2 * k + 1] = ef[2 * k + 1] + re * io + im * ro;
        out[2 * (k + h)] = ef[2 * k] - re * ro + im * io;
        out[2 * (k + h) + 1] = ef[2 * k + 1] - re * io - im * ro;
    }
}

int wo_log_mel(const wo_model *m, const float *pcm, size_t n, float **mel_out, int *n_len_out, int *n_len_org_out) {
    init_tables();
    const int n_mel = m->hp.n_mels, n_fft = 1 + WO_N_FFT / 2;
    const size_t pad1 = (size_t)WO_SAMPLE_RATE * WO_CHUNK, pad2 = WO_N_FFT / 2;
    const size_t np = n + pad1 + 2 * pad2;
    float *sp = (float *)calloc(np, sizeof(float));
    memcpy(sp + pad2, pcm, n * sizeof(float));
    /* reflective pad at the beginning: reverse of samples[1..=200] */
    for (size_t i = 0; i < pad2; i++) sp[i] = (pad2 - i < n) ? pcm[pad2 - i] : 0.f;
    const int n_len = (int)((np - WO_N_FFT) / WO_HOP);
    const int n_len_org = 1 + (int)(((long)n + (long)pad2 - WO_N_FFT) / WO_HOP);
    const long n_eff = (long)(n + pad2);          /* worker's n_samples */
    float *mel = (float *)malloc((size_t)n_mel * n_len * sizeof(float));
    long n_calc = n_eff / WO_HOP + 1; if (n_calc > n_len) n_calc = n_len;
#pragma omp parallel
    {
        float fin[WO_N_FFT], fout[2 * WO_N_FFT], scratch[8 * WO_N_FFT];
#pragma omp for schedule(static)
        for (int i = 0; i < n_len; i++) {
            if (i >= n_calc) { for (int j = 0; j < n_mel; j++) mel[(size_t)j * n_len + i] = (float)log10(1e-10); continue; }
            const long off = (long)i * WO_HOP;
            long lim = n_eff - off; if (lim > WO_N_FFT) lim = WO_N_FFT;
            for (long j = 0; j < lim; j++) fin[j] = g_hann[j] * sp[off + j];
            for (long j = lim < 0 ? 0 : lim; j < WO_N_FFT; j++) fin[j] = 0.f;
            fft_rec(fin, WO_N_FFT, fout, scratch);
            for (int j = 0; j < WO_N_FFT; j++) fout[j] = fout[2 * j] * fout[2 * j] + fout[2 * j + 1] * fout[2 * j + 1];
            for (int j = 0; j < n_mel; j++) {
                const float *fl = m->filters + (size_t)j * n_fft;
                double sum = 0.0; int k = 0;
                for (; k < n_fft - 3; k += 4)
                    sum += fout[k] * fl[k] + fout[k + 1] * fl[k + 1] + fout[k + 2] * fl[k + 2] + fout[k + 3] * fl[k + 3];
                for (; k < n_fft; k++) sum += fout[k] * fl[k];
                sum = log10(sum > 1e-10 ? sum : 1e-10);
                mel[(size_t)j * n_len + i] = (float)sum;
            }
        }
    }
    double mmax = -1e20;
    for (size_t i = 0; i < (size_t)n_mel * n_len; i++) if (mel[i] > mmax) mmax = mel[i];
    mmax -= 8.0;
    for (size_t i = 0; i < (size_t)n_mel * n_len; i++) {
        if (mel[i] < mmax) mel[i] = (float)mmax;
        mel[i] = (float)((mel[i] + 4.0) / 4.0);
    }
    free(sp);
    *mel_out = mel; *n_len_out = n_len; *n_len_org_out = n_len_org;
    return 0;
}
void wo_free_buf(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------ */
/* state                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int id, tid; float p, plog, pt, ptsum; } tokdata_t;
typedef struct { tokdata_t *tokens; int n, cap; int result_len; double sum_logprobs_all, sum_logprobs, avg_logprobs, entropy, score; } sequence_t;
typedef struct {
    sequence_t seq; int seek_delta, failed, completed, has_ts;
    float *probs, *logits, *logprobs;
    uint32_t mt[624]; int mti;
    double min_margin; long n_draws;   /* diagnostics of the draws since the last wo_full started (wo_min_sample_margin) */
} decoder_t;
typedef struct { int64_t t0, t1; char *text; int speaker_turn_next; } segment_t;

struct wo_state {
    wo_model *m;
    float *enc_out;                    /* [T][d] */
    f16 *cross_k, *cross_v;            /* [layer][T][d] */
    f16 *self_k, *self_v;              /* [seq][layer][n_text_ctx][d] */
    decoder_t dec[WO_MAX_DECODERS];
    int *prompt_past; int n_prompt_past, cap_prompt_past;
    segment_t *segs; int n_segs, cap_segs;
    int *res_tok; float *res_p, *res_plog; int n_res, cap_res;
    int n_fallbacks, n_decoded, n_windows;
    float *kept; int n_kept, cap_kept;
};

static void mt_seed(decoder_t *d, uint32_t s) {
    d->mt[0] = s;
    for (int i = 1; i < 624; i++) d->mt[i] = 1812433253u * (d->mt[i - 1] ^ (d->mt[i - 1] >> 30)) + (uint32_t)i;
    d->mti = 624;
}
static uint32_t mt_next(decoder_t *d) {
    if (d->mti >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (d->mt[i] & 0x80000000u) | (d->mt[(i + 1) % 624] & 0x7fffffffu);
            d->mt[i] = d->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        d->mti = 0;
    }
    uint32_t y = d->mt[d->mti++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

wo_state *wo_state_new(wo_model *m) {
    const wo_hparams *hp = &m->hp;
    wo_state *s = (wo_state *)calloc(1, sizeof *s);
    s->m = m;
    size_t T = (size_t)hp->n_audio_ctx, d = (size_t)hp->n_text_state;
    s->enc_out = (float *)malloc(T * hp->n_audio_state * sizeof(float));
    s->cross_k = (f16 *)malloc((size_t)hp->n_text_layer * T * d * sizeof(f16));
    s->cross_v = (f16 *)malloc((size_t)hp->n_text_layer * T * d * sizeof(f16));
    size_t kv = (size_t)WO_MAX_DECODERS * hp->n_text_layer * hp->n_text_ctx * d;
    s->self_k = (f16 *)calloc(kv, sizeof(f16));
    s->self_v = (f16 *)calloc(kv, sizeof(f16));
    for (int j = 0; j < WO_MAX_DECODERS; j++) {
        decoder_t *dc = &s->dec[j];
        dc->probs = (float *)malloc((size_t)hp->n_vocab * sizeof(float));
        dc->logits = (float *)malloc((size_t)hp->n_vocab * sizeof(float));
        dc->logprobs = (float *)malloc((size_t)hp->n_vocab * sizeof(float));
        mt_seed(dc, 0);
    }
    return s;
}
static void clear_segments(wo_state *s) {
    for (int i = 0; i < s->n_segs; i++) free(s->segs[i].text);
    s->n_segs = 0;
}
void wo_state_free(wo_state *s) {
    if (!s) return;
    clear_segments(s);
    free(s->segs); free(s->enc_out); free(s->cross_k); free(s->cross_v); free(s->self_k); free(s->self_v);
    for (int j = 0; j < WO_MAX_DECODERS; j++) { free(s->dec[j].probs); free(s->dec[j].logits); free(s->dec[j].logprobs); free(s->dec[j].seq.tokens); }
    free(s->prompt_past); free(s->res_tok); free(s->res_p); free(s->res_plog); free(s->kept); free(s);
}
const float *wo_encoder_out(const wo_state *s) { return s->enc_out; }

/* ------------------------------------------------------------------------------------------ */
/* encoder (Appendix A.3) + cross-KV (A.4)                                                      */
/* ------------------------------------------------------------------------------------------ */
static void linear(const float *x_r16, int M, const lin_t *L, float *y) {
    gemm_f16w(x_r16, (size_t)L->n_in, L->w, (size_t)L->n_in, y, (size_t)L->n_out, M, L->n_out, L->n_in, L->b);
}

int wo_encode(wo_state *s, const float *mel, int n_len, int seek) {
    const wo_model *m = s->m; const wo_hparams *hp = &m->hp;
    const int T = hp->n_audio_ctx, T2 = 2 * T, d = hp->n_audio_state, nm = hp->n_mels, H = hp->n_audio_head, dh = d / H;
    /* conv1: im2col in f16 (zero padded), K index = c*3 + k */
    float *col1 = (float *)calloc((size_t)T2 * nm * 3, sizeof(float));
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T2; t++)
        for (int c = 0; c < nm; c++)
            for (int k = 0; k < 3; k++) {
                int tt = t + k - 1; float v = 0.f;
                if (tt >= 0 && tt < T2) { int fr = seek + tt; v = fr < n_len ? mel[(size_t)c * n_len + fr] : 0.f; }
                col1[((size_t)t * nm + c) * 3 + k] = r16(v);
            }
    float *x1 = (float *)malloc((size_t)T2 * d * sizeof(float));
    linear(col1, T2, &m->conv1, x1);
    free(col1);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < (size_t)T2 * d; i++) x1[i] = gelu_ggml(x1[i]);
    /* conv2 stride 2 */
    float *col2 = (float *)malloc((size_t)T * d * 3 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int t = 0; t < T; t++)
        for (int c = 0; c < d; c++)
            for (int k = 0; k < 3; k++) {
                int tt = 2 * t + k - 1;
                col2[((size_t)t * d + c) * 3 + k] = (tt >= 0 && tt < T2) ? r16(x1[(size_t)tt * d + c]) : 0.f;
            }
    free(x1);
    float *x = (float *)malloc((size_t)T * d * sizeof(float));
    linear(col2, T, &m->conv2, x);
    free(col2);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < (size_t)T * d; i++) x[i] = gelu_ggml(x[i]) + m->e_pos[i];

    float *h = (float *)malloc((size_t)T * d * sizeof(float));
    float *q = (float *)malloc((size_t)T * d * sizeof(float));
    float *kk = (float *)malloc((size_t)T * d * sizeof(float));
    float *v = (float *)malloc((size_t)T * d * sizeof(float));
    float *att = (float *)malloc((size_t)T * d * sizeof(float));
    float *y = (float *)malloc((size_t)T * d * sizeof(float));
    float *ff = (float *)malloc((size_t)T * 4 * d * sizeof(float));
    const float kq_scale = 1.0f / sqrtf((float)dh);
    for (int il = 0; il < hp->n_audio_layer; il++) {
        const enc_layer_t *L = &m->enc[il];
        layer_norm(x, h, T, d, &L->attn_ln, 1);
        linear(h, T, &L->q, q); linear(h, T, &L->k, kk); linear(h, T, &L->v, v);
        round_f16_inplace(q, (size_t)T * d); round_f16_inplace(kk, (size_t)T * d); round_f16_inplace(v, (size_t)T * d);
#pragma omp parallel
        {
            float *S = (float *)malloc((size_t)T * T * sizeof(float));
            float *vt = (float *)malloc((size_t)dh * T * sizeof(float));
            float *o = (float *)malloc((size_t)T * dh * sizeof(float));
#pragma omp for schedule(dynamic, 1)
            for (int hh = 0; hh < H; hh++) {
                gemm_f32(q + hh * dh, (size_t)d, kk + hh * dh, (size_t)d, S, (size_t)T, T, T, dh);
                for (int i = 0; i < T; i++) {
                    float *row = S + (size_t)i * T;
                    for (int j = 0; j < T; j++) row[j] *= kq_scale;
                    softmax_row(row, T);
                    for (int j = 0; j < T; j++) row[j] = r16(row[j]);
                }
                for (int t = 0; t < T; t++) for (int c = 0; c < dh; c++) vt[(size_t)c * T + t] = v[(size_t)t * d + hh * dh + c];
                gemm_f32(S, (size_t)T, vt, (size_t)T, o, (size_t)dh, T, dh, T);
                for (int t = 0; t < T; t++) for (int c = 0; c < dh; c++) att[(size_t)t * d + hh * dh + c] = r16(o[(size_t)t * dh + c]);
            }
            free(S); free(vt); free(o);
        }
        linear(att, T, &L->o, y);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)T * d; i++) x[i] += y[i];
        layer_norm(x, h, T, d, &L->mlp_ln, 1);
        linear(h, T, &L->fc1, ff);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)T * 4 * d; i++) ff[i] = gelu_ggml(ff[i]);
        linear(ff, T, &L->fc2, y);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)T * d; i++) x[i] += y[i];
    }
    layer_norm(x, s->enc_out, T, d, &m->ln_post, 0);

    /* cross-KV: K pre-scaled by dh^-1/4, stored f16 */
    const int dd = hp->n_text_state; const float s4 = powf((float)(dd / hp->n_text_head), -0.25f);
    memcpy(h, s->enc_out, (size_t)T * d * sizeof(float));
    round_f16_inplace(h, (size_t)T * d);
    for (int il = 0; il < hp->n_text_layer; il++) {
        const dec_layer_t *L = &m->dec[il];
        linear(h, T, &L->ck, y);
        f16 *ck = s->cross_k + (size_t)il * T * dd, *cv = s->cross_v + (size_t)il * T * dd;
        for (size_t i = 0; i < (size_t)T * dd; i++) ck[i] = (f16)(y[i] * s4);
        linear(h, T, &L->cv, y);
        for (size_t i = 0; i < (size_t)T * dd; i++) cv[i] = (f16)y[i];
    }
    free(x); free(h); free(q); free(kk); free(v); free(att); free(y); free(ff);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* decoder (Appendix A.4)                                                                       */
/* ------------------------------------------------------------------------------------------ */
static void attend(const float *q, const f16 *K, const f16 *V, size_t ld, int n_keys, int dh, float *out, float *sc) {
    /* q: [dh] f16-rounded; K,V rows at stride ld; out [dh] */
    for (int j = 0; j < n_keys; j++) {
        const f16 *kj = K + (size_t)j * ld; float a = 0.f;
        for (int c = 0; c < dh; c++) a += q[c] * (float)kj[c];
        sc[j] = a;
    }
    softmax_row(sc, n_keys);
    for (int c = 0; c < dh; c++) out[c] = 0.f;
    for (int j = 0; j < n_keys; j++) {
        const float p = r16(sc[j]); const f16 *vj = V + (size_t)j * ld;
        for (int c = 0; c < dh; c++) out[c] += p * (float)vj[c];
    }
}

int wo_decode(wo_state *s, int seq, const int *tokens, int n, int n_past, float *logits_out) {
    const wo_model *m = s->m; const wo_hparams *hp = &m->hp;
    const int d = hp->n_text_state, H = hp->n_text_head, dh = d / H, T = hp->n_audio_ctx, nctx = hp->n_text_ctx;
    if (n_past + n > nctx || seq < 0 || seq >= WO_MAX_DECODERS) { set_err("decode: context overflow"); return -1; }
    const float s4 = powf((float)dh, -0.25f);
    float *x = (float *)malloc((size_t)n * d * sizeof(float)), *h = (float *)malloc((size_t)n * d * sizeof(float));
    float *q = (float *)malloc((size_t)n * d * sizeof(float)), *kk = (float *)malloc((size_t)n * d * sizeof(float));
    float *v = (float *)malloc((size_t)n * d * sizeof(float)), *att = (float *)malloc((size_t)n * d * sizeof(float));
    float *y = (float *)malloc((size_t)n * d * sizeof(float)), *ff = (float *)malloc((size_t)n * 4 * d * sizeof(float));
    for (int i = 0; i < n; i++) {
        const f16 *e = m->tok_emb + (size_t)tokens[i] * d; const float *pe = m->d_pos + (size_t)(n_past + i) * d;
        for (int c = 0; c < d; c++) x[(size_t)i * d + c] = (float)e[c] + pe[c];
    }
    for (int il = 0; il < hp->n_text_layer; il++) {
        const dec_layer_t *L = &m->dec[il];
        f16 *sk = s->self_k + (((size_t)seq * hp->n_text_layer + il) * nctx) * d;
        f16 *sv = s->self_v + (((size_t)seq * hp->n_text_layer + il) * nctx) * d;
        layer_norm(x, h, n, d, &L->attn_ln, 1);
        linear(h, n, &L->q, q); linear(h, n, &L->k, kk); linear(h, n, &L->v, v);
        for (int i = 0; i < n; i++)
            for (int c = 0; c < d; c++) {
                q[(size_t)i * d + c] = r16(q[(size_t)i * d + c] * s4);
                sk[(size_t)(n_past + i) * d + c] = (f16)(kk[(size_t)i * d + c] * s4);
                sv[(size_t)(n_past + i) * d + c] = (f16)v[(size_t)i * d + c];
            }
#pragma omp parallel
        {
            float *sc = (float *)malloc((size_t)(nctx > T ? nctx : T) * sizeof(float)); float o[256];
#pragma omp for schedule(static) collapse(2)
            for (int i = 0; i < n; i++)
                for (int hh = 0; hh < H; hh++) {
                    attend(q + (size_t)i * d + hh * dh, sk + hh * dh, sv + hh * dh, (size_t)d, n_past + i + 1, dh, o, sc);
                    for (int c = 0; c < dh; c++) att[(size_t)i * d + hh * dh + c] = r16(o[c]);
                }
            free(sc);
        }
        linear(att, n, &L->o, y);
        for (size_t i = 0; i < (size_t)n * d; i++) x[i] += y[i];
        /* cross attention */
        layer_norm(x, h, n, d, &L->cross_ln, 1);
        linear(h, n, &L->cq, q);
        for (size_t i = 0; i < (size_t)n * d; i++) q[i] = r16(q[i] * s4);
        const f16 *ck = s->cross_k + (size_t)il * T * d, *cv = s->cross_v + (size_t)il * T * d;
#pragma omp parallel
        {
            float *sc = (float *)malloc((size_t)(nctx > T ? nctx : T) * sizeof(float)); float o[256];
#pragma omp for schedule(static) collapse(2)
            for (int i = 0; i < n; i++)
                for (int hh = 0; hh < H; hh++) {
                    attend(q + (size_t)i * d + hh * dh, ck + hh * dh, cv + hh * dh, (size_t)d, T, dh, o, sc);
                    for (int c = 0; c < dh; c++) att[(size_t)i * d + hh * dh + c] = r16(o[c]);
                }
            free(sc);
        }
        linear(att, n, &L->co, y);
        for (size_t i = 0; i < (size_t)n * d; i++) x[i] += y[i];
        /* mlp */
        layer_norm(x, h, n, d, &L->mlp_ln, 1);
        linear(h, n, &L->fc1, ff);
        for (size_t i = 0; i < (size_t)n * 4 * d; i++) ff[i] = gelu_ggml(ff[i]);
        linear(ff, n, &L->fc2, y);
        for (size_t i = 0; i < (size_t)n * d; i++) x[i] += y[i];
    }
    layer_norm(x + (size_t)(n - 1) * d, h, 1, d, &m->d_ln, 1);
    gemm_f16w(h, (size_t)d, m->tok_emb, (size_t)d, logits_out, (size_t)hp->n_vocab, 1, hp->n_vocab, d, NULL);
    s->n_decoded += 1;
    free(x); free(h); free(q); free(kk); free(v); free(att); free(y); free(ff);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* whisper_full (Appendix A.5)                                                                  */
/* ------------------------------------------------------------------------------------------ */
void wo_default_params(wo_params *p) {
    memset(p, 0, sizeof *p);
    p->language = NULL; p->tdrz_enable = 0; p->no_context = 0; p->single_segment = 0; p->best_of = 5; p->beam_size = 0;
    p->temperature = 0.0f; p->temperature_inc = 0.2f; p->entropy_thold = 2.4f; p->logprob_thold = -1.0f;
    p->max_initial_ts = 1.0f; p->length_penalty = -1.0f; p->suppress_blank = 1; p->n_max_text_ctx = 16384;
    p->max_tokens = 0; p->n_threads = 16; p->keep_logits = 0;
}

static void seq_push(sequence_t *q, tokdata_t t) {
    if (q->n == q->cap) { q->cap = q->cap ? 2 * q->cap : 256; q->tokens = (tokdata_t *)realloc(q->tokens, (size_t)q->cap * sizeof(tokdata_t)); }
    q->tokens[q->n++] = t;
}
static void seq_copy(sequence_t *dst, const sequence_t *src) {
    tokdata_t *buf = dst->tokens; int cap = dst->cap;
    if (cap < src->n) { cap = src->n + 64; buf = (tokdata_t *)realloc(buf, (size_t)cap * sizeof(tokdata_t)); }
    memcpy(buf, src->tokens, (size_t)src->n * sizeof(tokdata_t));
    *dst = *src; dst->tokens = buf; dst->cap = cap;
}

static void process_logits(wo_state *s, const wo_params *P, decoder_t *dc, const float *raw, float temperature) {
    const wo_model *m = s->m; const int nv = m->hp.n_vocab;
    float *logits = dc->logits, *logprobs = dc->logprobs, *probs = dc->probs;
    const sequence_t *q = &dc->seq;
    const int is_initial = q->n == 0;
    memcpy(logits, raw, (size_t)nv * sizeof(float));
    if (temperature > 0.0f) for (int i = 0; i < nv; i++) logits[i] /= temperature;
    if (P->suppress_blank && is_initial) { logits[m->eot] = -INFINITY; if (m->blank >= 0) logits[m->blank] = -INFINITY; }
    logits[m->not_] = -INFINITY;
    logits[m->sot] = -INFINITY; logits[m->nosp] = -INFINITY;
    if (!P->tdrz_enable) logits[m->solm] = -INFINITY;
    logits[m->translate] = -INFINITY; logits[m->transcribe] = -INFINITY; logits[m->prev] = -INFINITY;
    for (int i = 0; i < N_LANG_TABLE; i++) { int t = m->sot + 1 + i; if (t < nv) logits[t] = -INFINITY; }
    {
        const int last_ts = q->n > 0 && q->tokens[q->n - 1].id >= m->beg;
        const int penult_ts = q->n < 2 || q->tokens[q->n - 2].id >= m->beg;
        if (last_ts) {
            if (penult_ts) for (int i = m->beg; i < nv; i++) logits[i] = -INFINITY;
            else for (int i = 0; i < m->eot; i++) logits[i] = -INFINITY;
        }
    }
    if (is_initial && P->max_initial_ts > 0.0f) {
        const float precision = (float)WO_CHUNK / m->hp.n_audio_ctx;
        const int tid0 = (int)roundf(P->max_initial_ts / precision);
        for (int i = m->beg + tid0 + 1; i < nv; i++) logits[i] = -INFINITY;
    }
    if (dc->has_ts) {
        const int tid0 = dc->seek_delta / 2;
        for (int i = m->beg; i < m->beg + tid0 && i < nv; i++) logits[i] = -INFINITY;
    }
    {
        float mx = -INFINITY; for (int i = 0; i < nv; i++) if (logits[i] > mx) mx = logits[i];
        float lse = 0.0f;
        for (int i = 0; i < nv; i++) if (logits[i] > -INFINITY) lse += expf(logits[i] - mx);
        lse = logf(lse) + mx;
        for (int i = 0; i < nv; i++) logprobs[i] = logits[i] > -INFINITY ? logits[i] - lse : -INFINITY;
    }
    {
        float ts_lp = -INFINITY;
        {
            float lse = 0.0f, mx = -INFINITY;
            for (int i = m->beg; i < nv; i++) if (logprobs[i] > mx) mx = logprobs[i];
            for (int i = m->beg; i < nv; i++) if (logprobs[i] > -INFINITY) lse += expf(logprobs[i] - mx);
            if (lse > 0.0f) ts_lp = logf(lse) + mx;
        }
        float mt = -INFINITY; for (int i = 0; i < m->beg; i++) if (logprobs[i] > mt) mt = logprobs[i];
        if (ts_lp > mt) for (int i = 0; i < m->beg; i++) { logits[i] = -INFINITY; logprobs[i] = -INFINITY; }
    }
    for (int i = 0; i < nv; i++) probs[i] = logits[i] == -INFINITY ? 0.0f : expf(logprobs[i]);
}

/* test probe: whisper_process_logits on a given token history (ids of the tokens sampled so far in this window),
 * decoder timestamp state {has_ts, seek_delta} and raw logits; writes the filtered logits (-inf = masked).
 * Used by tests/test_oracle_golden.py to hold the timestamp grammar against HuggingFace's independent
 * WhisperTimeStampLogitsProcessor (tests/golden/logits_filter.npz). */
int wo_probe_process_logits(wo_state *s, const wo_params *P, const int *ids, int n_ids, int has_ts, int seek_delta,
                            const float *raw, float temperature, float *logits_out) {
    decoder_t *dc = &s->dec[0];
    if (n_ids < 0) return -1;
    if (n_ids > dc->seq.cap) { dc->seq.cap = n_ids + 16; dc->seq.tokens = (tokdata_t *)realloc(dc->seq.tokens, (size_t)dc->seq.cap * sizeof(tokdata_t)); }
    dc->seq.n = n_ids;
    for (int i = 0; i < n_ids; i++) { memset(&dc->seq.tokens[i], 0, sizeof dc->seq.tokens[i]); dc->seq.tokens[i].id = ids[i]; }
    dc->has_ts = has_ts; dc->seek_delta = seek_delta;
    process_logits(s, P, dc, raw, temperature);
    memcpy(logits_out, dc->logits, (size_t)s->m->hp.n_vocab * sizeof(float));
    dc->seq.n = 0; dc->has_ts = 0;
    return 0;
}

/* libstdc++ std::discrete_distribution<>(probs) driven by std::mt19937 via generate_canonical<double,53> */
static int sample_discrete(decoder_t *dc, const float *probs, int n) {
    double sum = 0.0; for (int i = 0; i < n; i++) sum += (double)probs[i];
    uint32_t a = mt_next(dc), b = mt_next(dc);
    double r = ((double)a + (double)b * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    double c = 0.0;
    dc->n_draws++;
    for (int i = 0; i < n - 1; i++) {
        const double lo = c;
        c += (double)probs[i] / sum;
        if (!(c < r)) {      /* distance of the uniform to the nearest boundary of the winner's interval: how much the cumulative */
            const double mg = (r - lo) < (c - r) ? (r - lo) : (c - r);      /* probabilities may move before the draw changes */
            if (mg < dc->min_margin) dc->min_margin = mg;
            return i;
        }
    }
    if (1.0 - r < dc->min_margin) dc->min_margin = 1.0 - r;
    return n - 1;
}

/* test probe: `count` draws of std::discrete_distribution<>(probs) from a std::mt19937 seeded with `seed`, as restated above;
 * tests/test_oracle_full.py compares them with the real libstdc++ classes (a C++ snippet compiled at test time). */
int wo_probe_sample(wo_state *s, uint32_t seed, const float *probs, int n, int count, int *out) {
    decoder_t *dc = &s->dec[0];
    mt_seed(dc, seed);
    for (int i = 0; i < count; i++) out[i] = sample_discrete(dc, probs, n);
    mt_seed(dc, 0);
    return 0;
}

double wo_min_sample_margin(const wo_state *s) {
    double m = 1.0;
    for (int j = 0; j < WO_MAX_DECODERS; j++) if (s->dec[j].n_draws > 0 && s->dec[j].min_margin < m) m = s->dec[j].min_margin;
    return m;
}
long wo_n_draws(const wo_state *s) { long n = 0; for (int j = 0; j < WO_MAX_DECODERS; j++) n += s->dec[j].n_draws; return n; }

static tokdata_t sample_token(wo_state *s, decoder_t *dc, int best) {
    const wo_model *m = s->m; const int nv = m->hp.n_vocab;
    tokdata_t r = {0, 0, 0.f, 0.f, 0.f, 0.f};
    {
        double sum_ts = 0.0, max_ts = 0.0;
        for (int i = m->beg; i < nv; i++) {
            sum_ts += dc->probs[i];
            if (max_ts < dc->probs[i]) { max_ts = dc->probs[i]; r.tid = i; }
        }
        r.pt = (float)(max_ts / (sum_ts + 1e-10)); r.ptsum = (float)sum_ts;
    }
    if (best) {
        for (int i = 0; i < nv; i++) if (r.p < dc->probs[i]) { r.id = i; r.p = dc->probs[i]; r.plog = dc->logprobs[i]; }
    } else {
        r.id = sample_discrete(dc, dc->probs, nv); r.p = dc->probs[r.id]; r.plog = dc->logprobs[r.id];
    }
    if (r.id >= m->beg) { r.tid = r.id; r.pt = r.p; }
    return r;
}

static void sequence_score(const wo_params *P, sequence_t *q) {
    if (q->result_len == 0) return;
    double result = 0.0;
    for (int i = 0; i < q->result_len; i++) result += q->tokens[i].plog;
    q->sum_logprobs = result; q->avg_logprobs = result / q->result_len;
    double penalty = q->result_len;
    if (P->length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, P->length_penalty);
    q->score = result / penalty;
    {
        const int n = 32; int cnt = 0; double entropy = 0.0;
        int i0 = q->result_len - n > 0 ? q->result_len - n : 0;
        int ids[32], counts[32], nu = 0;
        for (int i = i0; i < q->result_len; i++) {
            int id = q->tokens[i].id, j;
            for (j = 0; j < nu; j++) if (ids[j] == id) { counts[j]++; break; }
            if (j == nu) { ids[nu] = id; counts[nu] = 1; nu++; }
            cnt++;
        }
        /* std::map iterates in key order; summation order follows it */
        for (int a = 0; a < nu; a++) for (int b = a + 1; b < nu; b++) if (ids[b] < ids[a]) { int t = ids[a]; ids[a] = ids[b]; ids[b] = t; t = counts[a]; counts[a] = counts[b]; counts[b] = t; }
        for (int j = 0; j < nu; j++) { double p = counts[j] / (double)cnt; entropy -= p * log(p); }
        q->entropy = entropy;
    }
}

/* whisper_full's per-token decoder bookkeeping for the token `id` sampled at step i of a window: the timestamp
 * / seek_delta / result_len update, the completion rule (EOT, max_tokens, a timestamp within 1 s of the end of the audio) and the
 * two failure rules (timestamps going backwards; the token cap reached without a usable timestamp).  Used by wo_full and by the
 * wo_probe_bookkeeping test probe. */
static void token_bookkeeping(const wo_model *m, const wo_params *P, decoder_t *dc, int id, int i, int seek, int seek_end, int n_max) {
    if (id > m->beg) {
        const int sd_new = 2 * (id - m->beg);
        if (dc->has_ts && dc->seek_delta > sd_new && dc->seq.result_len < i) { dc->failed = 1; return; }
        dc->seek_delta = sd_new; dc->seq.result_len = i + 1; dc->has_ts = 1;
    }
    if (id == m->eot || (P->max_tokens > 0 && i >= P->max_tokens) || (dc->has_ts && seek + dc->seek_delta + 100 >= seek_end)) {
        if (dc->seq.result_len == 0) {
            if (seek + dc->seek_delta + 100 >= seek_end) dc->seq.result_len = i + 1;
            else { dc->failed = 1; return; }
        }
        if (P->single_segment) { dc->seq.result_len = i + 1; dc->seek_delta = 100 * WO_CHUNK; }
        dc->completed = 1; return;
    }
    if (i == n_max - 1 && (dc->seq.result_len == 0 || dc->seek_delta < 100 * WO_CHUNK / 2)) { dc->failed = 1; return; }
}

/* test probe: token_bookkeeping over a given list of sampled ids (decoder state as at the start of a window);
 * out = {failed, completed, result_len, seek_delta, has_ts, steps consumed} */
int wo_probe_bookkeeping(wo_state *s, const wo_params *P, const int *ids, int n, int seek, int seek_end, int n_max, int *out) {
    decoder_t tmp; memset(&tmp, 0, sizeof tmp);
    tmp.seek_delta = 100 * WO_CHUNK;
    int i = 0;
    for (; i < n && i < n_max; i++) {
        token_bookkeeping(s->m, P, &tmp, ids[i], i, seek, seek_end, n_max);
        if (tmp.failed || tmp.completed) { i++; break; }
    }
    out[0] = tmp.failed; out[1] = tmp.completed; out[2] = tmp.seq.result_len; out[3] = tmp.seek_delta; out[4] = tmp.has_ts; out[5] = i;
    return 0;
}

/* test probe: sequence_score (whisper_sequence_score + the entropy of the last 32 tokens) on given ids / log-probs;
 * out = {sum_logprobs, avg_logprobs, entropy, score} */
int wo_probe_score(wo_state *s, const wo_params *P, const int *ids, const float *plogs, int n, int result_len, double *out) {
    (void)s;
    sequence_t q; memset(&q, 0, sizeof q);
    q.tokens = (tokdata_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(tokdata_t));
    for (int i = 0; i < n; i++) { q.tokens[i].id = ids[i]; q.tokens[i].plog = plogs[i]; }
    q.n = n; q.result_len = result_len;
    sequence_score(P, &q);
    out[0] = q.sum_logprobs; out[1] = q.avg_logprobs; out[2] = q.entropy; out[3] = q.score;
    free(q.tokens);
    return 0;
}

static void kv_seq_copy(wo_state *s, int from, int to) {
    if (from == to) return;
    const wo_hparams *hp = &s->m->hp;
    size_t per = (size_t)hp->n_text_layer * hp->n_text_ctx * hp->n_text_state;
    memcpy(s->self_k + (size_t)to * per, s->self_k + (size_t)from * per, per * sizeof(f16));
    memcpy(s->self_v + (size_t)to * per, s->self_v + (size_t)from * per, per * sizeof(f16));
}

static void push_segment(wo_state *s, int64_t t0, int64_t t1, const char *text, size_t len, int turn) {
    if (s->n_segs == s->cap_segs) { s->cap_segs = s->cap_segs ? 2 * s->cap_segs : 16; s->segs = (segment_t *)realloc(s->segs, (size_t)s->cap_segs * sizeof(segment_t)); }
    segment_t *g = &s->segs[s->n_segs++];
    g->t0 = t0; g->t1 = t1; g->speaker_turn_next = turn;
    g->text = (char *)malloc(len + 1); memcpy(g->text, text, len); g->text[len] = 0;
}
static void push_result_token(wo_state *s, const tokdata_t *t) {
    if (s->n_res == s->cap_res) {
        s->cap_res = s->cap_res ? 2 * s->cap_res : 512;
        s->res_tok = (int *)realloc(s->res_tok, (size_t)s->cap_res * sizeof(int));
        s->res_p = (float *)realloc(s->res_p, (size_t)s->cap_res * sizeof(float));
        s->res_plog = (float *)realloc(s->res_plog, (size_t)s->cap_res * sizeof(float));
    }
    s->res_tok[s->n_res] = t->id; s->res_p[s->n_res] = t->p; s->res_plog[s->n_res] = t->plog; s->n_res++;
}

typedef struct { int decoder_idx, seek_delta, has_ts; sequence_t seq; } beam_cand_t;

/* One beam-search step's candidate assignment (whisper.cpp whisper_full_with_state, BEAM_SEARCH branch behind "update each decoder"):
 * the candidates of all live decoders are sorted by sum_logprobs_all, descending (stable here; std::sort there), and handed to the
 * live decoders in order; from the second sampled token on (i > 0) a decoder skips the candidates that follow its own and carry the
 * same token sequence; the candidate cursor wraps to 0 when it runs past the end.  pick[j] = index into the SORTED array of the
 * candidate decoder j continues with, -1 for a decoder that has completed / failed.  Factored out for wo_full and the
 * wo_probe_beam_assign test probe. */
static void beam_assign(beam_cand_t *cands, int n_cands, int n_cur, const int *live, int i, int *pick) {
    for (int a = 1; a < n_cands; a++) {
        beam_cand_t tmp = cands[a]; int b = a - 1;
        while (b >= 0 && cands[b].seq.sum_logprobs_all < tmp.seq.sum_logprobs_all) { cands[b + 1] = cands[b]; b--; }
        cands[b + 1] = tmp;
    }
    int cur_c = 0;
    for (int j = 0; j < n_cur; j++) {
        pick[j] = -1;
        if (!live[j]) continue;
        if (cur_c >= n_cands) cur_c = 0;
        const beam_cand_t *c = &cands[cur_c];
        pick[j] = cur_c++;
        while (n_cands > cur_c && i > 0 && cands[cur_c].seq.n == c->seq.n) {
            int eq = 1; for (int a = 0; a < c->seq.n; a++) if (cands[cur_c].seq.tokens[a].id != c->seq.tokens[a].id) { eq = 0; break; }
            if (!eq) break;
            ++cur_c;
        }
    }
}

/* test probe: beam_assign over n_cands candidates given as token-id rows (ids[c * max_len ..], len[c] tokens each), their
 * sum_logprobs_all and the decoder each one came from; out[j] = ORIGINAL index of the candidate decoder j continues with (-1: not live) */
int wo_probe_beam_assign(const int *ids, const int *len, int max_len, const double *sums, const int *decoder_idx, int n_cands,
                         const int *live, int n_cur, int i, int *out) {
    if (n_cands < 0 || n_cands > WO_MAX_DECODERS * WO_MAX_DECODERS || n_cur < 0 || n_cur > WO_MAX_DECODERS) return -1;
    beam_cand_t *cands = (beam_cand_t *)calloc((size_t)(n_cands > 0 ? n_cands : 1), sizeof(beam_cand_t));
    for (int c = 0; c < n_cands; c++) {
        cands[c].decoder_idx = decoder_idx[c];
        cands[c].seek_delta = c;      /* carries the original index through the sort */
        for (int a = 0; a < len[c]; a++) { tokdata_t t = {ids[(size_t)c * max_len + a], 0, 0.f, 0.f, 0.f, 0.f}; seq_push(&cands[c].seq, t); }
        cands[c].seq.sum_logprobs_all = sums[c];
    }
    int pick[WO_MAX_DECODERS];
    beam_assign(cands, n_cands, n_cur, live, i, pick);
    for (int j = 0; j < n_cur; j++) out[j] = pick[j] >= 0 ? cands[pick[j]].seek_delta : -1;
    for (int c = 0; c < n_cands; c++) free(cands[c].seq.tokens);
    free(cands);
    return 0;
}

int wo_full(wo_state *s, const float *pcm, size_t n_samples, const wo_params *P) {
    wo_model *m = s->m; const wo_hparams *hp = &m->hp; const int nv = hp->n_vocab;
    clear_segments(s);
    s->n_res = 0; s->n_fallbacks = 0; s->n_decoded = 0; s->n_windows = 0; s->n_kept = 0;
    for (int j = 0; j < WO_MAX_DECODERS; j++) { s->dec[j].min_margin = 1.0; s->dec[j].n_draws = 0; }
    if (P->n_threads > 0) {
#ifdef _OPENMP
        int hw = omp_get_num_procs(); omp_set_num_threads(P->n_threads < hw ? P->n_threads : hw);
#endif
    }
    float *mel = NULL; int n_len = 0, n_len_org = 0;
    wo_log_mel(m, pcm, n_samples, &mel, &n_len, &n_len_org);
    const int seek_start = 0, seek_end = n_len_org;
    if (seek_end < seek_start + 100) { free(mel); return 0; }

    float temps[16]; int n_temps = 0;
    if (P->temperature_inc > 0.0f) { for (float t = P->temperature; t < 1.0f + 1e-6f && n_temps < 16; t += P->temperature_inc) temps[n_temps++] = t; }
    else temps[n_temps++] = P->temperature;

    const int beam = P->beam_size > 1;
    int n_decoders = beam ? (P->best_of > P->beam_size ? P->best_of : P->beam_size) : P->best_of;
    if (n_decoders < 1) n_decoders = 1;
    if (n_decoders > WO_MAX_DECODERS) { set_err("too many decoders"); free(mel); return -4; }

    if (P->no_context) s->n_prompt_past = 0;

    int prompt_init[4], n_init = 0;
    prompt_init[n_init++] = m->sot;
    if (m->multilingual) {
        int lid = wo_lang_id(P->language ? P->language : "en");
        if (lid < 0) { set_err("unknown language '%s'", P->language); free(mel); return -2; }
        prompt_init[n_init++] = m->sot + 1 + lid;
        prompt_init[n_init++] = m->transcribe;
    }
    int seek = seek_start;
    int *prompt = (int *)malloc((size_t)(hp->n_text_ctx + 8) * sizeof(int)); int n_prompt = 0;
    float *raw = (float *)malloc((size_t)nv * sizeof(float));
    beam_cand_t *cands = beam ? (beam_cand_t *)calloc((size_t)WO_MAX_DECODERS * WO_MAX_DECODERS, sizeof(beam_cand_t)) : NULL;
    int rc = 0;

    while (1) {
        if (seek + 100 >= seek_end) break;
        wo_encode(s, mel, n_len, seek);
        s->n_windows++;
        if (seek > seek_start && seek + 500 >= seek_end) s->n_prompt_past = 0;
        int best_decoder_id = 0;
        for (int it = 0; it < n_temps; it++) {
            const float t_cur = temps[it];
            int n_cur = 1;
            if (!beam) { if (t_cur > 0.0f) n_cur = P->best_of; }
            else { n_cur = t_cur > 0.0f ? P->best_of : P->beam_size; }
            if (n_cur < 1) n_cur = 1;
            if (it > 0) s->n_fallbacks++;
            for (int j = 0; j < n_cur; j++) {
                decoder_t *dc = &s->dec[j];
                dc->seq.n = 0; dc->seq.result_len = 0; dc->seq.sum_logprobs_all = 0.0; dc->seq.sum_logprobs = -INFINITY;
                dc->seq.avg_logprobs = -INFINITY; dc->seq.entropy = 0.0; dc->seq.score = -INFINITY;
                dc->seek_delta = 100 * WO_CHUNK; dc->failed = 0; dc->completed = 0; dc->has_ts = 0;
            }
            n_prompt = 0;
            if (s->n_prompt_past > 0 && t_cur < 0.5f && P->n_max_text_ctx > 0) {
                int n_take = P->n_max_text_ctx < hp->n_text_ctx / 2 ? P->n_max_text_ctx : hp->n_text_ctx / 2;
                if (n_take > s->n_prompt_past) n_take = s->n_prompt_past;
                prompt[n_prompt++] = m->prev;
                memcpy(prompt + n_prompt, s->prompt_past + s->n_prompt_past - n_take, (size_t)n_take * sizeof(int)); n_prompt += n_take;
            }
            memcpy(prompt + n_prompt, prompt_init, (size_t)n_init * sizeof(int)); n_prompt += n_init;
            if ((rc = wo_decode(s, 0, prompt, n_prompt, 0, raw)) != 0) { rc = -7; goto done; }
            if (P->keep_logits && it == 0) {
                if (s->n_kept == s->cap_kept) { s->cap_kept = s->cap_kept ? 2 * s->cap_kept : 64; s->kept = (float *)realloc(s->kept, (size_t)s->cap_kept * nv * sizeof(float)); }
                memcpy(s->kept + (size_t)s->n_kept++ * nv, raw, (size_t)nv * sizeof(float));
            }
            process_logits(s, P, &s->dec[0], raw, t_cur);
            for (int j = 1; j < n_cur; j++) {
                decoder_t *dc = &s->dec[j];
                kv_seq_copy(s, 0, j);
                memcpy(dc->probs, s->dec[0].probs, (size_t)nv * sizeof(float));
                memcpy(dc->logits, s->dec[0].logits, (size_t)nv * sizeof(float));
                memcpy(dc->logprobs, s->dec[0].logprobs, (size_t)nv * sizeof(float));
            }
            const int n_max = hp->n_text_ctx / 2 - 4;
            for (int i = 0; i < n_max; i++) {
                int n_cands = 0;
                for (int j = 0; j < n_cur; j++) {
                    decoder_t *dc = &s->dec[j];
                    if (dc->completed || dc->failed) continue;
                    if (!beam) {
                        tokdata_t t = sample_token(s, dc, t_cur < 1e-6f);
                        seq_push(&dc->seq, t); dc->seq.sum_logprobs_all += t.plog;
                    } else {
                        /* whisper_sample_token_topk: top-k by (logprob desc, id asc); tid/pt as in sample_token */
                        int k = P->beam_size;
                        int top[WO_MAX_DECODERS]; int nt = 0;
                        for (int a = 0; a < k; a++) {
                            int bi = -1; float bv = -INFINITY;
                            for (int v = 0; v < nv; v++) {
                                int used = 0; for (int b = 0; b < nt; b++) if (top[b] == v) used = 1;
                                if (used) continue;
                                if (bi < 0 || dc->logprobs[v] > bv) { bi = v; bv = dc->logprobs[v]; }
                            }
                            top[nt++] = bi;
                        }
                        double sum_ts = 0.0, max_ts = 0.0; int tid = 0;
                        for (int v = m->beg; v < nv; v++) { sum_ts += dc->probs[v]; if (max_ts < dc->probs[v]) { max_ts = dc->probs[v]; tid = v; } }
                        for (int a = 0; a < nt; a++) {
                            beam_cand_t *c = &cands[n_cands++];
                            c->decoder_idx = j; c->seek_delta = dc->seek_delta; c->has_ts = dc->has_ts;
                            seq_copy(&c->seq, &dc->seq);
                            tokdata_t t = {top[a], tid, dc->probs[top[a]], dc->logprobs[top[a]], (float)(max_ts / (sum_ts + 1e-10)), (float)sum_ts};
                            if (t.id >= m->beg) { t.tid = t.id; t.pt = t.p; }
                            seq_push(&c->seq, t); c->seq.sum_logprobs_all += t.plog;
                        }
                    }
                }
                if (beam) {
                    int live[WO_MAX_DECODERS], pick[WO_MAX_DECODERS], src[WO_MAX_DECODERS];
                    for (int j = 0; j < n_cur; j++) live[j] = !(s->dec[j].completed || s->dec[j].failed);
                    beam_assign(cands, n_cands, n_cur, live, i, pick);
                    for (int j = 0; j < n_cur; j++) {
                        decoder_t *dc = &s->dec[j]; src[j] = -1;
                        if (pick[j] < 0) continue;
                        const beam_cand_t *c = &cands[pick[j]];
                        dc->seek_delta = c->seek_delta; dc->has_ts = c->has_ts; seq_copy(&dc->seq, &c->seq);
                        src[j] = c->decoder_idx;
                    }
                    /* KV shuffle through temporaries (seq ids MAX+j in whisper.cpp): 2-phase copy */
                    {
                        size_t per = (size_t)hp->n_text_layer * hp->n_text_ctx * hp->n_text_state;
                        f16 *tk = (f16 *)malloc((size_t)n_cur * per * sizeof(f16)), *tv = (f16 *)malloc((size_t)n_cur * per * sizeof(f16));
                        for (int j = 0; j < n_cur; j++) if (src[j] >= 0) { memcpy(tk + (size_t)j * per, s->self_k + (size_t)src[j] * per, per * sizeof(f16)); memcpy(tv + (size_t)j * per, s->self_v + (size_t)src[j] * per, per * sizeof(f16)); }
                        for (int j = 0; j < n_cur; j++) if (src[j] >= 0) { memcpy(s->self_k + (size_t)j * per, tk + (size_t)j * per, per * sizeof(f16)); memcpy(s->self_v + (size_t)j * per, tv + (size_t)j * per, per * sizeof(f16)); }
                        free(tk); free(tv);
                    }
                }
                for (int j = 0; j < n_cur; j++) {
                    decoder_t *dc = &s->dec[j];
                    if (dc->completed || dc->failed) continue;
                    token_bookkeeping(m, P, dc, dc->seq.tokens[dc->seq.n - 1].id, i, seek, seek_end, n_max);
                }
                {
                    int all = 1;
                    for (int j = 0; j < n_cur; j++) if (!(s->dec[j].completed || s->dec[j].failed)) all = 0;
                    if (all) break;
                }
                const int n_past = n_prompt + i;
                for (int j = 0; j < n_cur; j++) {
                    decoder_t *dc = &s->dec[j];
                    if (dc->failed || dc->completed) continue;
                    int tok = dc->seq.tokens[dc->seq.n - 1].id;
                    if ((rc = wo_decode(s, j, &tok, 1, n_past, raw)) != 0) { rc = -8; goto done; }
                    if (P->keep_logits && it == 0 && j == 0) {
                        if (s->n_kept == s->cap_kept) { s->cap_kept = s->cap_kept ? 2 * s->cap_kept : 64; s->kept = (float *)realloc(s->kept, (size_t)s->cap_kept * nv * sizeof(float)); }
                        memcpy(s->kept + (size_t)s->n_kept++ * nv, raw, (size_t)nv * sizeof(float));
                    }
                    process_logits(s, P, dc, raw, t_cur);
                }
            }
            {
                double best_score = -INFINITY;
                for (int j = 0; j < n_cur; j++) {
                    decoder_t *dc = &s->dec[j];
                    if (dc->failed) continue;
                    dc->seq.n = dc->seq.result_len < dc->seq.n ? dc->seq.result_len : dc->seq.n;
                    sequence_score(P, &dc->seq);
                    if (dc->seq.entropy < P->entropy_thold) { dc->failed = 1; continue; }
                    if (best_score < dc->seq.score) { best_score = dc->seq.score; best_decoder_id = j; }
                }
            }
            if (getenv("WO_TRACE")) for (int j = 0; j < n_cur; j++) fprintf(stderr, "[wo] seek %d t=%.1f dec %d: n=%d result_len=%d failed=%d completed=%d avg_lp=%.4f entropy=%.3f seek_delta=%d\n", seek, t_cur, j, s->dec[j].seq.n, s->dec[j].seq.result_len, s->dec[j].failed, s->dec[j].completed, s->dec[j].seq.avg_logprobs, s->dec[j].seq.entropy, s->dec[j].seek_delta);
            int success = 1;
            if (it != n_temps - 1) {
                const decoder_t *dc = &s->dec[best_decoder_id];
                if (dc->failed || dc->seq.avg_logprobs < P->logprob_thold) success = 0;
            }
            if (success) break;
        }
        {
            const decoder_t *bd = &s->dec[best_decoder_id];
            const int seek_delta = bd->seek_delta, result_len = bd->seq.result_len;
            const tokdata_t *tc = bd->seq.tokens; const int ntc = bd->seq.n;
            /* update prompt_past */
            {
                int keep_n = 0; int *keep = NULL;
                if (prompt[0] == m->prev) { keep_n = n_prompt - 1 - n_init; keep = prompt + 1; }
                int need = keep_n + result_len;
                if (need > s->cap_prompt_past) { s->cap_prompt_past = need + 256; s->prompt_past = (int *)realloc(s->prompt_past, (size_t)s->cap_prompt_past * sizeof(int)); }
                if (keep_n > 0) memmove(s->prompt_past, keep, (size_t)keep_n * sizeof(int));
                s->n_prompt_past = keep_n;
                for (int i = 0; i < result_len && i < ntc; i++) s->prompt_past[s->n_prompt_past++] = tc[i].id;
            }
            if (ntc > 0) {
                for (int i = 0; i < ntc; i++) push_result_token(s, &tc[i]);
                int64_t t0 = seek + 2 * (tc[0].tid - m->beg);
                char *text = (char *)malloc(16); size_t tl = 0, tcap = 16; int turn = 0;
                for (int i = 0; i < ntc; i++) {
                    if (tc[i].id < m->eot) {
                        int l = m->tok[tc[i].id].len;
                        if (tl + (size_t)l + 1 > tcap) { tcap = 2 * (tl + (size_t)l + 1); text = (char *)realloc(text, tcap); }
                        memcpy(text + tl, m->tok[tc[i].id].s, (size_t)l); tl += (size_t)l;
                    }
                    if (P->tdrz_enable && tc[i].id == m->solm) turn = 1;
                    if (tc[i].id > m->beg && !P->single_segment) {
                        const int64_t t1 = seek + 2 * (tc[i].tid - m->beg);
                        if (tl > 0) push_segment(s, t0, t1, text, tl, turn);
                        tl = 0;
                        while (i < ntc && tc[i].id > m->beg) i++;
                        i--;
                        t0 = t1; turn = 0;
                    }
                }
                if (tl > 0) push_segment(s, t0, (int64_t)seek + seek_delta, text, tl, turn);
                free(text);
            }
            seek += seek_delta;
        }
    }
done:
    if (cands) { for (int i = 0; i < WO_MAX_DECODERS * WO_MAX_DECODERS; i++) free(cands[i].seq.tokens); free(cands); }
    free(prompt); free(raw); free(mel);
    return rc;
}

int wo_n_segments(const wo_state *s) { return s->n_segs; }
const char *wo_segment_text(const wo_state *s, int i) { return (i >= 0 && i < s->n_segs) ? s->segs[i].text : NULL; }
int64_t wo_segment_t0(const wo_state *s, int i) { return (i >= 0 && i < s->n_segs) ? s->segs[i].t0 : -1; }
int64_t wo_segment_t1(const wo_state *s, int i) { return (i >= 0 && i < s->n_segs) ? s->segs[i].t1 : -1; }
int wo_segment_speaker_turn_next(const wo_state *s, int i) { return (i >= 0 && i < s->n_segs) ? s->segs[i].speaker_turn_next : 0; }
int wo_n_result_tokens(const wo_state *s) { return s->n_res; }
int wo_result_token(const wo_state *s, int i, float *p, float *plog) {
    if (i < 0 || i >= s->n_res) return -1;
    if (p) *p = s->res_p[i];
    if (plog) *plog = s->res_plog[i];
    return s->res_tok[i];
}
int wo_n_fallbacks(const wo_state *s) { return s->n_fallbacks; }
int wo_n_decoded(const wo_state *s) { return s->n_decoded; }
int wo_n_windows(const wo_state *s) { return s->n_windows; }
int wo_n_kept_logits(const wo_state *s) { return s->n_kept; }
const float *wo_kept_logits(const wo_state *s, int step) { return (step >= 0 && step < s->n_kept) ? s->kept + (size_t)step * s->m->hp.n_vocab : NULL; }
