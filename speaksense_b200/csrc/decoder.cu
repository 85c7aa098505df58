// decoder.cu - batch-1 autoregressive text decoder step: HBM-bandwidth-bound vectorised kernels
// over a persistent KV cache (BASELINE.json north_star stage 3).
//
// One step = embed -> 32 x [LN+QKV GEMV (KV append) -> self-attn partials -> out-proj GEMV (+combine,
// +residual) -> LN+crossQ GEMV -> cross-attn partials over 1500 keys -> cross-out GEMV (+combine,
// +residual) -> LN+FC1 GEMV+GELU -> FC2 GEMV (+residual)] -> LN+LM-head GEMV -> logits filter +
// greedy sample + decoder-state update, all on the device.  The step reads its token / position
// from a device-resident control block (DecCtl), so the whole step is one static CUDA graph that
// the host replays without a round trip per token.
//
// Arithmetic follows the oracle (oracle/whisper_oracle.c wo_decode / process_logits): f16 weights,
// activations rounded to f16 in front of every mat-vec, f32 accumulation.
#include "kernels.h"

namespace ss {

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ uint4 ldg_stream(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 h2f(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2 *>(&u);
    return __half22float2(h);
}
__device__ __forceinline__ float dot8(const uint4 &w, const float4 &a, const float4 &b, float acc) {
    float2 f;
    f = h2f(w.x); acc = fmaf(f.x, a.x, acc); acc = fmaf(f.y, a.y, acc);
    f = h2f(w.y); acc = fmaf(f.x, a.z, acc); acc = fmaf(f.y, a.w, acc);
    f = h2f(w.z); acc = fmaf(f.x, b.x, acc); acc = fmaf(f.y, b.y, acc);
    f = h2f(w.w); acc = fmaf(f.x, b.z, acc); acc = fmaf(f.y, b.w, acc);
    return acc;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum over 256 threads; scratch >= 8 floats; all threads get the result
__device__ __forceinline__ float block_sum_256(float v, float *scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += scratch[i];
    return t;
}
__device__ __forceinline__ float gelu16(float x) {
    const float xh = r16(x);
    return r16(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh))));
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_embed_kernel(const __half *__restrict__ tok_emb, const float *__restrict__ pos_emb,
                                                         float *__restrict__ x, int d, const DecCtl *__restrict__ ctl) {
    if (ctl->done) return;
    const int token = ctl->token, pos = ctl->pos;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < d; c += gridDim.x * blockDim.x)
        x[c] = __half2float(tok_emb[(size_t)token * d + c]) + pos_emb[(size_t)pos * d + c];
}

// ------------------------------------------------------------------------------------------------
// GEMV: v[n] = W[n][:] . xs + bias[n], xs staged in shared memory by one of three prologues.
// ------------------------------------------------------------------------------------------------
extern __shared__ float gemv_smem[];

__global__ void __launch_bounds__(256) dec_gemv_kernel(GemvArgs a) {
    const DecCtl *ctl = a.ctl;
    if (ctl->done) return;
    if (a.logits_gate && (ctl->pos - ctl->pos0) < ctl->n_prompt - 1) return;   // prompt token: logits unused
    float *xs = gemv_smem;               // [K]
    __shared__ float red[8];
    const int K = a.K, tid = threadIdx.x;

    if (a.pro == PRO_PLAIN) {
        for (int i = tid; i < K; i += 256) xs[i] = a.xin[i];
    } else if (a.pro == PRO_LN) {
        // ggml_norm: mean, then variance of centred values, eps 1e-5; affine; rounded to f16 for the mat-vec
        float s = 0.f;
        for (int i = tid; i < K; i += 256) { float v = a.xin[i]; xs[i] = v; s += v; }
        const float mean = block_sum_256(s, red) / K;
        float s2 = 0.f;
        for (int i = tid; i < K; i += 256) { float v = xs[i] - mean; xs[i] = v; s2 += v * v; }
        const float var = block_sum_256(s2, red) / K;
        const float scale = rsqrtf(var + 1e-5f);
        for (int i = tid; i < K; i += 256) xs[i] = r16(xs[i] * scale * a.lnw[i] + a.lnb[i]);
    } else {   // PRO_ATTN: combine split-softmax partials [H][n_split][66] = {m, l, o[64]}
        const int ns = a.n_split;
        for (int n = tid; n < K; n += 256) {
            const int h = n >> 6, c = n & 63;
            const float *p = a.part + (size_t)h * ns * 66;
            float M = -INFINITY;
            for (int s = 0; s < ns; s++) M = fmaxf(M, p[s * 66]);
            float L = 0.f, o = 0.f;
            for (int s = 0; s < ns; s++) {
                const float ms = p[s * 66];
                if (ms == -INFINITY) continue;
                const float e = __expf(ms - M);
                L += p[s * 66 + 1] * e; o += p[s * 66 + 2 + c] * e;
            }
            xs[n] = r16(o / L);
        }
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    const int nchunk = K >> 3;
    const float4 *xs4 = reinterpret_cast<const float4 *>(xs);
    for (int row = blockIdx.x * 8 + warp; row < a.N; row += gridDim.x * 8) {
        const uint4 *wr = reinterpret_cast<const uint4 *>(a.W + (size_t)row * K);
        float acc = 0.f;
        for (int c0 = 0; c0 < nchunk; c0 += 32 * 5) {
            uint4 w[5];
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int c = c0 + u * 32 + lane;
                w[u] = c < nchunk ? ldg_stream(wr + c) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 5; u++) {
                const int c = c0 + u * 32 + lane;
                if (c < nchunk) acc = dot8(w[u], xs4[2 * c], xs4[2 * c + 1], acc);
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            float v = acc + (a.bias ? a.bias[row] : 0.f);
            switch (a.epi) {
                case EPI_STORE: a.out[row] = v; break;
                case EPI_RESID: a.out[row] += v; break;
                case EPI_GELU: a.out[row] = gelu16(v); break;
                case EPI_QSCALE: a.out[row] = r16(v * a.s4); break;
                case EPI_QKV: {
                    const int d = a.N / 3, pos = ctl->pos;
                    if (row < d) a.out[row] = r16(v * a.s4);
                    else if (row < 2 * d) { const int n = row - d; a.kcache[((size_t)(n >> 6) * a.ctx + pos) * 64 + (n & 63)] = __float2half_rn(v * a.s4); }
                    else { const int n = row - 2 * d; a.vcache[((size_t)(n >> 6) * a.ctx + pos) * 64 + (n & 63)] = __float2half_rn(v); }
                } break;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Attention partials for one query vector: grid (n_split, H), 8 lanes per key row (64 x f16 = 128 B).
// K/V layout: [head][ctx][64] f16.  Emits {m, l, o[64]} per (head, split).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dec_attn_kernel(AttnArgs a) {
    const DecCtl *ctl = a.ctl;
    if (ctl->done) return;
    __shared__ float sc[512];
    __shared__ float red[8][64];
    __shared__ float red1[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane >> 3, l8 = lane & 7;
    const int h = blockIdx.y, s = blockIdx.x, ns = gridDim.x;
    const int n_keys = a.n_keys >= 0 ? a.n_keys : ctl->pos + 1;
    const int per = (n_keys + ns - 1) / ns;
    const int j0 = s * per, j1 = min(n_keys, j0 + per);
    float *out = a.part + ((size_t)h * ns + s) * 66;
    if (j0 >= j1) {
        if (tid == 0) { out[0] = -INFINITY; out[1] = 0.f; }
        if (tid < 64) out[2 + tid] = 0.f;
        return;
    }
    const float4 qa = *reinterpret_cast<const float4 *>(a.q + h * 64 + l8 * 8);
    const float4 qb = *reinterpret_cast<const float4 *>(a.q + h * 64 + l8 * 8 + 4);
    const __half *Kh = a.K + (size_t)h * a.ctx * 64, *Vh = a.V + (size_t)h * a.ctx * 64;
    const int n = j1 - j0;
    // scores
    float lmax = -INFINITY;
    for (int jb = 0; jb < n; jb += 32 * 4) {
        uint4 kv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            kv[u] = j < n ? ldg_stream(Kh + (size_t)(j0 + j) * 64 + l8 * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            float d = dot8(kv[u], qa, qb, 0.f);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            if (j < n) { if (l8 == 0) sc[j] = d; lmax = fmaxf(lmax, d); }
        }
    }
    lmax = warp_max(lmax);
    if (lane == 0) red1[warp] = lmax;
    __syncthreads();
    float m = red1[0];
#pragma unroll
    for (int i = 1; i < 8; i++) m = fmaxf(m, red1[i]);
    float lsum = 0.f;
    for (int j = tid; j < n; j += 256) { const float e = __expf(sc[j] - m); sc[j] = e; lsum += e; }
    const float l = block_sum_256(lsum, red1);   // includes the barriers that publish sc[]
    // P.V
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int jb = 0; jb < n; jb += 32 * 4) {
        uint4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            vv[u] = j < n ? ldg_stream(Vh + (size_t)(j0 + j) * 64 + l8 * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) {
                const float p = sc[j];
                float2 f;
                f = h2f(vv[u].x); acc[0] = fmaf(p, f.x, acc[0]); acc[1] = fmaf(p, f.y, acc[1]);
                f = h2f(vv[u].y); acc[2] = fmaf(p, f.x, acc[2]); acc[3] = fmaf(p, f.y, acc[3]);
                f = h2f(vv[u].z); acc[4] = fmaf(p, f.x, acc[4]); acc[5] = fmaf(p, f.y, acc[5]);
                f = h2f(vv[u].w); acc[6] = fmaf(p, f.x, acc[6]); acc[7] = fmaf(p, f.y, acc[7]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
        acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) red[warp][l8 * 8 + i] = acc[i];
    }
    __syncthreads();
    if (tid < 64) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) o += red[w][tid];
        out[2 + tid] = o;
    }
    if (tid == 0) { out[0] = m; out[1] = l; }
}

// ------------------------------------------------------------------------------------------------
// Logits filter + greedy sample + decoder-state update (whisper_process_logits /
// whisper_sample_token(best) / the per-token bookkeeping of whisper_full).
// Single CTA of 1024 threads, logits held in registers (n_vocab <= 52 * 1024).
// ------------------------------------------------------------------------------------------------
struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

__global__ void __launch_bounds__(1024) dec_sample_kernel(SampleArgs a) {
    DecCtl *ctl = a.ctl;
    if (ctl->done) return;
    const int tid = threadIdx.x;
    const int jrel = ctl->pos - ctl->pos0;
    if (jrel < ctl->n_prompt - 1) {   // still feeding the prompt
        if (tid == 0) { ctl->token = ctl->prompt[jrel + 1]; ctl->pos = ctl->pos + 1; }
        return;
    }
    const int nv = a.n_vocab;
    if (ctl->keep_logits) {
        const int slot = ctl->n_kept;
        if (slot < a.keep_cap) for (int i = tid; i < nv; i += 1024) a.keep[(size_t)slot * nv + i] = a.logits[i];
        __syncthreads();
        if (tid == 0) ctl->n_kept = slot + 1;
    }
    if (!ctl->sample) { if (tid == 0) ctl->done = 1; return; }

    const int n_s = ctl->n_sampled;
    const bool is_initial = n_s == 0;
    const bool last_ts = n_s > 0 && ctl->last_id >= a.beg;
    const bool penult_ts = n_s < 2 || ctl->penult_id >= a.beg;
    const int ts_floor = ctl->has_ts ? a.beg + ctl->seek_delta / 2 : a.beg;   // timestamps below are masked
    const int ts_ceil = (is_initial && a.tid0_init >= 0) ? a.beg + a.tid0_init + 1 : nv;
    constexpr int MAXJ = 52;
    float v[MAXJ];
    MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};   // text (< beg), timestamps (>= beg)
#pragma unroll
    for (int j = 0; j < MAXJ; j++) {
        const int i = tid + j * 1024;
        float x = -INFINITY;
        if (i < nv) {
            bool masked = false;
            if (is_initial && a.suppress_blank && (i == a.eot || i == a.blank)) masked = true;
            if (i == a.not_ || i == a.sot || i == a.nosp || i == a.translate || i == a.transcribe || i == a.prev) masked = true;
            if (!a.tdrz && i == a.solm) masked = true;
            if (i > a.sot && i <= a.sot + kNumLangSuppress) masked = true;
            if (last_ts) { if (penult_ts) { if (i >= a.beg) masked = true; } else { if (i < a.eot) masked = true; } }
            if (i >= ts_ceil) masked = true;
            if (i >= a.beg && i < ts_floor) masked = true;
            if (!masked) x = a.logits[i] * a.inv_temperature;
        }
        v[j] = x;
        if (i < a.beg) { if (x > mt.v) mt = MaxIdx{x, i}; } else { if (x > ms.v) ms = MaxIdx{x, i}; }
    }
    __shared__ MaxIdx s_mt[32], s_ms[32];
    __shared__ float s_a[32], s_b[32];
    __shared__ float bc[4]; __shared__ int bci[2];
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx t{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, t);
        MaxIdx u{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, u);
    }
    if (lane == 0) { s_mt[warp] = mt; s_ms[warp] = ms; }
    __syncthreads();
    if (warp == 0) {
        mt = s_mt[lane]; ms = s_ms[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MaxIdx t{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, t);
            MaxIdx u{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, u);
        }
        if (lane == 0) { bc[0] = mt.v; bc[1] = ms.v; bci[0] = mt.i; bci[1] = ms.i; }
    }
    __syncthreads();
    const float max_text = bc[0], max_ts = bc[1];
    const int idx_text = bci[0], idx_ts = bci[1];
    const float max_all = fmaxf(max_text, max_ts);
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < MAXJ; j++) {
        const int i = tid + j * 1024;
        if (v[j] > -INFINITY) { sa += expf(v[j] - max_all); if (i >= a.beg) sb += expf(v[j] - max_ts); }
    }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane == 0) { s_a[warp] = sa; s_b[warp] = sb; }
    __syncthreads();
    if (tid == 0) {
        float sum_all = 0.f, sum_ts = 0.f;
        for (int w = 0; w < 32; w++) { sum_all += s_a[w]; sum_ts += s_b[w]; }
        const float lse = logf(sum_all) + max_all;
        const float ts_lp = sum_ts > 0.f ? logf(sum_ts) + (max_ts - lse) : -INFINITY;
        const float text_lp = max_text - lse;
        TokData t;
        if (ts_lp > text_lp) { t.id = idx_ts; t.plog = max_ts - lse; }
        else if (max_text >= max_ts) { t.id = idx_text; t.plog = text_lp; }
        else { t.id = idx_ts; t.plog = max_ts - lse; }
        if (t.id == 0x7fffffff) { t.id = 0; t.plog = -INFINITY; }
        t.p = expf(t.plog);
        const float p_ts_max = max_ts > -INFINITY ? expf(max_ts - lse) : 0.f;
        const float p_ts_sum = sum_ts * p_ts_max;
        t.tid = (max_ts > -INFINITY && p_ts_max > 0.f) ? idx_ts : 0;
        t.pt = p_ts_max / (p_ts_sum + 1e-10f); t.ptsum = p_ts_sum;
        if (t.id >= a.beg) { t.tid = t.id; t.pt = t.p; }
        const int i = n_s;
        a.out[i] = t;
        // ---- whisper_full per-token decoder bookkeeping
        int has_ts = ctl->has_ts, seek_delta = ctl->seek_delta, result_len = ctl->result_len;
        int failed = 0, completed = 0;
        if (t.id > a.beg) {
            const int sd_new = 2 * (t.id - a.beg);
            if (has_ts && seek_delta > sd_new && result_len < i) failed = 1;
            else { seek_delta = sd_new; result_len = i + 1; has_ts = 1; }
        }
        if (!failed) {
            if (t.id == a.eot || (has_ts && ctl->seek + seek_delta + 100 >= ctl->seek_end)) {
                if (result_len == 0) {
                    if (ctl->seek + seek_delta + 100 >= ctl->seek_end) result_len = i + 1; else failed = 1;
                }
                if (!failed) completed = 1;
            }
        }
        if (!failed && !completed && i == ctl->n_max - 1 && (result_len == 0 || seek_delta < 100 * kChunkSec / 2)) failed = 1;
        ctl->has_ts = has_ts; ctl->seek_delta = seek_delta; ctl->result_len = result_len;
        ctl->failed = failed; ctl->completed = completed;
        ctl->penult_id = ctl->last_id; ctl->last_id = t.id; ctl->n_sampled = i + 1;
        if (failed || completed || i + 1 >= ctl->n_max) ctl->done = 1;
        else { ctl->token = t.id; ctl->pos = ctl->pos + 1; }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static int gemv_grid(int N) { int g = (N + 7) / 8; const int cap = 148 * 4; return g < cap ? g : cap; }

static void launch_gemv(GemvArgs a, cudaStream_t st, int *launches) {
    dec_gemv_kernel<<<gemv_grid(a.N), 256, (size_t)a.K * sizeof(float), st>>>(a);
    (*launches)++;
}
static void launch_attn(AttnArgs a, int n_split, int H, cudaStream_t st, int *launches) {
    dec_attn_kernel<<<dim3(n_split, H), 256, 0, st>>>(a);
    (*launches)++;
}

void decode_step_enqueue(const Model &m, const DecodeBuffers &b, cudaStream_t st, int *launches) {
    const HParams &hp = m.hp;
    const int d = hp.n_text_state, H = hp.n_text_head, T = hp.n_audio_ctx, ctx = hp.n_text_ctx;
    const float s4 = powf((float)(d / H), -0.25f);
    dec_embed_kernel<<<ceil_div(d, 256), 256, 0, st>>>(m.tok_emb, m.d_pos, b.x, d, b.ctl);
    (*launches)++;
    for (int il = 0; il < hp.n_text_layer; il++) {
        const DecLayer &L = m.dec[il];
        __half *sk = b.self_k + (size_t)il * ctx * d, *sv = b.self_v + (size_t)il * ctx * d;
        const __half *ck = b.cross_k + (size_t)il * T * d, *cv = b.cross_v + (size_t)il * T * d;
        GemvArgs g{};
        g.ctl = b.ctl; g.s4 = s4; g.ctx = ctx;
        // LN + QKV, K/V appended to the self cache
        g.W = L.qkv.w; g.bias = L.qkv.b; g.N = 3 * d; g.K = d; g.pro = PRO_LN; g.xin = b.x; g.lnw = L.attn_ln.w; g.lnb = L.attn_ln.b;
        g.epi = EPI_QKV; g.out = b.q; g.kcache = sk; g.vcache = sv;
        launch_gemv(g, st, launches);
        AttnArgs at{}; at.ctl = b.ctl; at.q = b.q; at.K = sk; at.V = sv; at.ctx = ctx; at.n_keys = -1; at.part = b.part;
        launch_attn(at, kSelfSplit, H, st, launches);
        // out-proj (+ combine, + residual)
        g = GemvArgs{}; g.ctl = b.ctl; g.W = L.o.w; g.bias = L.o.b; g.N = d; g.K = d; g.pro = PRO_ATTN; g.part = b.part; g.n_split = kSelfSplit;
        g.epi = EPI_RESID; g.out = b.x;
        launch_gemv(g, st, launches);
        // LN + cross Q
        g = GemvArgs{}; g.ctl = b.ctl; g.s4 = s4; g.W = L.cq.w; g.bias = L.cq.b; g.N = d; g.K = d; g.pro = PRO_LN; g.xin = b.x; g.lnw = L.cross_ln.w; g.lnb = L.cross_ln.b;
        g.epi = EPI_QSCALE; g.out = b.q;
        launch_gemv(g, st, launches);
        at = AttnArgs{}; at.ctl = b.ctl; at.q = b.q; at.K = ck; at.V = cv; at.ctx = T; at.n_keys = T; at.part = b.part;
        launch_attn(at, kCrossSplit, H, st, launches);
        g = GemvArgs{}; g.ctl = b.ctl; g.W = L.co.w; g.bias = L.co.b; g.N = d; g.K = d; g.pro = PRO_ATTN; g.part = b.part; g.n_split = kCrossSplit;
        g.epi = EPI_RESID; g.out = b.x;
        launch_gemv(g, st, launches);
        // MLP
        g = GemvArgs{}; g.ctl = b.ctl; g.W = L.fc1.w; g.bias = L.fc1.b; g.N = 4 * d; g.K = d; g.pro = PRO_LN; g.xin = b.x; g.lnw = L.mlp_ln.w; g.lnb = L.mlp_ln.b;
        g.epi = EPI_GELU; g.out = b.h;
        launch_gemv(g, st, launches);
        g = GemvArgs{}; g.ctl = b.ctl; g.W = L.fc2.w; g.bias = L.fc2.b; g.N = d; g.K = 4 * d; g.pro = PRO_PLAIN; g.xin = b.h;
        g.epi = EPI_RESID; g.out = b.x;
        launch_gemv(g, st, launches);
    }
    GemvArgs g{};
    g.ctl = b.ctl; g.W = m.tok_emb; g.bias = nullptr; g.N = hp.n_vocab; g.K = d; g.pro = PRO_LN; g.xin = b.x; g.lnw = m.d_ln.w; g.lnb = m.d_ln.b;
    g.epi = EPI_STORE; g.out = b.logits; g.logits_gate = 1;
    launch_gemv(g, st, launches);
    SampleArgs s{};
    const Vocab &v = m.vocab;
    s.ctl = b.ctl; s.logits = b.logits; s.n_vocab = hp.n_vocab; s.out = b.tok_out; s.keep = b.keep; s.keep_cap = b.keep_cap;
    s.eot = v.eot; s.sot = v.sot; s.translate = v.translate; s.transcribe = v.transcribe; s.solm = v.solm; s.prev = v.prev;
    s.nosp = v.nosp; s.not_ = v.not_; s.beg = v.beg; s.blank = v.blank;
    s.suppress_blank = b.suppress_blank; s.tdrz = b.tdrz; s.tid0_init = b.tid0_init; s.inv_temperature = 1.0f;
    dec_sample_kernel<<<1, 1024, 0, st>>>(s);
    (*launches)++;
}

int decode_step_num_launches(const Model &m) { return 1 + 8 * m.hp.n_text_layer + 2; }

void decoder_configure() {
    // FC2 of large-v3 stages 5120 floats (20 KB) - below the 48 KB default; nothing to opt into yet.
}

}  // namespace ss
