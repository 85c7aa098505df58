// decoder_mega.cu - the batch-1 autoregressive decoder as ONE persistent cooperative kernel
// (BASELINE.json north_star stage 3: "HBM-bandwidth-bound vectorised kernels with a persistent
// KV-cache"; SURVEY.md §8a rows a8/a9, §7.2 "batch-1 decode latency").
//
// Why one kernel: a decoder step is ~260 dependent mat-vec phases of 0.5-2 us of HBM traffic each.
// As separate launches the dependency latency of each launch (6.8 us measured in round-1 v0, 16 % of
// the HBM roofline) dominates.  Here one CTA per SM stays resident for the whole decode loop:
//   * warp 8 (producer) streams this CTA's static slice of the weights and of the cross-attention K/V
//     cache through a 9 x 20 KB shared-memory ring with cp.async.bulk + mbarriers.  Its schedule does
//     not depend on activations, so it runs ahead across phase, layer and token boundaries and keeps
//     the HBM pipe busy while the consumers wait at grid barriers.
//   * warps 0..7 (consumers) compute dot products straight out of the ring against an activation
//     vector held in registers, publish the phase's outputs (a few KB) to L2 and meet at a grid-wide
//     barrier (one monotonic counter, release/acquire).
//   * logits filter, greedy sampling and the decoder-state update (whisper_process_logits /
//     whisper_sample_token / whisper_full bookkeeping) are folded in: every CTA
//     reduces its slice of the vocabulary, all CTAs combine the per-CTA records redundantly, so the
//     next token is known everywhere without another barrier and the loop never returns to the host.
//
// Arithmetic is the oracle's (oracle/whisper_oracle.c wo_decode / process_logits): f16 weights,
// activations rounded to f16 in front of every mat-vec, f32 accumulation.
#include <cooperative_groups.h>

#include "kernels.h"

namespace ss {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kMegaThreads = kConsumerThreads + 32;
constexpr int kSlots = 9;
constexpr int kChunkBytes = 20480;
constexpr int kMaxXs = 5120;          // largest mat-vec input (4 * d, d <= 1280)
constexpr int kMaxRowsPerCta = 512;   // per phase, x KQ partials
constexpr int kMaxScores = 512;
constexpr int kMaxJ = 5;              // d / 8 / 32 uint4 chunks per lane, d <= 1280

struct __align__(16) MegaSmem {
    uint8_t ring[kSlots][kChunkBytes];
    float xs[kMaxXs];
    float acc[kMaxRowsPerCta];
    float sc[kMaxScores];
    float red[kConsumerWarps][64];
    float red1[32];
    int redi[32];
    uint64_t full[kSlots];
    uint64_t empty[kSlots];
    volatile int stop_req;      // consumers -> producer: stop issuing
    volatile int prod_done;     // producer -> consumers: `issued` is final
    volatile uint32_t issued;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float gelu16(float x) {
    const float xh = r16(x);
    return r16(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh))));
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
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ float2 h2f(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }

// ------------------------------------------------------------------------------------------------
// static work schedule: which bytes CTA `c` streams for segment kind `k` of layer `l`
// ------------------------------------------------------------------------------------------------
enum SegKind : int { SEG_QKV = 0, SEG_O, SEG_CQ, SEG_XK, SEG_XV, SEG_CO, SEG_FC1, SEG_FC2, SEG_LM, SEG_COUNT };

struct Seg {
    const uint8_t *base;   // first byte of this CTA's slice
    int rows;              // rows (or keys) in the slice
    int row0;              // first row index (global)
    int row_bytes;
    int rows_per_chunk;
    int n_chunks;
};

__device__ __forceinline__ Seg make_seg(const MegaParams &P, int kind, int layer, int cta, int ncta) {
    Seg s;
    const int d = P.d;
    const __half *w = nullptr;
    int N = 0, K = d;
    if (kind == SEG_XK || kind == SEG_XV) {
        const int unit = cta;   // (head, split) unit; CTAs beyond H * xsplit idle in this phase
        if (unit >= P.H * P.xsplit) { s.base = nullptr; s.rows = 0; s.row0 = 0; s.row_bytes = 128; s.rows_per_chunk = kChunkBytes / 128; s.n_chunks = 0; return s; }
        const int h = unit / P.xsplit, sp = unit % P.xsplit;
        const int per = (P.T + P.xsplit - 1) / P.xsplit;
        const int j0 = sp * per, j1 = min(P.T, j0 + per);
        const __half *b = (kind == SEG_XK ? P.cross_k : P.cross_v) + (size_t)layer * P.T * d + ((size_t)h * P.T + j0) * 64;
        s.base = reinterpret_cast<const uint8_t *>(b);
        s.rows = max(0, j1 - j0); s.row0 = j0; s.row_bytes = 128;
    } else {
        const MegaLayer &L = P.layer[layer];
        switch (kind) {
            case SEG_QKV: w = L.qkv_w; N = 3 * d; break;
            case SEG_O: w = L.o_w; N = d; break;
            case SEG_CQ: w = L.cq_w; N = d; break;
            case SEG_CO: w = L.co_w; N = d; break;
            case SEG_FC1: w = L.fc1_w; N = 4 * d; break;
            case SEG_FC2: w = L.fc2_w; N = d; K = 4 * d; break;
            default: w = P.tok_emb; N = P.n_vocab; break;
        }
        const int r0 = (int)((long)cta * N / ncta), r1 = (int)((long)(cta + 1) * N / ncta);
        s.base = reinterpret_cast<const uint8_t *>(w + (size_t)r0 * K);
        s.rows = r1 - r0; s.row0 = r0; s.row_bytes = K * 2;
    }
    s.rows_per_chunk = max(1, kChunkBytes / s.row_bytes);
    s.n_chunks = (s.rows + s.rows_per_chunk - 1) / s.rows_per_chunk;
    return s;
}

// ------------------------------------------------------------------------------------------------
// grid barrier: one monotonic counter (reset by the host before launch)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_sync(unsigned int *bar, unsigned int &target, int ncta) {
    consumer_sync();
    if (threadIdx.x == 0) {
        target += (unsigned int)ncta;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        __threadfence();
    }
    consumer_sync();
}

// consumer: rows of the current segment against the x vector staged in sm.xs; KQ warps share a row
// (K split in KQ slices of d) so the slice of x a lane needs lives in registers.
template <int KQ>
__device__ __forceinline__ void gemv_rows(MegaSmem &sm, uint32_t &cons, const Seg &seg, int d, int warp, int lane) {
    const int quarter = KQ == 1 ? 0 : (warp & (KQ - 1));
    const int group = warp / KQ;                 // row group of this warp
    constexpr int kGroups = kConsumerWarps / KQ;
    const int nchunk = d >> 3;                   // uint4 chunks per row slice
    float4 xa[kMaxJ], xb[kMaxJ];
    {
        const float4 *x4 = reinterpret_cast<const float4 *>(sm.xs + quarter * d);
#pragma unroll
        for (int j = 0; j < kMaxJ; j++) {
            const int c = lane + 32 * j;
            if (c < nchunk) { xa[j] = x4[2 * c]; xb[j] = x4[2 * c + 1]; }
            else { xa[j] = make_float4(0, 0, 0, 0); xb[j] = xa[j]; }
        }
    }
    for (int ch = 0; ch < seg.n_chunks; ch++) {
        const int slot = cons % kSlots;
        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const int rbase = ch * seg.rows_per_chunk;
        const int nrows = min(seg.rows_per_chunk, seg.rows - rbase);
        for (int r = 0; r < nrows; r++) {
            const int R = rbase + r;
            if (R % kGroups != group) continue;
            const uint4 *w = reinterpret_cast<const uint4 *>(sm.ring[slot] + (size_t)r * seg.row_bytes + (size_t)quarter * d * 2);
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) {
                const int c = lane + 32 * j;
                if (c < nchunk) {
                    const uint4 u = w[c];
                    float2 f;
                    f = h2f(u.x); a = fmaf(f.x, xa[j].x, a); a = fmaf(f.y, xa[j].y, a);
                    f = h2f(u.y); a = fmaf(f.x, xa[j].z, a); a = fmaf(f.y, xa[j].w, a);
                    f = h2f(u.z); a = fmaf(f.x, xb[j].x, a); a = fmaf(f.y, xb[j].y, a);
                    f = h2f(u.w); a = fmaf(f.x, xb[j].z, a); a = fmaf(f.y, xb[j].w, a);
                }
            }
            a = warp_sum(a);
            if (lane == 0) sm.acc[R * KQ + quarter] = a;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
}

// block-wide (256 consumer threads) sum; result broadcast
__device__ __forceinline__ float consumer_sum(MegaSmem &sm, float v) {
    v = warp_sum(v);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) sm.red1[threadIdx.x >> 5] = v;
    consumer_sync();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) t += sm.red1[i];
    return t;
}
__device__ __forceinline__ float consumer_max(MegaSmem &sm, float v) {
    v = warp_max(v);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) sm.red1[threadIdx.x >> 5] = v;
    consumer_sync();
    float t = sm.red1[0];
#pragma unroll
    for (int i = 1; i < kConsumerWarps; i++) t = fmaxf(t, sm.red1[i]);
    return t;
}

// LayerNorm of the K-vector already in sm.xs (ggml_norm + affine, eps 1e-5), rounded to f16
__device__ __forceinline__ void ln_inplace(MegaSmem &sm, int K, const float *__restrict__ w, const float *__restrict__ b) {
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int i = tid; i < K; i += kConsumerThreads) s += sm.xs[i];
    const float mean = consumer_sum(sm, s) / K;
    float s2 = 0.f;
    for (int i = tid; i < K; i += kConsumerThreads) { const float v = sm.xs[i] - mean; sm.xs[i] = v; s2 += v * v; }
    const float var = consumer_sum(sm, s2) / K;
    const float scale = rsqrtf(var + 1e-5f);
    for (int i = tid; i < K; i += kConsumerThreads) sm.xs[i] = r16(sm.xs[i] * scale * __ldg(w + i) + __ldg(b + i));
    consumer_sync();
}

// combine split-softmax partials [H][ns][66] into the attention vector (f16-rounded) in sm.xs
__device__ __forceinline__ void combine_partials(MegaSmem &sm, const float *part, int d, int ns) {
    for (int n = threadIdx.x; n < d; n += kConsumerThreads) {
        const int h = n >> 6, c = n & 63;
        const float *p = part + (size_t)h * ns * 66;
        float M = -INFINITY;
        for (int s = 0; s < ns; s++) M = fmaxf(M, __ldcg(p + s * 66));
        float L = 0.f, o = 0.f;
        for (int s = 0; s < ns; s++) {
            const float ms = __ldcg(p + s * 66);
            if (ms == -INFINITY) continue;
            const float e = __expf(ms - M);
            L += __ldcg(p + s * 66 + 1) * e; o += __ldcg(p + s * 66 + 2 + c) * e;
        }
        sm.xs[n] = r16(o / L);
    }
    consumer_sync();
}

struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

struct DecState {   // replicated in every CTA's registers (thread-uniform)
    int pos, token, n_sampled, has_ts, seek_delta, result_len, last_id, penult_id, n_kept;
};

__device__ __forceinline__ bool token_masked(const MegaParams &P, const DecState &st, int i) {
    const bool is_initial = st.n_sampled == 0;
    if (is_initial && P.suppress_blank && (i == P.eot || i == P.blank)) return true;
    if (i == P.not_ || i == P.sot || i == P.nosp || i == P.translate || i == P.transcribe || i == P.prev) return true;
    if (!P.tdrz && i == P.solm) return true;
    if (i > P.sot && i <= P.sot + kNumLangSuppress) return true;
    const bool last_ts = st.n_sampled > 0 && st.last_id >= P.beg;
    const bool penult_ts = st.n_sampled < 2 || st.penult_id >= P.beg;
    if (last_ts) { if (penult_ts) { if (i >= P.beg) return true; } else { if (i < P.eot) return true; } }
    if (is_initial && P.tid0_init >= 0 && i >= P.beg + P.tid0_init + 1) return true;
    if (st.has_ts && i >= P.beg && i < P.beg + st.seek_delta / 2) return true;
    return false;
}

}  // namespace

// ================================================================================================
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaParams *__restrict__ Pp, int max_steps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    MegaSmem &sm = *reinterpret_cast<MegaSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const MegaParams &P = *Pp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int d = P.d, H = P.H, L = P.L;

    DecCtl *ctl = P.ctl;
    // snapshot of the control block (identical in every CTA)
    const int pos_start = ctl->pos, pos0 = ctl->pos0, n_prompt = ctl->n_prompt, do_sample = ctl->sample;
    const int seek = ctl->seek, seek_end = ctl->seek_end, n_max = ctl->n_max, keep_logits = ctl->keep_logits, all_logits = ctl->all_logits;
    const int already_done = ctl->done;
    const int steps_left = max_steps;

    if (tid == 0) {
        for (int s = 0; s < kSlots; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps); }
        sm.stop_req = 0; sm.prod_done = 0; sm.issued = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (already_done || steps_left <= 0) return;

    if (warp == kConsumerWarps) {
        // ======================= producer =======================
        if (lane == 0) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            uint32_t issued = 0;
            bool stopped = false;
            for (int t = 0; t < steps_left && !stopped; t++) {
                const int jrel = pos_start - pos0 + t;
                const bool need_logits = all_logits || jrel >= n_prompt - 1;
                const int nseg = L * (SEG_LM) + (need_logits ? 1 : 0);
                for (int si = 0; si < nseg && !stopped; si++) {
                    const int layer = si / SEG_LM, kind = si < L * SEG_LM ? si % SEG_LM : SEG_LM;
                    const Seg seg = make_seg(P, kind, layer < L ? layer : 0, cta, ncta);
                    for (int ch = 0; ch < seg.n_chunks; ch++) {
                        const int slot = issued % kSlots;
                        const uint32_t par = ((issued / kSlots) & 1) ^ 1;
                        while (!mbar_try_wait(&sm.empty[slot], par)) { if (sm.stop_req) { stopped = true; break; } }
                        if (stopped || sm.stop_req) { stopped = true; break; }
                        const int rbase = ch * seg.rows_per_chunk;
                        const uint32_t bytes = (uint32_t)min(seg.rows_per_chunk, seg.rows - rbase) * seg.row_bytes;
                        mbar_expect_tx(&sm.full[slot], bytes);
                        bulk_g2s(sm.ring[slot], seg.base + (size_t)rbase * seg.row_bytes, bytes, &sm.full[slot], policy);
                        issued++;
                    }
                }
            }
            // hand the issue count to the consumers so that they can drain in-flight copies before exit
            sm.issued = issued;
            __threadfence_block();
            sm.prod_done = 1;
        }
        return;
    }

    // ======================= consumers =======================
    uint32_t cons = 0;
    unsigned int bar_target = 0;
    DecState st;
    st.pos = pos_start; st.token = ctl->token; st.n_sampled = ctl->n_sampled; st.has_ts = ctl->has_ts; st.seek_delta = ctl->seek_delta;
    st.result_len = ctl->result_len; st.last_id = ctl->last_id; st.penult_id = ctl->penult_id; st.n_kept = ctl->n_kept;
    int failed = 0, completed = 0, done = 0;

    for (int t = 0; t < steps_left && !done; t++) {
        const int jrel = st.pos - pos0;
        const bool need_logits = all_logits || jrel >= n_prompt - 1;
        const int n_keys = st.pos + 1;

        for (int il = 0; il < L; il++) {
            const MegaLayer &ly = P.layer[il];
            // ---------------- P0: LN1 + QKV, append K/V ----------------
            {
                if (il == 0) {
                    const __half *e = P.tok_emb + (size_t)st.token * d;
                    const float *pe = P.d_pos + (size_t)st.pos * d;
                    for (int i = tid; i < d; i += kConsumerThreads) {
                        const float v = __half2float(__ldg(e + i)) + __ldg(pe + i);
                        sm.xs[i] = v;
                        if (cta == 0) P.x[i] = v;
                    }
                } else {
                    for (int i = tid; i < d; i += kConsumerThreads) sm.xs[i] = __ldcg(P.x + i);
                }
                ln_inplace(sm, d, ly.ln1_w, ly.ln1_b);
                const Seg seg = make_seg(P, SEG_QKV, il, cta, ncta);
                gemv_rows<1>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                __half *sk = P.self_k + (size_t)il * P.ctx * d, *sv = P.self_v + (size_t)il * P.ctx * d;
                for (int R = tid; R < seg.rows; R += kConsumerThreads) {
                    const int row = seg.row0 + R;
                    const float v = sm.acc[R] + __ldg(ly.qkv_b + row);
                    if (row < d) P.q[row] = r16(v * P.s4);
                    else if (row < 2 * d) { const int n = row - d; sk[((size_t)(n >> 6) * P.ctx + st.pos) * 64 + (n & 63)] = __float2half_rn(v * P.s4); }
                    else { const int n = row - 2 * d; sv[((size_t)(n >> 6) * P.ctx + st.pos) * 64 + (n & 63)] = __float2half_rn(v); }
                }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P1: self-attention partials over the KV cache (direct global reads) ----------------
            {
                const int ns = P.ssplit;
                if (cta < H * ns) {
                    const int h = cta / ns, sp = cta % ns;
                    const int per = (n_keys + ns - 1) / ns;
                    const int j0 = sp * per, j1 = min(n_keys, j0 + per), n = max(0, j1 - j0);
                    float *out = P.part + ((size_t)h * ns + sp) * 66;
                    if (n == 0) {
                        if (tid == 0) { out[0] = -INFINITY; out[1] = 0.f; }
                        if (tid < 64) out[2 + tid] = 0.f;
                    } else {
                        const int sub = lane >> 3, l8 = lane & 7;
                        const float4 qa = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8));
                        const float4 qb = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8 + 4));
                        const __half *Kh = P.self_k + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64;
                        const __half *Vh = P.self_v + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64;
                        float lmax = -INFINITY;
                        for (int jb = 0; jb < n; jb += 32) {
                            const int j = jb + warp * 4 + sub;
                            uint4 kv = make_uint4(0, 0, 0, 0);
                            if (j < n) kv = __ldcg(reinterpret_cast<const uint4 *>(Kh + (size_t)(j0 + j) * 64 + l8 * 8));
                            float2 f; float dsum = 0.f;
                            f = h2f(kv.x); dsum = fmaf(f.x, qa.x, dsum); dsum = fmaf(f.y, qa.y, dsum);
                            f = h2f(kv.y); dsum = fmaf(f.x, qa.z, dsum); dsum = fmaf(f.y, qa.w, dsum);
                            f = h2f(kv.z); dsum = fmaf(f.x, qb.x, dsum); dsum = fmaf(f.y, qb.y, dsum);
                            f = h2f(kv.w); dsum = fmaf(f.x, qb.z, dsum); dsum = fmaf(f.y, qb.w, dsum);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
                            if (j < n) { if (l8 == 0) sm.sc[j] = dsum; lmax = fmaxf(lmax, dsum); }
                        }
                        const float m = consumer_max(sm, lmax);
                        float lsum = 0.f;
                        for (int j = tid; j < n; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; lsum += e; }
                        const float l = consumer_sum(sm, lsum);
                        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                        for (int jb = 0; jb < n; jb += 32) {
                            const int j = jb + warp * 4 + sub;
                            if (j < n) {
                                const uint4 vv = __ldcg(reinterpret_cast<const uint4 *>(Vh + (size_t)(j0 + j) * 64 + l8 * 8));
                                const float p = sm.sc[j];
                                float2 f;
                                f = h2f(vv.x); acc[0] = fmaf(p, f.x, acc[0]); acc[1] = fmaf(p, f.y, acc[1]);
                                f = h2f(vv.y); acc[2] = fmaf(p, f.x, acc[2]); acc[3] = fmaf(p, f.y, acc[3]);
                                f = h2f(vv.z); acc[4] = fmaf(p, f.x, acc[4]); acc[5] = fmaf(p, f.y, acc[5]);
                                f = h2f(vv.w); acc[6] = fmaf(p, f.x, acc[6]); acc[7] = fmaf(p, f.y, acc[7]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; i++) { acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8); acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16); }
                        if (sub == 0) {
#pragma unroll
                            for (int i = 0; i < 8; i++) sm.red[warp][l8 * 8 + i] = acc[i];
                        }
                        consumer_sync();
                        if (tid < 64) { float o = 0.f; for (int w = 0; w < kConsumerWarps; w++) o += sm.red[w][tid]; out[2 + tid] = o; }
                        if (tid == 0) { out[0] = m; out[1] = l; }
                    }
                }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P2: combine + out-proj + residual ----------------
            {
                combine_partials(sm, P.part, d, P.ssplit);
                const Seg seg = make_seg(P, SEG_O, il, cta, ncta);
                gemv_rows<1>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                for (int R = tid; R < seg.rows; R += kConsumerThreads) { const int row = seg.row0 + R; P.x[row] = __ldcg(P.x + row) + sm.acc[R] + __ldg(ly.o_b + row); }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P3: LN2 + cross query ----------------
            {
                for (int i = tid; i < d; i += kConsumerThreads) sm.xs[i] = __ldcg(P.x + i);
                ln_inplace(sm, d, ly.ln2_w, ly.ln2_b);
                const Seg seg = make_seg(P, SEG_CQ, il, cta, ncta);
                gemv_rows<1>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                for (int R = tid; R < seg.rows; R += kConsumerThreads) { const int row = seg.row0 + R; P.q[row] = r16((sm.acc[R] + __ldg(ly.cq_b + row)) * P.s4); }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P4: cross-attention partials, K then V streamed through the ring ----------------
            {
                const Seg sk = make_seg(P, SEG_XK, il, cta, ncta), sv = make_seg(P, SEG_XV, il, cta, ncta);
                if (sk.n_chunks > 0) {
                    const int ns = P.xsplit, h = cta / ns, sp = cta % ns;
                    const int sub = lane >> 3, l8 = lane & 7, n = sk.rows;
                    const float4 qa = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8));
                    const float4 qb = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8 + 4));
                    float lmax = -INFINITY;
                    for (int ch = 0; ch < sk.n_chunks; ch++) {
                        const int slot = cons % kSlots;
                        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
                        const int kbase = ch * sk.rows_per_chunk, nk = min(sk.rows_per_chunk, n - kbase);
                        for (int jb = 0; jb < nk; jb += 32) {
                            const int j = jb + warp * 4 + sub;
                            uint4 kv = make_uint4(0, 0, 0, 0);
                            if (j < nk) kv = *reinterpret_cast<const uint4 *>(sm.ring[slot] + (size_t)j * 128 + l8 * 16);
                            float2 f; float dsum = 0.f;
                            f = h2f(kv.x); dsum = fmaf(f.x, qa.x, dsum); dsum = fmaf(f.y, qa.y, dsum);
                            f = h2f(kv.y); dsum = fmaf(f.x, qa.z, dsum); dsum = fmaf(f.y, qa.w, dsum);
                            f = h2f(kv.z); dsum = fmaf(f.x, qb.x, dsum); dsum = fmaf(f.y, qb.y, dsum);
                            f = h2f(kv.w); dsum = fmaf(f.x, qb.z, dsum); dsum = fmaf(f.y, qb.w, dsum);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                            dsum += __shfl_xor_sync(0xffffffffu, dsum, 4);
                            if (j < nk) { if (l8 == 0) sm.sc[kbase + j] = dsum; lmax = fmaxf(lmax, dsum); }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.empty[slot]);
                        cons++;
                    }
                    const float m = consumer_max(sm, lmax);
                    float lsum = 0.f;
                    for (int j = tid; j < n; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; lsum += e; }
                    const float l = consumer_sum(sm, lsum);
                    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                    for (int ch = 0; ch < sv.n_chunks; ch++) {
                        const int slot = cons % kSlots;
                        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
                        const int kbase = ch * sv.rows_per_chunk, nk = min(sv.rows_per_chunk, n - kbase);
                        for (int jb = 0; jb < nk; jb += 32) {
                            const int j = jb + warp * 4 + sub;
                            if (j < nk) {
                                const uint4 vv = *reinterpret_cast<const uint4 *>(sm.ring[slot] + (size_t)j * 128 + l8 * 16);
                                const float p = sm.sc[kbase + j];
                                float2 f;
                                f = h2f(vv.x); acc[0] = fmaf(p, f.x, acc[0]); acc[1] = fmaf(p, f.y, acc[1]);
                                f = h2f(vv.y); acc[2] = fmaf(p, f.x, acc[2]); acc[3] = fmaf(p, f.y, acc[3]);
                                f = h2f(vv.z); acc[4] = fmaf(p, f.x, acc[4]); acc[5] = fmaf(p, f.y, acc[5]);
                                f = h2f(vv.w); acc[6] = fmaf(p, f.x, acc[6]); acc[7] = fmaf(p, f.y, acc[7]);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sm.empty[slot]);
                        cons++;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) { acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8); acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16); }
                    if (sub == 0) {
#pragma unroll
                        for (int i = 0; i < 8; i++) sm.red[warp][l8 * 8 + i] = acc[i];
                    }
                    consumer_sync();
                    float *out = P.part + ((size_t)h * ns + sp) * 66;
                    if (tid < 64) { float o = 0.f; for (int w = 0; w < kConsumerWarps; w++) o += sm.red[w][tid]; out[2 + tid] = o; }
                    if (tid == 0) { out[0] = m; out[1] = l; }
                }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P5: combine + cross out-proj + residual ----------------
            {
                combine_partials(sm, P.part, d, P.xsplit);
                const Seg seg = make_seg(P, SEG_CO, il, cta, ncta);
                gemv_rows<1>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                for (int R = tid; R < seg.rows; R += kConsumerThreads) { const int row = seg.row0 + R; P.x[row] = __ldcg(P.x + row) + sm.acc[R] + __ldg(ly.co_b + row); }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P6: LN3 + FC1 + GELU ----------------
            {
                for (int i = tid; i < d; i += kConsumerThreads) sm.xs[i] = __ldcg(P.x + i);
                ln_inplace(sm, d, ly.ln3_w, ly.ln3_b);
                const Seg seg = make_seg(P, SEG_FC1, il, cta, ncta);
                gemv_rows<1>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                for (int R = tid; R < seg.rows; R += kConsumerThreads) { const int row = seg.row0 + R; P.h[row] = gelu16(sm.acc[R] + __ldg(ly.fc1_b + row)); }
                grid_sync(P.bar, bar_target, ncta);
            }
            // ---------------- P7: FC2 + residual (K = 4d, four warps per row) ----------------
            {
                for (int i = tid; i < 4 * d; i += kConsumerThreads) sm.xs[i] = __ldcg(P.h + i);
                consumer_sync();
                const Seg seg = make_seg(P, SEG_FC2, il, cta, ncta);
                gemv_rows<4>(sm, cons, seg, d, warp, lane);
                consumer_sync();
                for (int R = tid; R < seg.rows; R += kConsumerThreads) {
                    const int row = seg.row0 + R;
                    P.x[row] = __ldcg(P.x + row) + (sm.acc[4 * R] + sm.acc[4 * R + 1] + sm.acc[4 * R + 2] + sm.acc[4 * R + 3]) + __ldg(ly.fc2_b + row);
                }
                grid_sync(P.bar, bar_target, ncta);
            }
        }

        // ---------------- final LN + LM head + per-CTA softmax statistics ----------------
        if (need_logits) {
            for (int i = tid; i < d; i += kConsumerThreads) sm.xs[i] = __ldcg(P.x + i);
            ln_inplace(sm, d, P.lnf_w, P.lnf_b);
            const Seg seg = make_seg(P, SEG_LM, 0, cta, ncta);
            gemv_rows<1>(sm, cons, seg, d, warp, lane);
            consumer_sync();
            const bool keep = keep_logits && st.n_kept < P.keep_cap;
            MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
            for (int R = tid; R < seg.rows; R += kConsumerThreads) {
                const int i = seg.row0 + R;
                const float raw = sm.acc[R];
                P.logits[i] = raw;
                if (keep) P.keep[(size_t)st.n_kept * P.n_vocab + i] = raw;
                float x = -INFINITY;
                if (do_sample && jrel >= n_prompt - 1 && !token_masked(P, st, i)) x = raw;
                sm.acc[R] = x;
                if (i < P.beg) { if (x > mt.v) mt = MaxIdx{x, i}; } else { if (x > ms.v) ms = MaxIdx{x, i}; }
            }
            if (do_sample && jrel >= n_prompt - 1) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
                    MaxIdx b{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, b);
                }
                consumer_sync();
                if (lane == 0) { sm.red[0][warp] = mt.v; sm.redi[warp] = mt.i; sm.red[1][warp] = ms.v; sm.redi[8 + warp] = ms.i; }
                consumer_sync();
                mt = MaxIdx{sm.red[0][0], sm.redi[0]}; ms = MaxIdx{sm.red[1][0], sm.redi[8]};
                for (int w = 1; w < kConsumerWarps; w++) { mt = better(mt, MaxIdx{sm.red[0][w], sm.redi[w]}); ms = better(ms, MaxIdx{sm.red[1][w], sm.redi[8 + w]}); }
                const float m_all = fmaxf(mt.v, ms.v);
                float sa = 0.f, sb = 0.f;
                for (int R = tid; R < seg.rows; R += kConsumerThreads) {
                    const float x = sm.acc[R];
                    if (x > -INFINITY) { sa += expf(x - m_all); if (seg.row0 + R >= P.beg) sb += expf(x - ms.v); }
                }
                sa = consumer_sum(sm, sa);
                sb = consumer_sum(sm, sb);
                if (tid == 0) {
                    float *rec = P.stats + (size_t)cta * 8;
                    rec[0] = mt.v; rec[1] = __int_as_float(mt.i); rec[2] = ms.v; rec[3] = __int_as_float(ms.i); rec[4] = sa; rec[5] = sb;
                }
            }
            grid_sync(P.bar, bar_target, ncta);
        }
        if (need_logits && keep_logits && st.n_kept < P.keep_cap) st.n_kept++;
        if (jrel < n_prompt - 1) { st.token = ctl->prompt[jrel + 1]; st.pos++; continue; }   // prompt token: feed the next one
        if (!do_sample) { done = 1; break; }
        // ---------------- combine the per-CTA records (every CTA, identically) + sample + bookkeeping ----------------
        {
            if (warp == 0) {
                MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
                for (int c = lane; c < ncta; c += 32) {
                    const float *rec = P.stats + (size_t)c * 8;
                    mt = better(mt, MaxIdx{__ldcg(rec + 0), __float_as_int(__ldcg(rec + 1))});
                    ms = better(ms, MaxIdx{__ldcg(rec + 2), __float_as_int(__ldcg(rec + 3))});
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
                    MaxIdx b{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, b);
                }
                const float m_all = fmaxf(mt.v, ms.v);
                float sa = 0.f, sb = 0.f;
                for (int c = lane; c < ncta; c += 32) {
                    const float *rec = P.stats + (size_t)c * 8;
                    const float cm = fmaxf(__ldcg(rec + 0), __ldcg(rec + 2)), cs = __ldcg(rec + 2);
                    if (cm > -INFINITY) sa += __ldcg(rec + 4) * expf(cm - m_all);
                    if (cs > -INFINITY) sb += __ldcg(rec + 5) * expf(cs - ms.v);
                }
                sa = warp_sum(sa); sb = warp_sum(sb);
                if (lane == 0) {
                    const float max_text = mt.v, max_ts = ms.v;
                    const float lse = logf(sa) + m_all;
                    const float ts_lp = sb > 0.f ? logf(sb) + (max_ts - lse) : -INFINITY;
                    const float text_lp = max_text - lse;
                    TokData tk;
                    if (ts_lp > text_lp) { tk.id = ms.i; tk.plog = max_ts - lse; }
                    else if (max_text >= max_ts) { tk.id = mt.i; tk.plog = text_lp; }
                    else { tk.id = ms.i; tk.plog = max_ts - lse; }
                    if (tk.id == 0x7fffffff) { tk.id = 0; tk.plog = -INFINITY; }
                    tk.p = expf(tk.plog);
                    const float p_ts_max = max_ts > -INFINITY ? expf(max_ts - lse) : 0.f;
                    const float p_ts_sum = sb * p_ts_max;
                    tk.tid = (max_ts > -INFINITY && p_ts_max > 0.f) ? ms.i : 0;
                    tk.pt = p_ts_max / (p_ts_sum + 1e-10f); tk.ptsum = p_ts_sum;
                    if (tk.id >= P.beg) { tk.tid = tk.id; tk.pt = tk.p; }
                    const int i = st.n_sampled;
                    if (cta == 0) P.tok_out[i] = tk;
                    int f = 0, cpl = 0;
                    int has_ts = st.has_ts, seek_delta = st.seek_delta, result_len = st.result_len;
                    if (tk.id > P.beg) {
                        const int sd_new = 2 * (tk.id - P.beg);
                        if (has_ts && seek_delta > sd_new && result_len < i) f = 1;
                        else { seek_delta = sd_new; result_len = i + 1; has_ts = 1; }
                    }
                    if (!f) {
                        if (tk.id == P.eot || (has_ts && seek + seek_delta + 100 >= seek_end)) {
                            if (result_len == 0) { if (seek + seek_delta + 100 >= seek_end) result_len = i + 1; else f = 1; }
                            if (!f) cpl = 1;
                        }
                    }
                    if (!f && !cpl && i == n_max - 1 && (result_len == 0 || seek_delta < 100 * kChunkSec / 2)) f = 1;
                    sm.redi[16] = tk.id; sm.redi[17] = has_ts; sm.redi[18] = seek_delta; sm.redi[19] = result_len; sm.redi[20] = f; sm.redi[21] = cpl;
                }
            }
            consumer_sync();
            const int id = sm.redi[16];
            st.has_ts = sm.redi[17]; st.seek_delta = sm.redi[18]; st.result_len = sm.redi[19]; failed = sm.redi[20]; completed = sm.redi[21];
            st.penult_id = st.last_id; st.last_id = id; st.n_sampled++;
            if (failed || completed || st.n_sampled >= n_max) done = 1;
            else { st.token = id; st.pos++; }
            consumer_sync();
        }
    }

    // ---------------- shutdown: stop the producer, drain copies still in flight, publish the state ----------------
    if (tid == 0) {
        sm.stop_req = 1;
        while (!sm.prod_done) {}
        __threadfence_block();
        const uint32_t issued = sm.issued;
        for (uint32_t c = cons; c < issued; c++) mbar_wait(&sm.full[c % kSlots], (c / kSlots) & 1);
        if (cta == 0) {
            ctl->pos = st.pos; ctl->token = st.token; ctl->n_sampled = st.n_sampled; ctl->has_ts = st.has_ts; ctl->seek_delta = st.seek_delta;
            ctl->result_len = st.result_len; ctl->last_id = st.last_id; ctl->penult_id = st.penult_id; ctl->n_kept = st.n_kept;
            ctl->failed = failed; ctl->completed = completed; ctl->done = done;
        }
    }
}

// ------------------------------------------------------------------------------------------------
size_t decode_mega_smem_bytes() { return sizeof(MegaSmem) + 128; }

void decode_mega_configure() {
    CUDA_CHECK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decode_mega_smem_bytes()));
}

int decode_mega_grid(int device) {
    int sms = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_mega_kernel, kMegaThreads, decode_mega_smem_bytes()));
    if (per_sm < 1) SS_THROW(-4, "decode_mega_kernel does not fit on an SM");
    return sms;
}

void decode_mega_launch(const MegaParams *d_params, unsigned int *d_bar, int max_steps, int grid, cudaStream_t st) {
    CUDA_CHECK(cudaMemsetAsync(d_bar, 0, sizeof(unsigned int), st));
    void *args[] = {(void *)&d_params, (void *)&max_steps};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)decode_mega_kernel, dim3(grid), dim3(kMegaThreads), args, decode_mega_smem_bytes(), st));
}

}  // namespace ss
