// decoder_mega.cu - the batch-1 autoregressive decoder as ONE persistent cooperative kernel
// (BASELINE.json north_star stage 3: "HBM-bandwidth-bound vectorised kernels with a persistent
// KV-cache"; SURVEY.md §8a rows a8/a9, §7.2 "batch-1 decode latency").
//
// Why one kernel: a decoder step is ~260 dependent mat-vec phases of 0.5-2 us of HBM traffic each.
// As separate launches the dependency latency of each launch (6.8 us measured in round-1 v0, 16 % of
// the HBM roofline) dominates.  Here one CTA per SM stays resident for the whole decode loop:
//   * warp 8 (producer) streams this CTA's static slice of the weights and of the cross-attention K/V
//     cache through an 8 x 20 KB shared-memory ring with cp.async.bulk + mbarriers.  Its schedule does
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
constexpr int kSlots = 8;
constexpr int kChunkBytes = 20480;
constexpr int kMaxXs = 5120;          // largest mat-vec input (4 * d, d <= 1280)
constexpr int kMaxRowsPerCta = 512;   // per phase, x KQ partials
constexpr int kMaxScores = 512;
constexpr int kMaxJ = 5;              // d / 8 / 32 uint4 chunks per lane, d <= 1280

struct DecState {   // replicated per CTA (thread-uniform, lives in shared memory)
    int pos, token, n_sampled, has_ts, seek_delta, result_len, last_id, penult_id, n_kept, failed, completed, done;
};

struct __align__(128) MegaSmem {
    uint8_t ring[kSlots][kChunkBytes];
    float xs[kMaxXs];
    float lnw[1280], lnb[1280];     // LayerNorm affine of the coming phase (cp.async prefetched across the barrier)
    float bias[kMaxRowsPerCta];     // bias slice of this CTA's rows for the coming phase
    MegaParams P;                   // descriptor copy: no pointer chasing through L2 on the critical path
    float acc[kMaxRowsPerCta];
    float sc[kMaxScores];
    float red[kConsumerWarps][64];
    float red1[32];
    int redi[32];
    uint64_t full[kSlots];
    uint64_t empty[kSlots];
    DecState st;
    unsigned int bar_target, xuse;
    volatile int stop_req;      // consumers -> producer: stop issuing
    volatile int prod_done;     // producer -> consumers: `issued` is final
    volatile uint32_t issued;
    long long prof[8];          // thread-0 cycle counters: arrive, spin, gemv, total, barriers
};

// The kernel's code is deliberately small: every phase of every layer runs through the SAME few
// non-inlined routines.  (A fully inlined first version was 183 KB of SASS and ran at instruction-fetch
// speed: each phase's private copy of the mat-vec loop missed the instruction caches every time.)
extern __shared__ __align__(128) uint8_t mega_smem_raw[];
#define SM (*reinterpret_cast<MegaSmem *>(mega_smem_raw))

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
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ float2 h2f(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }
__device__ __forceinline__ float dot8(const uint4 &w, const float4 &a, const float4 &b, float acc) {
    float2 f;
    f = h2f(w.x); acc = fmaf(f.x, a.x, acc); acc = fmaf(f.y, a.y, acc);
    f = h2f(w.y); acc = fmaf(f.x, a.z, acc); acc = fmaf(f.y, a.w, acc);
    f = h2f(w.z); acc = fmaf(f.x, b.x, acc); acc = fmaf(f.y, b.y, acc);
    f = h2f(w.w); acc = fmaf(f.x, b.z, acc); acc = fmaf(f.y, b.w, acc);
    return acc;
}
__device__ __forceinline__ void axpy8(const uint4 &v, float p, float (&acc)[8]) {
    float2 f;
    f = h2f(v.x); acc[0] = fmaf(p, f.x, acc[0]); acc[1] = fmaf(p, f.y, acc[1]);
    f = h2f(v.y); acc[2] = fmaf(p, f.x, acc[2]); acc[3] = fmaf(p, f.y, acc[3]);
    f = h2f(v.z); acc[4] = fmaf(p, f.x, acc[4]); acc[5] = fmaf(p, f.y, acc[5]);
    f = h2f(v.w); acc[6] = fmaf(p, f.x, acc[6]); acc[7] = fmaf(p, f.y, acc[7]);
}

// ------------------------------------------------------------------------------------------------
// static work schedule: which bytes CTA `c` streams for segment kind `k` of layer `l`
// ------------------------------------------------------------------------------------------------
enum SegKind : int { SEG_QKV = 0, SEG_O, SEG_CQ, SEG_XK, SEG_XV, SEG_CO, SEG_FC1, SEG_FC2, SEG_LM, SEG_COUNT };
enum Phase : int { PH_QKV = 0, PH_SELF, PH_O, PH_CQ, PH_CROSS, PH_CO, PH_FC1, PH_FC2, PH_LM, PH_NONE };

struct Seg {
    const uint8_t *base;   // first byte of this CTA's slice
    int rows;              // rows (or keys) in the slice
    int row0;              // first row index (global)
    int row_bytes;
    int rows_per_chunk;
    int n_chunks;
};

__device__ __noinline__ Seg make_seg(int kind, int layer) {
    const MegaParams &P = SM.P;
    const int cta = blockIdx.x, ncta = gridDim.x;
    Seg s;
    const int d = P.d;
    const __half *w = nullptr;
    int N = 0, K = d;
    if (kind == SEG_XK || kind == SEG_XV) {
        if (cta >= P.H * P.xsplit) { s.base = nullptr; s.rows = 0; s.row0 = 0; s.row_bytes = 128; s.rows_per_chunk = kChunkBytes / 128; s.n_chunks = 0; return s; }
        const int h = cta / P.xsplit, sp = cta % P.xsplit;
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

// static (weight-side) small operands of phase `ph`: issued before a grid barrier, landed after it
__device__ __noinline__ void prefetch_static(int ph, int il) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int d = P.d, tid = threadIdx.x, cta = blockIdx.x, ncta = gridDim.x;
    const float *lw = nullptr, *lb = nullptr, *bias = nullptr;
    int N = 0;
    if (ph < PH_LM) {
        const MegaLayer &ly = P.layer[il];
        switch (ph) {
            case PH_QKV: lw = ly.ln1_w; lb = ly.ln1_b; bias = ly.qkv_b; N = 3 * d; break;
            case PH_O: bias = ly.o_b; N = d; break;
            case PH_CQ: lw = ly.ln2_w; lb = ly.ln2_b; bias = ly.cq_b; N = d; break;
            case PH_CO: bias = ly.co_b; N = d; break;
            case PH_FC1: lw = ly.ln3_w; lb = ly.ln3_b; bias = ly.fc1_b; N = 4 * d; break;
            case PH_FC2: bias = ly.fc2_b; N = d; break;
            default: break;
        }
    } else if (ph == PH_LM) { lw = P.lnf_w; lb = P.lnf_b; }
    if (lw) for (int i = tid; i < (d >> 2); i += kConsumerThreads) { cp_async16(&sm.lnw[4 * i], lw + 4 * i); cp_async16(&sm.lnb[4 * i], lb + 4 * i); }
    if (bias) {
        const int r0 = (int)((long)cta * N / ncta), r1 = (int)((long)(cta + 1) * N / ncta);
        for (int R = tid; R < r1 - r0; R += kConsumerThreads) cp_async4(&sm.bias[R], bias + r0 + R);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// grid barrier: one monotonic counter (reset by the host before launch).  Arrive = release-atomic after
// the CTA barrier (cumulative over the CTA's stores, which are therefore in L2 before the count moves);
// wait = relaxed gpu-scope poll.  Everything another CTA produced is read with ld.global.cg (L2, the
// point of coherence) after the poll exits, so no acquire fence / L1 invalidation (CCTL.IVALL) is
// needed.  The next phase's static operands are prefetched with cp.async while thread 0 polls.
__device__ __noinline__ void grid_sync(int next_ph, int next_il) {
    MegaSmem &sm = SM;
    consumer_sync();
    if (next_ph != PH_NONE) prefetch_static(next_ph, next_il);
    if (threadIdx.x == 0) {
        unsigned int *bar = sm.P.bar;
        const unsigned int target = sm.bar_target + gridDim.x;
        sm.bar_target = target;
        const long long c0 = clock64();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        const long long c1 = clock64();
        unsigned int v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while (v < target);
        sm.prof[0] += c1 - c0; sm.prof[1] += clock64() - c1; sm.prof[4] += 1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    consumer_sync();
}

// consumer: rows of the current segment against the x vector staged in sm.xs; kq (1 or 4) warps share a
// row (K split in kq slices of d) so the slice of x a lane needs lives in registers.
__device__ __noinline__ uint32_t gemv_rows(uint32_t cons, int rows, int row_bytes, int rows_per_chunk, int n_chunks, int d, int kq) {
    MegaSmem &sm = SM;
    const long long tg0 = clock64();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & (kq - 1);
    const int group = kq == 1 ? warp : (warp >> 2);
    const int gmask = kConsumerWarps / kq - 1;   // groups are a power of two
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
    for (int ch = 0; ch < n_chunks; ch++) {
        const int slot = cons % kSlots;
        const long long tw0 = clock64();
        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        if (threadIdx.x == 0) sm.prof[5] += clock64() - tw0;
        const int rbase = ch * rows_per_chunk;
        const int nrows = min(rows_per_chunk, rows - rbase);
        for (int r = 0; r < nrows; r++) {
            const int R = rbase + r;
            if ((R & gmask) != group) continue;
            const long long tr0 = clock64();
            const uint4 *w = reinterpret_cast<const uint4 *>(sm.ring[slot] + (size_t)r * row_bytes + (size_t)quarter * d * 2);
            uint4 u[kMaxJ];
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) { const int c = lane + 32 * j; u[j] = c < nchunk ? w[c] : make_uint4(0, 0, 0, 0); }
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) { if (j & 1) a1 = dot8(u[j], xa[j], xb[j], a1); else a0 = dot8(u[j], xa[j], xb[j], a0); }
            const float a = warp_sum(a0 + a1);
            if (lane == 0) sm.acc[R * kq + quarter] = a;
            if (threadIdx.x == 0) { sm.prof[6] += clock64() - tr0; sm.prof[7] += 1; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
    if (threadIdx.x == 0) sm.prof[2] += clock64() - tg0;
    return cons;
}

// block-wide (256 consumer threads) reductions; result broadcast
__device__ __forceinline__ float consumer_sum(float v) {
    MegaSmem &sm = SM;
    v = warp_sum(v);
    consumer_sync();
    if ((threadIdx.x & 31) == 0) sm.red1[threadIdx.x >> 5] = v;
    consumer_sync();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) t += sm.red1[i];
    return t;
}
__device__ __forceinline__ float consumer_max(float v) {
    MegaSmem &sm = SM;
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
__device__ __noinline__ void ln_inplace(int K) {
    MegaSmem &sm = SM;
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int i = tid; i < K; i += kConsumerThreads) s += sm.xs[i];
    const float mean = consumer_sum(s) / K;
    float s2 = 0.f;
    for (int i = tid; i < K; i += kConsumerThreads) { const float v = sm.xs[i] - mean; sm.xs[i] = v; s2 += v * v; }
    const float var = consumer_sum(s2) / K;
    const float scale = rsqrtf(var + 1e-5f);
    for (int i = tid; i < K; i += kConsumerThreads) sm.xs[i] = r16(sm.xs[i] * scale * sm.lnw[i] + sm.lnb[i]);
    consumer_sync();
}

// n floats (multiple of 4) from L2 (ld.global.cg: produced by other CTAs) into sm.xs with all loads of a
// thread in flight at once: one L2 round trip on the critical path.
__device__ __noinline__ void load_xs(const float *g, int n) {
    MegaSmem &sm = SM;
    float4 v[kMaxJ];
    const int n4 = n >> 2;
#pragma unroll
    for (int j = 0; j < kMaxJ; j++) { const int i = threadIdx.x + j * kConsumerThreads; v[j] = i < n4 ? __ldcg(reinterpret_cast<const float4 *>(g) + i) : make_float4(0, 0, 0, 0); }
#pragma unroll
    for (int j = 0; j < kMaxJ; j++) { const int i = threadIdx.x + j * kConsumerThreads; if (i < n4) reinterpret_cast<float4 *>(sm.xs)[i] = v[j]; }
}

// token embedding + positional embedding -> sm.xs (every CTA) and the residual stream in L2 (CTA 0)
__device__ __noinline__ void embed_xs() {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int d = P.d, tid = threadIdx.x;
    const __half *e = P.tok_emb + (size_t)sm.st.token * d;
    const float *pe = P.d_pos + (size_t)sm.st.pos * d;
    float v[kMaxJ];
#pragma unroll
    for (int j = 0; j < kMaxJ; j++) { const int i = tid + j * kConsumerThreads; v[j] = i < d ? __half2float(__ldg(e + i)) + __ldg(pe + i) : 0.f; }
#pragma unroll
    for (int j = 0; j < kMaxJ; j++) { const int i = tid + j * kConsumerThreads; if (i < d) { sm.xs[i] = v[j]; if (blockIdx.x == 0) P.x[i] = v[j]; } }
}

// scores of one query against n key rows (128 B each) at `K`: 8 lanes per row, 4 rows per warp per step.
// FROM_RING: rows come from a ring slot (shared memory), else from L2 with 4 rows per thread in flight.
template <bool FROM_RING>
__device__ __forceinline__ float attn_scores(const uint8_t *K, int n, int sc_base, const float4 &qa, const float4 &qb, float lmax) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    for (int jb = 0; jb < n; jb += 128) {
        uint4 kv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) kv[u] = FROM_RING ? *reinterpret_cast<const uint4 *>(K + (size_t)j * 128 + l8 * 16) : __ldcg(reinterpret_cast<const uint4 *>(K + (size_t)j * 128 + l8 * 16));
            else kv[u] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            float ds = dot8(kv[u], qa, qb, 0.f);
            ds += __shfl_xor_sync(0xffffffffu, ds, 1);
            ds += __shfl_xor_sync(0xffffffffu, ds, 2);
            ds += __shfl_xor_sync(0xffffffffu, ds, 4);
            if (j < n) { if (l8 == 0) sm.sc[sc_base + j] = ds; lmax = fmaxf(lmax, ds); }
        }
    }
    return lmax;
}
template <bool FROM_RING>
__device__ __forceinline__ void attn_pv(const uint8_t *V, int n, int sc_base, float (&acc)[8]) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    for (int jb = 0; jb < n; jb += 128) {
        uint4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) vv[u] = FROM_RING ? *reinterpret_cast<const uint4 *>(V + (size_t)j * 128 + l8 * 16) : __ldcg(reinterpret_cast<const uint4 *>(V + (size_t)j * 128 + l8 * 16));
            else vv[u] = make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) axpy8(vv[u], sm.sc[sc_base + j], acc);
        }
    }
}
// fold the per-lane P.V partial sums of the CTA into 64 channel sums (thread c < 64 returns channel c)
__device__ __forceinline__ float attn_fold(float (&acc)[8]) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8); acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16); }
    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) sm.red[warp][l8 * 8 + i] = acc[i];
    }
    consumer_sync();
    float o = 0.f;
    if (threadIdx.x < 64) for (int w = 0; w < kConsumerWarps; w++) o += sm.red[w][threadIdx.x];
    return o;
}

// P1: self-attention, one CTA per head, K/V straight from L2 (written by this kernel: ld.global.cg)
__device__ __noinline__ void self_attn(int il) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int h = blockIdx.x;
    if (h >= P.H) return;
    const int n = sm.st.pos + 1, d = P.d, l8 = threadIdx.x & 7;
    const uint8_t *Kh = reinterpret_cast<const uint8_t *>(P.self_k + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    const uint8_t *Vh = reinterpret_cast<const uint8_t *>(P.self_v + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    const float4 qa = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8));
    const float4 qb = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8 + 4));
    const float lmax = attn_scores<false>(Kh, n, 0, qa, qb, -INFINITY);
    const float m = consumer_max(lmax);
    float lsum = 0.f;
    for (int j = threadIdx.x; j < n; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; lsum += e; }
    const float l = consumer_sum(lsum);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    attn_pv<false>(Vh, n, 0, acc);
    const float o = attn_fold(acc);
    if (threadIdx.x < 64) P.att[h * 64 + threadIdx.x] = r16(o / l);
}

// P4: cross-attention; K then V streamed through the ring; the last split of a head to finish combines the
// head's partials into the attention vector
__device__ __noinline__ uint32_t cross_attn(int il, uint32_t cons) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const Seg sk = make_seg(SEG_XK, il);
    if (sk.n_chunks == 0) return cons;
    const int tid = threadIdx.x, lane = tid & 31, l8 = tid & 7;
    const int ns = P.xsplit, h = blockIdx.x / ns, sp = blockIdx.x % ns, n = sk.rows;
    const float4 qa = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8));
    const float4 qb = __ldcg(reinterpret_cast<const float4 *>(P.q + h * 64 + l8 * 8 + 4));
    float lmax = -INFINITY;
    for (int ch = 0; ch < sk.n_chunks; ch++) {
        const int slot = cons % kSlots;
        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const int kbase = ch * sk.rows_per_chunk, nk = min(sk.rows_per_chunk, n - kbase);
        lmax = attn_scores<true>(sm.ring[slot], nk, kbase, qa, qb, lmax);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
    const float m = consumer_max(lmax);
    float lsum = 0.f;
    for (int j = tid; j < n; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; lsum += e; }
    const float l = consumer_sum(lsum);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int ch = 0; ch < sk.n_chunks; ch++) {     // the V slice has the same chunking as the K slice
        const int slot = cons % kSlots;
        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const int kbase = ch * sk.rows_per_chunk, nk = min(sk.rows_per_chunk, n - kbase);
        attn_pv<true>(sm.ring[slot], nk, kbase, acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
    const float o = attn_fold(acc);
    float *out = P.part + ((size_t)h * ns + sp) * 66;
    if (tid < 64) out[2 + tid] = o;
    if (tid == 0) { out[0] = m; out[1] = l; }
    consumer_sync();
    if (tid == 0) {
        __threadfence();
        const unsigned int old = atomicAdd(P.bar + 1 + h, 1u);
        sm.redi[24] = (old == sm.xuse * (unsigned int)ns + (unsigned int)ns - 1u) ? 1 : 0;
    }
    consumer_sync();
    if (sm.redi[24] && tid < 64) {   // last split of head h: fold all ns partials (<= 8)
        const float *p = P.part + (size_t)h * ns * 66;
        float pm[8], pl[8], po[8];
#pragma unroll
        for (int s2 = 0; s2 < 8; s2++) {
            if (s2 < ns) { pm[s2] = __ldcg(p + s2 * 66); pl[s2] = __ldcg(p + s2 * 66 + 1); po[s2] = __ldcg(p + s2 * 66 + 2 + tid); }
            else { pm[s2] = -INFINITY; pl[s2] = 0.f; po[s2] = 0.f; }
        }
        float M = -INFINITY;
#pragma unroll
        for (int s2 = 0; s2 < 8; s2++) M = fmaxf(M, pm[s2]);
        float Lsum = 0.f, oo = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < 8; s2++) if (pm[s2] > -INFINITY) { const float e = __expf(pm[s2] - M); Lsum += pl[s2] * e; oo += po[s2] * e; }
        P.att[h * 64 + tid] = r16(oo / Lsum);
    }
    return cons;
}

struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

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

// LM-head epilogue: publish raw logits; when sampling, the logits filter (whisper_process_logits) and this
// CTA's softmax statistics {max text, argmax, max timestamp, argmax, sum exp, sum exp over timestamps}
__device__ __noinline__ void lm_epilogue(int rows, int row0, bool keep, bool sampling) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DecState st = sm.st;
    MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
    for (int R = tid; R < rows; R += kConsumerThreads) {
        const int i = row0 + R;
        const float raw = sm.acc[R];
        P.logits[i] = raw;
        if (keep) P.keep[(size_t)st.n_kept * P.n_vocab + i] = raw;
        float x = -INFINITY;
        if (sampling && !token_masked(P, st, i)) x = raw;
        sm.acc[R] = x;
        if (i < P.beg) { if (x > mt.v) mt = MaxIdx{x, i}; } else { if (x > ms.v) ms = MaxIdx{x, i}; }
    }
    if (!sampling) return;
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
    for (int R = tid; R < rows; R += kConsumerThreads) {
        const float x = sm.acc[R];
        if (x > -INFINITY) { sa += expf(x - m_all); if (row0 + R >= P.beg) sb += expf(x - ms.v); }
    }
    sa = consumer_sum(sa);
    sb = consumer_sum(sb);
    if (tid == 0) {
        float *rec = P.stats + (size_t)blockIdx.x * 8;
        rec[0] = mt.v; rec[1] = __int_as_float(mt.i); rec[2] = ms.v; rec[3] = __int_as_float(ms.i); rec[4] = sa; rec[5] = sb;
    }
}

// combine the per-CTA records (every CTA, identically), greedy sample (whisper_sample_token best=true) and
// the per-token decoder bookkeeping of whisper_full; updates sm.st
__device__ __noinline__ void sample_and_update(int seek, int seek_end, int n_max) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int tid = threadIdx.x, lane = tid & 31, ncta = gridDim.x;
    if (tid < 32) {
        float r0[5], r1[5], r2[5], r3[5], r4[5], r5[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int c = lane + 32 * k;
            if (c < ncta) {
                const float4 a = __ldcg(reinterpret_cast<const float4 *>(P.stats + (size_t)c * 8));
                const float2 b = __ldcg(reinterpret_cast<const float2 *>(P.stats + (size_t)c * 8 + 4));
                r0[k] = a.x; r1[k] = a.y; r2[k] = a.z; r3[k] = a.w; r4[k] = b.x; r5[k] = b.y;
            } else { r0[k] = -INFINITY; r1[k] = __int_as_float(0x7fffffff); r2[k] = -INFINITY; r3[k] = __int_as_float(0x7fffffff); r4[k] = 0.f; r5[k] = 0.f; }
        }
        MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
#pragma unroll
        for (int k = 0; k < 5; k++) { mt = better(mt, MaxIdx{r0[k], __float_as_int(r1[k])}); ms = better(ms, MaxIdx{r2[k], __float_as_int(r3[k])}); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
            MaxIdx b{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, b);
        }
        const float m_all = fmaxf(mt.v, ms.v);
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float cm = fmaxf(r0[k], r2[k]), cs = r2[k];
            if (cm > -INFINITY) sa += r4[k] * expf(cm - m_all);
            if (cs > -INFINITY) sb += r5[k] * expf(cs - ms.v);
        }
        sa = warp_sum(sa); sb = warp_sum(sb);
        if (lane == 0) {
            DecState st = sm.st;
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
            if (blockIdx.x == 0) P.tok_out[i] = tk;
            int f = 0, cpl = 0;
            if (tk.id > P.beg) {
                const int sd_new = 2 * (tk.id - P.beg);
                if (st.has_ts && st.seek_delta > sd_new && st.result_len < i) f = 1;
                else { st.seek_delta = sd_new; st.result_len = i + 1; st.has_ts = 1; }
            }
            if (!f) {
                if (tk.id == P.eot || (st.has_ts && seek + st.seek_delta + 100 >= seek_end)) {
                    if (st.result_len == 0) { if (seek + st.seek_delta + 100 >= seek_end) st.result_len = i + 1; else f = 1; }
                    if (!f) cpl = 1;
                }
            }
            if (!f && !cpl && i == n_max - 1 && (st.result_len == 0 || st.seek_delta < 100 * kChunkSec / 2)) f = 1;
            st.failed = f; st.completed = cpl;
            st.penult_id = st.last_id; st.last_id = tk.id; st.n_sampled = i + 1;
            if (f || cpl || st.n_sampled >= n_max) st.done = 1;
            else { st.token = tk.id; st.pos = st.pos + 1; }
            sm.st = st;
        }
    }
    consumer_sync();
}

}  // namespace

// ================================================================================================
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaParams *__restrict__ Pp, int max_steps) {
    MegaSmem &sm = SM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(Pp);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.P);
        for (int i = tid; i < (int)(sizeof(MegaParams) / 4); i += kMegaThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const MegaParams &P = sm.P;
    const int d = P.d, L = P.L;

    DecCtl *ctl = P.ctl;
    // snapshot of the control block (identical in every CTA)
    const int pos_start = ctl->pos, pos0 = ctl->pos0, n_prompt = ctl->n_prompt, do_sample = ctl->sample;
    const int seek = ctl->seek, seek_end = ctl->seek_end, n_max = ctl->n_max, keep_logits = ctl->keep_logits, all_logits = ctl->all_logits;
    const int already_done = ctl->done;
    const int steps_left = max_steps;

    if (tid == 0) {
        for (int s = 0; s < kSlots; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps); }
        sm.stop_req = 0; sm.prod_done = 0; sm.issued = 0; sm.bar_target = 0; sm.xuse = 0;
        for (int i = 0; i < 8; i++) sm.prof[i] = 0;
        DecState st;
        st.pos = pos_start; st.token = ctl->token; st.n_sampled = ctl->n_sampled; st.has_ts = ctl->has_ts; st.seek_delta = ctl->seek_delta;
        st.result_len = ctl->result_len; st.last_id = ctl->last_id; st.penult_id = ctl->penult_id; st.n_kept = ctl->n_kept;
        st.failed = 0; st.completed = 0; st.done = 0;
        sm.st = st;
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
                    const Seg seg = make_seg(kind, layer < L ? layer : 0);
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
    const long long t_begin = clock64();
    prefetch_static(PH_QKV, 0);
    asm volatile("cp.async.wait_all;" ::: "memory");
    consumer_sync();

    for (int t = 0; t < steps_left; t++) {
        if (sm.st.done) break;
        const int jrel = sm.st.pos - pos0;
        const bool need_logits = all_logits || jrel >= n_prompt - 1;

#pragma unroll 1
        for (int il = 0; il < L; il++) {
#pragma unroll 1
            for (int ph = 0; ph < 8; ph++) {
                if (ph == PH_SELF) {
                    self_attn(il);
                } else if (ph == PH_CROSS) {
                    cons = cross_attn(il, cons);
                    if (tid == 0) sm.xuse = sm.xuse + 1;
                } else {
                    const int kind = ph == PH_QKV ? SEG_QKV : ph == PH_O ? SEG_O : ph == PH_CQ ? SEG_CQ : ph == PH_CO ? SEG_CO : ph == PH_FC1 ? SEG_FC1 : SEG_FC2;
                    const Seg seg = make_seg(kind, il);
                    const bool resid = ph == PH_O || ph == PH_CO || ph == PH_FC2;
                    float xres = 0.f;
                    if (resid && tid < seg.rows) xres = __ldcg(P.x + seg.row0 + tid);   // residual rows of this CTA, in flight early
                    if (ph == PH_QKV && il == 0) embed_xs();
                    else load_xs(resid ? (ph == PH_FC2 ? P.h : P.att) : P.x, ph == PH_FC2 ? 4 * d : d);
                    consumer_sync();
                    if (!resid) ln_inplace(d);
                    cons = gemv_rows(cons, seg.rows, seg.row_bytes, seg.rows_per_chunk, seg.n_chunks, d, ph == PH_FC2 ? 4 : 1);
                    consumer_sync();
                    if (ph == PH_QKV) {
                        const int pos = sm.st.pos;
                        __half *sk = P.self_k + (size_t)il * P.ctx * d, *sv = P.self_v + (size_t)il * P.ctx * d;
                        for (int R = tid; R < seg.rows; R += kConsumerThreads) {
                            const int row = seg.row0 + R;
                            const float v = sm.acc[R] + sm.bias[R];
                            if (row < d) P.q[row] = r16(v * P.s4);
                            else if (row < 2 * d) { const int n = row - d; sk[((size_t)(n >> 6) * P.ctx + pos) * 64 + (n & 63)] = __float2half_rn(v * P.s4); }
                            else { const int n = row - 2 * d; sv[((size_t)(n >> 6) * P.ctx + pos) * 64 + (n & 63)] = __float2half_rn(v); }
                        }
                    } else if (tid < seg.rows) {
                        const int row = seg.row0 + tid;
                        if (ph == PH_FC2) P.x[row] = xres + (sm.acc[4 * tid] + sm.acc[4 * tid + 1] + sm.acc[4 * tid + 2] + sm.acc[4 * tid + 3]) + sm.bias[tid];
                        else if (resid) P.x[row] = xres + sm.acc[tid] + sm.bias[tid];
                        else if (ph == PH_CQ) P.q[row] = r16((sm.acc[tid] + sm.bias[tid]) * P.s4);
                        else P.h[row] = gelu16(sm.acc[tid] + sm.bias[tid]);
                    }
                }
                int nph = ph + 1, nil = il;
                if (ph == PH_FC2) { nil = il + 1 < L ? il + 1 : 0; nph = il + 1 < L ? PH_QKV : (need_logits ? PH_LM : PH_QKV); }
                grid_sync(nph, nil);
            }
        }

        // ---------------- final LN + LM head + per-CTA softmax statistics ----------------
        if (need_logits) {
            load_xs(P.x, d);
            consumer_sync();
            ln_inplace(d);
            const Seg seg = make_seg(SEG_LM, 0);
            cons = gemv_rows(cons, seg.rows, seg.row_bytes, seg.rows_per_chunk, seg.n_chunks, d, 1);
            consumer_sync();
            const bool keep = keep_logits && sm.st.n_kept < P.keep_cap;
            lm_epilogue(seg.rows, seg.row0, keep, do_sample && jrel >= n_prompt - 1);
            grid_sync(PH_QKV, 0);
            if (keep) { consumer_sync(); if (tid == 0) sm.st.n_kept = sm.st.n_kept + 1; consumer_sync(); }
        }
        if (jrel < n_prompt - 1) {   // prompt token: feed the next one
            consumer_sync();
            if (tid == 0) { sm.st.token = ctl->prompt[jrel + 1]; sm.st.pos = sm.st.pos + 1; }
            consumer_sync();
            continue;
        }
        if (!do_sample) { consumer_sync(); if (tid == 0) sm.st.done = 1; consumer_sync(); break; }
        sample_and_update(seek, seek_end, n_max);
    }

    // ---------------- shutdown: stop the producer, drain copies still in flight, publish the state ----------------
    if (tid == 0) {
        sm.stop_req = 1;
        while (!sm.prod_done) {}
        __threadfence_block();
        const uint32_t issued = sm.issued;
        for (uint32_t c = cons; c < issued; c++) mbar_wait(&sm.full[c % kSlots], (c / kSlots) & 1);
        const DecState st = sm.st;
        if (P.prof) {
            sm.prof[3] = clock64() - t_begin;
            for (int i = 0; i < 8; i++) P.prof[(size_t)cta * 8 + i] = sm.prof[i];
        }
        if (cta == 0) {
            ctl->pos = st.pos; ctl->token = st.token; ctl->n_sampled = st.n_sampled; ctl->has_ts = st.has_ts; ctl->seek_delta = st.seek_delta;
            ctl->result_len = st.result_len; ctl->last_id = st.last_id; ctl->penult_id = st.penult_id; ctl->n_kept = st.n_kept;
            ctl->failed = st.failed; ctl->completed = st.completed; ctl->done = st.done;
        }
    }
}

// ------------------------------------------------------------------------------------------------
size_t decode_mega_smem_bytes() { return sizeof(MegaSmem); }

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
    CUDA_CHECK(cudaMemsetAsync(d_bar, 0, kMegaBarWords * sizeof(unsigned int), st));
    void *args[] = {(void *)&d_params, (void *)&max_steps};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)decode_mega_kernel, dim3(grid), dim3(kMegaThreads), args, decode_mega_smem_bytes(), st));
}

}  // namespace ss
