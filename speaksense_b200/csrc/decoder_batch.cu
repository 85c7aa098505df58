// decoder_batch.cu - one decode step for a BATCH of independent sequences (ss_transcribe_batch; BASELINE.json
// configs 3/4: 32 clips per GPU).  SURVEY.md §8d "decoder step, batch B": 1.6008 GB of weights + B x 0.24576 GB of
// cross-KV per step, so the weights must be streamed ONCE per step for the whole batch - the batch-1 persistent kernel
// (decoder_mega.cu) streams them once per sequence.
//
// Shape of the work: M = B <= 32 tokens against [N][K] f16 weight matrices.  That is far too skinny for the 128-row
// tcgen05 tiles of gemm_sm100.cu (10 CTAs for N = 1280), and it is HBM-bound anyway, so the mat-muls run on
// mma.m16n8k16 with the WEIGHTS as the A operand (16 output features per warp tile) and the sequences on the 8-wide N
// dimension (1, 2 or 4 n-tiles).  Fragments are loaded straight from global memory as 16-byte vectors: thread (g, t) of
// a warp reads halves [8t, 8t+8) of a 32-wide K block of weight rows g and g+8 and of x rows g (+8 per n-tile).  The MMA
// then sees K in a permuted order - the same permutation on both operands, so every product is still formed exactly
// once - and each row's 64 bytes per K block are read by 4 adjacent lanes (whole 32-byte sectors, no shared-memory
// staging, no ldmatrix).  Warps of a CTA split N (WR row tiles) and K (WK slices, folded through shared memory).
//
// Arithmetic is the oracle's and decoder_mega.cu's: f16 weights, activations rounded to f16 in front of every mat-mul,
// f32 accumulation, f32 LayerNorm / softmax, ggml's f16 GELU; Q and K carry head_dim^-1/4 each.  The logits filter,
// greedy sampling and whisper_full's per-token bookkeeping are restated per sequence in bd_sample_kernel
// (== lm_epilogue + sample_and_update of decoder_mega.cu).
//
// Measured (round 1, tools/batch_bench.py, profiles/r1f_batch_bench_pdl.json): 32 x 30 s clips, large-v3, one B200: 3.45 ms per
// step (9.8 GB of weights + caches = 44 % of the HBM peak; the cross-attention alone runs at 86 %), RTF 1559x against 292x
// clip by clip, results identical (tests/test_gpu_batch.py).  SS_BATCH_DECODE=0 gives ss_transcribe_batch the clip-by-clip
// decode back, SS_BATCH_PDL=0 plain stream-ordered launches.
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace ss {

namespace {

enum : int { EPI_QKV = 0, EPI_RES, EPI_Q, EPI_GELU, EPI_LOGITS };

// Programmatic dependent launch: every kernel of a step lets its successor start at once (launch_dependents) and itself
// waits for its predecessor (wait) only where it first touches something the predecessor wrote.  What a GEMM does before
// the wait - issuing the loads of its first weight fragments, which are static - overlaps the predecessor's execution, and
// every kernel's launch latency and ramp-up hide behind the previous one.  Without the launch attribute both are no-ops.
// (pdl_trigger / pdl_wait: kernels.h)

// Loads whose position in the instruction stream matters (issued right behind the wait, BEFORE the branch on a sequence's finished
// flag, so that no L2 round trip waits for another): volatile asm keeps program order against the wait and against each other.
__device__ __forceinline__ int ldp_s32(const int *p) { int v; asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ldp_f32(const float *p) { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ unsigned short ldp_u16(const void *p) { unsigned short v; asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p)); return v; }
template <bool STREAM>
__device__ __forceinline__ uint4 ldp_v4(const uint4 *p) {
    uint4 v;
    if (STREAM) asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

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
// block-wide reductions over NW warps; `scratch` holds NW floats and must not be in use (callers alternate two rows)
template <int NW>
__device__ __forceinline__ float block_sum(float v, float *scratch) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NW; i++) t += scratch[i];
    return t;
}
template <int NW>
__device__ __forceinline__ float block_max(float v, float *scratch) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = scratch[0];
#pragma unroll
    for (int i = 1; i < NW; i++) t = fmaxf(t, scratch[i]);
    return t;
}
__device__ __forceinline__ float2 h2f(uint32_t u) { return __half22float2(*reinterpret_cast<__half2 *>(&u)); }
__device__ __forceinline__ float dot8(const uint4 &w, const float4 &a, const float4 &b) {
    float acc = 0.f;
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
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------------
// LayerNorm of every live sequence's residual row -> f16 operand (ggml_norm: mean, then the variance of the centred
// values).  embed: the row is first formed as token embedding + positional embedding (start of a step).
// ------------------------------------------------------------------------------------------------
constexpr int kLnThreads = 256;
constexpr int kLnPer = 5;     // d <= 1280
__global__ void __launch_bounds__(kLnThreads) bd_ln_kernel(const __grid_constant__ BatchParams P, const float *__restrict__ lw,
                                                           const float *__restrict__ lb, int embed) {
    __shared__ float red[2][8];
    const int b = blockIdx.x, tid = threadIdx.x, d = P.d;
    pdl_trigger();
    float w[kLnPer], bb[kLnPer];      // the affine parameters are static: fetched before the wait
#pragma unroll
    for (int k = 0; k < kLnPer; k++) { const int i = min(tid + k * kLnThreads, d - 1); w[k] = __ldg(lw + i); bb[k] = __ldg(lb + i); }
    pdl_wait();
    const DecCtl *ctl = P.seq[b].ctl;
    float *x = P.x + (size_t)b * d;
    float v[kLnPer];
    if (!embed) {      // the row is requested together with the finished flag (one L2 round trip, not two)
#pragma unroll
        for (int k = 0; k < kLnPer; k++) v[k] = ldp_f32(x + min(tid + k * kLnThreads, d - 1));
    }
    const int done = ldp_s32(&ctl->done), tok = ldp_s32(&ctl->token), pos = ldp_s32(&ctl->pos);
    if (done) return;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kLnPer; k++) {
        const int i = tid + k * kLnThreads;
        if (i < d) {
            if (embed) { v[k] = __half2float(P.tok_emb[(size_t)tok * d + i]) + P.d_pos[(size_t)pos * d + i]; x[i] = v[k]; }
            s += v[k];
        } else v[k] = 0.f;
    }
    const float mean = block_sum<8>(s, red[0]) / (float)d;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLnPer; k++) if (tid + k * kLnThreads < d) { const float c = v[k] - mean; s2 += c * c; }
    const float var = block_sum<8>(s2, red[1]) / (float)d;
    const float scale = 1.0f / sqrtf(var + 1e-5f);
    __half *y = P.xn + (size_t)b * d;
#pragma unroll
    for (int k = 0; k < kLnPer; k++) {
        const int i = tid + k * kLnThreads;
        if (i < d) y[i] = __float2half_rn((v[k] - mean) * scale * w[k] + bb[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// skinny GEMM: out[b][n] = sum_k W[n][k] * X[b][k]  (+ epilogue), b < 8 * NT
// ------------------------------------------------------------------------------------------------
// `bias` already added by the caller where the kind has one; `res` = the prefetched residual (EPI_RES)
template <int EPI>
__device__ __forceinline__ void bd_epilogue(const BatchParams &P, int il, int row, int b, float v, float res) {
    const int d = P.d;
    if (EPI == EPI_LOGITS) P.logits[(size_t)b * P.n_vocab + row] = v;
    else if (EPI == EPI_QKV) {
        if (row < d) P.q[(size_t)b * d + row] = __float2half_rn(v * P.s4);
        else {
            const bool is_k = row < 2 * d;
            const int n = row - (is_k ? d : 2 * d), pos = P.seq[b].ctl->pos;
            __half *cache = (is_k ? P.seq[b].self_k : P.seq[b].self_v) + (size_t)il * P.ctx * d;
            cache[((size_t)(n >> 6) * P.ctx + pos) * 64 + (n & 63)] = __float2half_rn(is_k ? v * P.s4 : v);
        }
    } else if (EPI == EPI_RES) P.x[(size_t)b * d + row] = res + v;
    else if (EPI == EPI_Q) P.q[(size_t)b * d + row] = __float2half_rn(v * P.s4);
    else if (EPI == EPI_GELU) P.hid[(size_t)b * 4 * d + row] = __float2half_rn(gelu16(v));
}

constexpr int kGemmThreads = 256;
constexpr int kGemmU = 5;      // 32-wide K blocks per warp and round (5 KB of weights in registers): K = 1280 split 8 ways is one round
constexpr int kGemmXD = 3;     // x blocks in flight per warp (registers; the weights of the NEXT round refill a block's registers as soon as
                               // its products are issued, so neither operand's latency is paid once per block)
// Latency structure (round 2, from the SASS of the first version: one L2 round trip for the finished counter, then one per K block
// because the tail guard put every block's x loads behind the previous block's MMAs, then one HBM round trip per extra round):
// everything the kernel needs from its predecessor is now requested right behind the wait - the counter, the live flags, the
// residual, the first kGemmXD blocks of x - and the body of a round is branch-free (blocks past the end of a short tail multiply
// zeroed weights), so a round costs about two dependent L2 round trips instead of six.
template <int WR, int WK, int NT, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 2) bd_gemm_kernel(const __grid_constant__ BatchParams P, const __half *__restrict__ W,
                                                                  const float *__restrict__ bias, const __half *__restrict__ X, int N, int K, int il) {
    static_assert(WR * WK == 8, "8 warps per CTA");
    constexpr int U = kGemmU, XD = kGemmXD < kGemmU ? kGemmXD : kGemmU;
    constexpr int TOTAL = WR * 16 * 8 * NT;                                   // outputs of the CTA
    constexpr int NOUT = (TOTAL + kGemmThreads - 1) / kGemmThreads;           // per thread in the epilogue
    __shared__ float red[WR][WK][8 * NT][17];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wr = warp / WK, wk = warp % WK;
    const int row_base = (blockIdx.x * WR + wr) * 16;
    const int nblk = K >> 5, blk0 = wk * nblk / WK, blk1 = (wk + 1) * nblk / WK;      // 32-wide K blocks of this warp
    const int ra = min(row_base + g, N - 1), rb = min(row_base + g + 8, N - 1);       // clamped: rows past N are computed, not stored
    const uint4 *wa = reinterpret_cast<const uint4 *>(W + (size_t)ra * K) + t;
    const uint4 *wb = reinterpret_cast<const uint4 *>(W + (size_t)rb * K) + t;
    const uint4 *xp = reinterpret_cast<const uint4 *>(X + (size_t)g * K) + t;         // n-tile nt: + nt * K (8 rows of K / 8 vectors)
    pdl_trigger();
    uint4 a[U], c[U];
#pragma unroll
    for (int u = 0; u < U; u++) {      // first round of weight fragments: static data, in flight while the predecessor still runs
        const int bi = min(blk0 + u, blk1 - 1);      // a short tail re-reads the last block; its products are zeroed below
        a[u] = __ldcs(wa + bi * 4); c[u] = __ldcs(wb + bi * 4);
    }
    float pb[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        const int idx = tid + k * kGemmThreads;
        const int row = (blockIdx.x * WR + idx / (16 * 8 * NT)) * 16 + (idx & 15);
        pb[k] = (EPI != EPI_LOGITS && idx < TOTAL && row < N) ? __ldg(bias + row) : 0.f;
    }
    pdl_wait();
    // ---- what the predecessor wrote, requested all at once
    uint4 xq[XD][NT];
#pragma unroll
    for (int u = 0; u < XD; u++) {
        const int bi = min(blk0 + u, blk1 - 1);
#pragma unroll
        for (int nt = 0; nt < NT; nt++) xq[u][nt] = __ldg(xp + (size_t)nt * K + bi * 4);
    }
    const int n_done = *P.n_done;      // (all sequences finished: nothing is stored - tested after the tiles, so that no load waits for it)
    bool live[2];
    float pr[NOUT];
#pragma unroll
    for (int k = 0; k < 2; k++) {      // a thread's outputs alternate between two sequences at most (512 / 16 is a multiple of 8 * NT)
        const int b = ((tid + k * kGemmThreads) >> 4) % (8 * NT);
        live[k] = b < P.B && (EPI == EPI_LOGITS || !P.seq[b].ctl->done);
    }
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        const int idx = tid + k * kGemmThreads;
        const int b = (idx >> 4) % (8 * NT), row = (blockIdx.x * WR + idx / (16 * 8 * NT)) * 16 + (idx & 15);
        pr[k] = (EPI == EPI_RES && idx < TOTAL && row < N && b < P.B) ? P.x[(size_t)b * P.d + row] : 0.f;
    }
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
    for (int blk = blk0; blk < blk1; blk += U) {
        const bool more = blk + U < blk1;      // (warp-uniform) another round follows
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool on = blk + u < blk1;    // blocks past the end of a short tail: zero weights, so the body has no branch
            const uint4 au = on ? a[u] : make_uint4(0u, 0u, 0u, 0u), cu = on ? c[u] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const uint4 xv = xq[u % XD][nt];
                mma16816(acc[nt], au.x, cu.x, au.y, cu.y, xv.x, xv.y);
                mma16816(acc[nt], au.z, cu.z, au.w, cu.w, xv.z, xv.w);
            }
            // refill slot u % XD: x of the block XD ahead in this round, else of the block of the NEXT round that uses this slot
            // (block u' < XD of a round sits in slot u'); the weights of the same block of the next round
            {
                const int bx = u + XD < U ? blk + u + XD : blk + U + (u % XD);
                if (bx < blk1) {      // (warp-uniform; nothing is fetched past the end)
#pragma unroll
                    for (int nt = 0; nt < NT; nt++) xq[u % XD][nt] = __ldg(xp + (size_t)nt * K + bx * 4);
                }
            }
            if (more) {
                const int bi = min(blk + U + u, blk1 - 1);
                a[u] = __ldcs(wa + bi * 4); c[u] = __ldcs(wb + bi * 4);
            }
        }
    }
    if (n_done >= P.B) return;
    // C fragment: c0/c1 = (row g, sequences 2t, 2t+1), c2/c3 = (row g + 8, ...)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
        red[wr][wk][nt * 8 + 2 * t][g] = acc[nt][0]; red[wr][wk][nt * 8 + 2 * t + 1][g] = acc[nt][1];
        red[wr][wk][nt * 8 + 2 * t][g + 8] = acc[nt][2]; red[wr][wk][nt * 8 + 2 * t + 1][g + 8] = acc[nt][3];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        const int idx = tid + k * kGemmThreads;
        const int r = idx & 15, b = (idx >> 4) % (8 * NT), w2 = idx / (16 * 8 * NT);
        if (idx >= TOTAL) break;
        float v = 0.f;
#pragma unroll
        for (int kk = 0; kk < WK; kk++) v += red[w2][kk][b][r];
        const int row = (blockIdx.x * WR + w2) * 16 + r;
        if (row < N && live[k & 1]) bd_epilogue<EPI>(P, il, row, b, v + pb[k], pr[k]);
    }
}

// ------------------------------------------------------------------------------------------------
// attention of one query (64 channels) over keys [0, n) of a head-major f16 cache: scores -> softmax statistics -> P.V
// 8 lanes per key row (16 bytes each), 4 keys per warp and step, 4 steps in flight.
// Returns in thread c < 64 the UNNORMALISED output channel c; m / l are the softmax maximum and sum.
// ------------------------------------------------------------------------------------------------
constexpr int kAttU = 4;
// the first batch of key and value rows (NW * 16 of each), requested before the query / the sequence's position are known: rows are
// clamped to the ALLOCATION (n_alloc rows exist behind Kh / Vh), not to the number of valid keys
template <int NW, bool STREAM>
__device__ __forceinline__ void attend_prefetch(const __half *__restrict__ Kh, const __half *__restrict__ Vh, int n_alloc, uint4 (&k0)[kAttU],
                                                uint4 (&v0)[kAttU], int swz_row0 = -1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
#pragma unroll
    for (int u = 0; u < kAttU; u++) {
        const int j = min(u * NW * 4 + warp * 4 + sub, n_alloc - 1);
        const int ch = swz_row0 >= 0 ? (l8 ^ ((swz_row0 + j) & 7)) : l8;
        k0[u] = ldp_v4<STREAM>(reinterpret_cast<const uint4 *>(Kh + (size_t)j * 64) + ch);
        v0[u] = ldp_v4<STREAM>(reinterpret_cast<const uint4 *>(Vh + (size_t)j * 64) + ch);
    }
}
// PRE: the first batch of rows comes from attend_prefetch (k0 / v0); otherwise the two arrays are ignored
template <int NW, bool STREAM, bool PRE>
__device__ __forceinline__ float attend(const __half *__restrict__ Kh, const __half *__restrict__ Vh, int n, const float *q, float *sc,
                                        float (*red)[64], float *red1, float &m_out, float &l_out, const uint4 (&k0)[kAttU],
                                        const uint4 (&v0)[kAttU], int swz_row0 = -1) {
    // swz_row0 >= 0: the rows are cross-KV cache rows starting at absolute key swz_row0 - chunk c of key m sits at c ^ (m & 7)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane >> 3, l8 = lane & 7;
    const float4 qa = *reinterpret_cast<const float4 *>(q + l8 * 8), qb = *reinterpret_cast<const float4 *>(q + l8 * 8 + 4);
    constexpr int STEP = NW * 4, U = kAttU;
    float lmax = -INFINITY;
    for (int jb = 0; jb < n; jb += STEP * U) {
        uint4 kv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (PRE && jb == 0) { kv[u] = k0[u]; continue; }      // (prefetched; rows >= n hold allocated, unused data and are skipped below)
            const int j = min(jb + u * STEP + warp * 4 + sub, n - 1);
            const uint4 *p = reinterpret_cast<const uint4 *>(Kh + (size_t)j * 64) + (swz_row0 >= 0 ? (l8 ^ ((swz_row0 + j) & 7)) : l8);
            kv[u] = STREAM ? __ldcs(p) : __ldcg(p);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + u * STEP + warp * 4 + sub;
            float ds = dot8(kv[u], qa, qb);
            ds += __shfl_xor_sync(0xffffffffu, ds, 1);
            ds += __shfl_xor_sync(0xffffffffu, ds, 2);
            ds += __shfl_xor_sync(0xffffffffu, ds, 4);
            if (j < n) { if (l8 == 0) sc[j] = ds; lmax = fmaxf(lmax, ds); }
        }
    }
    const float m = block_max<NW>(lmax, red1);          // (its barrier also publishes the scores)
    float lsum = 0.f;
    for (int j = tid; j < n; j += NW * 32) { const float e = __expf(sc[j] - m); sc[j] = e; lsum += e; }
    const float l = block_sum<NW>(lsum, red1 + NW);     // (its barrier also publishes the probabilities)
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int jb = 0; jb < n; jb += STEP * U) {
        uint4 vv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (PRE && jb == 0) { vv[u] = v0[u]; continue; }
            const int j = min(jb + u * STEP + warp * 4 + sub, n - 1);
            const uint4 *p = reinterpret_cast<const uint4 *>(Vh + (size_t)j * 64) + (swz_row0 >= 0 ? (l8 ^ ((swz_row0 + j) & 7)) : l8);
            vv[u] = STREAM ? __ldcs(p) : __ldcg(p);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = jb + u * STEP + warp * 4 + sub;
            if (j < n) axpy8(vv[u], sc[j], acc);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8); acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16); }
    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) red[warp][l8 * 8 + i] = acc[i];
    }
    __syncthreads();
    float o = 0.f;
    if (tid < 64) for (int w = 0; w < NW; w++) o += red[w][tid];
    m_out = m; l_out = l;
    return o;
}

// self-attention: grid (H, B); the current token's K / V are already in the cache (QKV epilogue), so n = pos + 1
constexpr int kSelfWarps = 4;
__global__ void __launch_bounds__(kSelfWarps * 32) bd_self_attn_kernel(const __grid_constant__ BatchParams P, int il) {
    __shared__ __align__(16) float q[64];
    __shared__ float sc[512];
    __shared__ float red[kSelfWarps][64];
    __shared__ float red1[2 * kSelfWarps];
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, d = P.d;
    pdl_trigger();
    const DecCtl *ctl = P.seq[b].ctl;
    const size_t off = (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64;
    pdl_wait();
    // one L2 round trip for everything: the first 64 key / value rows, the query, the finished flag and the position
    uint4 k0[kAttU], v0[kAttU];
    attend_prefetch<kSelfWarps, false>(P.seq[b].self_k + off, P.seq[b].self_v + off, P.ctx, k0, v0);
    const unsigned short qh = ldp_u16(P.q + (size_t)b * d + h * 64 + (tid & 63));
    const int done = ldp_s32(&ctl->done), n = ldp_s32(&ctl->pos) + 1;
    if (done) return;
    if (tid < 64) q[tid] = __half2float(__ushort_as_half(qh));
    __syncthreads();
    float m, l;
    const float o = attend<kSelfWarps, false, true>(P.seq[b].self_k + off, P.seq[b].self_v + off, n, q, sc, red, red1, m, l, k0, v0);
    if (tid < 64) P.att[(size_t)b * d + h * 64 + tid] = __float2half_rn(o / l);
}

// cross-attention over the T encoder positions: grid (H, B, S); S > 1 splits the keys and leaves {m, l, o[64]} records
constexpr int kCrossWarps = 8;
__global__ void __launch_bounds__(kCrossWarps * 32) bd_cross_attn_kernel(const __grid_constant__ BatchParams P, int il, int S) {
    __shared__ __align__(16) float q[64];
    __shared__ float sc[1536];
    __shared__ float red[kCrossWarps][64];
    __shared__ float red1[2 * kCrossWarps];
    const int h = blockIdx.x, b = blockIdx.y, sp = blockIdx.z, tid = threadIdx.x, d = P.d, T = P.T;
    pdl_trigger();
    const DecCtl *ctl = P.seq[b].ctl;
    const int per = (T + S - 1) / S, j0 = min(T, sp * per), n = min(T, j0 + per) - j0;
    const size_t off = (size_t)il * 2 * T * d + ((size_t)h * T + j0) * 64;
    // (no key / value prefetch here: 32 more registers would take the kernel from 5 to 3 CTAs per SM and B x H = 640 CTAs out of one
    //  wave - measured: 432 -> 522 ms of decode time for 32 clips)
    const uint4 k0[kAttU] = {}, v0[kAttU] = {};
    pdl_wait();
    const unsigned short qh = ldp_u16(P.q + (size_t)b * d + h * 64 + (tid & 63));
    const int done = ldp_s32(&ctl->done);
    if (done) return;
    if (tid < 64) q[tid] = __half2float(__ushort_as_half(qh));
    __syncthreads();
    float m = -INFINITY, l = 0.f, o = 0.f;
    if (n > 0) o = attend<kCrossWarps, true, false>(P.seq[b].cross_k + off, P.seq[b].cross_v + off, n, q, sc, red, red1, m, l, k0, v0, j0);
    if (S == 1) {
        if (tid < 64) P.att[(size_t)b * d + h * 64 + tid] = __float2half_rn(o / l);
    } else {
        float *rec = P.part + (((size_t)b * P.H + h) * S + sp) * 66;
        if (tid < 64) rec[2 + tid] = o;
        if (tid == 0) { rec[0] = m; rec[1] = l; }
    }
}
__global__ void __launch_bounds__(64) bd_cross_fold_kernel(const __grid_constant__ BatchParams P, int S) {
    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (P.seq[b].ctl->done) return;
    const float *rec = P.part + ((size_t)b * P.H + h) * S * 66;
    float M = -INFINITY;
    for (int s = 0; s < S; s++) M = fmaxf(M, rec[s * 66]);
    float L = 0.f, o = 0.f;
    for (int s = 0; s < S; s++) {
        const float pm = rec[s * 66];
        if (pm > -INFINITY) { const float e = __expf(pm - M); L += rec[s * 66 + 1] * e; o += rec[s * 66 + 2 + tid] * e; }
    }
    P.att[(size_t)b * P.d + h * 64 + tid] = __float2half_rn(o / L);
}

// ------------------------------------------------------------------------------------------------
// end of a step, one CTA per sequence: feed the next prompt token, or filter the logits (whisper_process_logits), take
// the greedy token (whisper_sample_token, best = true) and apply whisper_full's per-token bookkeeping
// ------------------------------------------------------------------------------------------------
struct MaxIdx { float v; int i; };
__device__ __forceinline__ MaxIdx better(MaxIdx a, MaxIdx b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

struct SeqState { int pos, pos0, token, done, n_sampled, has_ts, seek_delta, result_len, last_id, penult_id, n_prompt, seek, seek_end, n_max, sample; float temperature; };

__device__ __forceinline__ bool token_masked(const BatchParams &P, const SeqState &st, int i) {
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

constexpr int kSampleWarps = 32;
__global__ void __launch_bounds__(kSampleWarps * 32) bd_sample_kernel(const __grid_constant__ BatchParams P) {
    __shared__ SeqState S;
    __shared__ float rv[2][kSampleWarps];
    __shared__ int ri[2][kSampleWarps];
    __shared__ float rs[2][kSampleWarps];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DecCtl *ctl = P.seq[b].ctl;
    pdl_trigger();
    pdl_wait();
    if (tid == 0) {
        SeqState s;
        s.pos = ctl->pos; s.pos0 = ctl->pos0; s.token = ctl->token; s.done = ctl->done; s.n_sampled = ctl->n_sampled; s.has_ts = ctl->has_ts;
        s.seek_delta = ctl->seek_delta; s.result_len = ctl->result_len; s.last_id = ctl->last_id; s.penult_id = ctl->penult_id;
        s.n_prompt = ctl->n_prompt; s.seek = ctl->seek; s.seek_end = ctl->seek_end; s.n_max = ctl->n_max; s.sample = ctl->sample;
        s.temperature = ctl->temperature;
        S = s;
    }
    __syncthreads();
    const SeqState st = S;
    if (st.done) return;
    const int jrel = st.pos - st.pos0;
    if (jrel < st.n_prompt - 1) {      // prompt token: feed the next one
        if (tid == 0) { ctl->token = ctl->prompt[jrel + 1]; ctl->pos = st.pos + 1; }
        return;
    }
    if (!st.sample) {                  // teacher-forced run: the logits of the last token are the result
        if (tid == 0) { ctl->done = 1; atomicAdd(P.n_done, 1); }
        return;
    }
    const float *logits = P.logits + (size_t)b * P.n_vocab;
    const bool drawn = st.sample == 3;                                  // t > 0: whisper_sample_token(best = false)
    const float inv_div = st.temperature;                               // (the host divides: logits[i] /= temperature)
    auto value = [&](int i) -> float {
        if (token_masked(P, st, i)) return -INFINITY;
        return drawn ? logits[i] / inv_div : logits[i];
    };
    MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};      // best text token / best timestamp token
    for (int i = tid; i < P.n_vocab; i += kSampleWarps * 32) {
        const float x = value(i);
        if (i < P.beg) { if (x > mt.v) mt = MaxIdx{x, i}; } else { if (x > ms.v) ms = MaxIdx{x, i}; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
        MaxIdx c{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, c);
    }
    if (lane == 0) { rv[0][warp] = mt.v; ri[0][warp] = mt.i; rv[1][warp] = ms.v; ri[1][warp] = ms.i; }
    __syncthreads();
    mt = MaxIdx{rv[0][0], ri[0][0]}; ms = MaxIdx{rv[1][0], ri[1][0]};
    for (int w = 1; w < kSampleWarps; w++) { mt = better(mt, MaxIdx{rv[0][w], ri[0][w]}); ms = better(ms, MaxIdx{rv[1][w], ri[1][w]}); }
    const float m_all = fmaxf(mt.v, ms.v);
    float sa = 0.f, sb = 0.f;      // sum exp over everything (relative to m_all) / over the timestamps (relative to their maximum)
    for (int i = tid; i < P.n_vocab; i += kSampleWarps * 32) {
        const float x = value(i);
        if (x > -INFINITY) { sa += expf(x - m_all); if (i >= P.beg) sb += expf(x - ms.v); }
    }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane == 0) { rs[0][warp] = sa; rs[1][warp] = sb; }
    __syncthreads();
    sa = 0.f; sb = 0.f;
    for (int w = 0; w < kSampleWarps; w++) { sa += rs[0][w]; sb += rs[1][w]; }
    const float max_text = mt.v, max_ts = ms.v;
    const float lse = logf(sa) + m_all;
    const float ts_lp = sb > 0.f ? logf(sb) + (max_ts - lse) : -INFINITY;      // logsumexp of the timestamp log-probs
    const float text_lp = max_text - lse;
    int drawn_id = -1;
    if (drawn) {
        // std::discrete_distribution<>(probs)(rng) of libstdc++ (whisper_sample_token, best = false): the probabilities are
        // normalised by their sum in double, accumulated, and the first index whose cumulative probability is not below the
        // uniform wins.  Thread t owns the contiguous slice [t * per, (t + 1) * per); slice sums are scanned across the CTA.
        __shared__ double ds[kSampleWarps];
        __shared__ double dtot;
        __shared__ int dmin[kSampleWarps];
        const bool text_off = ts_lp > text_lp;                                  // the timestamps outweigh every text token
        auto prob = [&](int i) -> float {
            const float x = value(i);
            if (x == -INFINITY || (text_off && i < P.beg)) return 0.f;
            return expf(x - lse);
        };
        const int per = (P.n_vocab + kSampleWarps * 32 - 1) / (kSampleWarps * 32);
        const int i0 = tid * per, i1 = min(P.n_vocab, i0 + per);
        double loc = 0.0;
        for (int i = i0; i < i1; i++) loc += (double)prob(i);
        double incl = loc;                                                      // inclusive scan of the slice sums: warp, then CTA
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) ds[warp] = incl;
        __syncthreads();
        if (tid == 0) { double a = 0.0; for (int w = 0; w < kSampleWarps; w++) { const double t = ds[w]; ds[w] = a; a += t; } dtot = a; }
        __syncthreads();
        const double total = dtot;
        double c = (ds[warp] + incl - loc) / total;                             // cumulative probability in front of this slice
        const double r = ctl->u[st.n_sampled < kMaxDraws ? st.n_sampled : kMaxDraws - 1];
        int mine = 0x7fffffff;
        for (int i = i0; i < i1; i++) { c += (double)prob(i) / total; if (!(c < r)) { mine = i; break; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        if (lane == 0) dmin[warp] = mine;
        __syncthreads();
        mine = dmin[0];
        for (int w = 1; w < kSampleWarps; w++) mine = min(mine, dmin[w]);
        drawn_id = min(mine, P.n_vocab - 1);                                    // (rounding left the last cumulative value below r)
    }
    if (tid != 0) return;
    SeqState s = st;
    TokData tk;
    if (drawn) { tk.id = drawn_id; const float x = value(drawn_id); tk.plog = x - lse; }
    else if (ts_lp > text_lp) { tk.id = ms.i; tk.plog = max_ts - lse; }         // timestamps outweigh every text token
    else if (max_text >= max_ts) { tk.id = mt.i; tk.plog = text_lp; }
    else { tk.id = ms.i; tk.plog = max_ts - lse; }
    if (tk.id == 0x7fffffff) { tk.id = 0; tk.plog = -INFINITY; }
    tk.p = expf(tk.plog);
    const float p_ts_max = max_ts > -INFINITY ? expf(max_ts - lse) : 0.f;
    const float p_ts_sum = sb * p_ts_max;
    tk.tid = (max_ts > -INFINITY && p_ts_max > 0.f) ? ms.i : 0;
    tk.pt = p_ts_max / (p_ts_sum + 1e-10f); tk.ptsum = p_ts_sum;
    if (tk.id >= P.beg) { tk.tid = tk.id; tk.pt = tk.p; }
    const int i = s.n_sampled;
    P.seq[b].tok_out[i] = tk;
    int f = 0, cpl = 0;
    if (tk.id > P.beg) {
        const int sd_new = 2 * (tk.id - P.beg);
        if (s.has_ts && s.seek_delta > sd_new && s.result_len < i) f = 1;
        else { s.seek_delta = sd_new; s.result_len = i + 1; s.has_ts = 1; }
    }
    if (!f) {
        if (tk.id == P.eot || (s.has_ts && s.seek + s.seek_delta + 100 >= s.seek_end)) {
            if (s.result_len == 0) { if (s.seek + s.seek_delta + 100 >= s.seek_end) s.result_len = i + 1; else f = 1; }
            if (!f) cpl = 1;
        }
    }
    if (!f && !cpl && i == s.n_max - 1 && (s.result_len == 0 || s.seek_delta < 100 * kChunkSec / 2)) f = 1;
    s.penult_id = s.last_id; s.last_id = tk.id; s.n_sampled = i + 1;
    const int done = (f || cpl || s.n_sampled >= s.n_max) ? 1 : 0;
    if (!done) { s.token = tk.id; s.pos = s.pos + 1; }
    ctl->pos = s.pos; ctl->token = s.token; ctl->n_sampled = s.n_sampled; ctl->has_ts = s.has_ts; ctl->seek_delta = s.seek_delta;
    ctl->result_len = s.result_len; ctl->last_id = s.last_id; ctl->penult_id = s.penult_id; ctl->failed = f; ctl->completed = cpl;
    if (done) { ctl->done = 1; atomicAdd(P.n_done, 1); }
}

// ------------------------------------------------------------------------------------------------
// beam search (engine_batch.cc decode_beam_batched; SS_BATCH_BEAM=0 turns it off): end of a step whose
// sequences are the live beams of one window.  Same filter and statistics as bd_sample_kernel (whisper_process_logits, with
// the temperature division of the t > 0 rungs), then whisper_sample_token_topk: the k most likely tokens in (log-prob
// descending, id ascending) order - masked tokens stay at -inf, so a list that runs out of allowed tokens fills up with the
// lowest ids exactly like the host's partial sort (engine.cc sample_topk_host).  k <= 8 candidates per sequence are left in
// `cand`; the host does the beam bookkeeping and re-arms the control blocks for the next step.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSampleWarps * 32) bd_topk_kernel(const __grid_constant__ BatchParams P, float temperature, int K, TokData *cand) {
    __shared__ SeqState S;
    __shared__ float rv[2][kSampleWarps];
    __shared__ int ri[2][kSampleWarps];
    __shared__ float rs[2][kSampleWarps];
    __shared__ MaxIdx pick;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DecCtl *ctl = P.seq[b].ctl;
    pdl_trigger();
    pdl_wait();
    if (tid == 0) {
        SeqState s;
        s.pos = ctl->pos; s.pos0 = ctl->pos0; s.token = ctl->token; s.done = ctl->done; s.n_sampled = ctl->n_sampled; s.has_ts = ctl->has_ts;
        s.seek_delta = ctl->seek_delta; s.result_len = ctl->result_len; s.last_id = ctl->last_id; s.penult_id = ctl->penult_id;
        s.n_prompt = ctl->n_prompt; s.seek = ctl->seek; s.seek_end = ctl->seek_end; s.n_max = ctl->n_max; s.sample = ctl->sample;
        S = s;
    }
    __syncthreads();
    const SeqState st = S;
    if (st.done) return;
    const float *logits = P.logits + (size_t)b * P.n_vocab;
    const bool scaled = temperature > 0.0f;
    auto value = [&](int i) -> float {
        if (token_masked(P, st, i)) return -INFINITY;
        return scaled ? logits[i] / temperature : logits[i];
    };
    MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
    for (int i = tid; i < P.n_vocab; i += kSampleWarps * 32) {
        const float x = value(i);
        if (i < P.beg) { if (x > mt.v) mt = MaxIdx{x, i}; } else { if (x > ms.v) ms = MaxIdx{x, i}; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
        MaxIdx c{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, c);
    }
    if (lane == 0) { rv[0][warp] = mt.v; ri[0][warp] = mt.i; rv[1][warp] = ms.v; ri[1][warp] = ms.i; }
    __syncthreads();
    mt = MaxIdx{rv[0][0], ri[0][0]}; ms = MaxIdx{rv[1][0], ri[1][0]};
    for (int w = 1; w < kSampleWarps; w++) { mt = better(mt, MaxIdx{rv[0][w], ri[0][w]}); ms = better(ms, MaxIdx{rv[1][w], ri[1][w]}); }
    const float m_all = fmaxf(mt.v, ms.v);
    float sa = 0.f, sb = 0.f;
    for (int i = tid; i < P.n_vocab; i += kSampleWarps * 32) {
        const float x = value(i);
        if (x > -INFINITY) { sa += expf(x - m_all); if (i >= P.beg) sb += expf(x - ms.v); }
    }
    sa = warp_sum(sa); sb = warp_sum(sb);
    if (lane == 0) { rs[0][warp] = sa; rs[1][warp] = sb; }
    __syncthreads();
    sa = 0.f; sb = 0.f;
    for (int w = 0; w < kSampleWarps; w++) { sa += rs[0][w]; sb += rs[1][w]; }
    const float max_text = mt.v, max_ts = ms.v;
    const float lse = logf(sa) + m_all;
    const float ts_lp = sb > 0.f ? logf(sb) + (max_ts - lse) : -INFINITY;
    const bool text_off = ts_lp > max_text - lse;                  // the timestamps outweigh every text token: text is suppressed
    const float p_ts_max = max_ts > -INFINITY ? expf(max_ts - lse) : 0.f;
    const float p_ts_sum = sb * p_ts_max;
    const int tid_best = (max_ts > -INFINITY && p_ts_max > 0.f) ? ms.i : 0;
    // k selection passes: the best (value, id) strictly after the previous pick in (value descending, id ascending) order
    float prev_v = INFINITY; int prev_i = -1;
    for (int k = 0; k < K; k++) {
        MaxIdx best{-INFINITY, 0x7fffffff};
        bool have = false;
        for (int i = tid; i < P.n_vocab; i += kSampleWarps * 32) {
            float x = value(i);
            if (text_off && i < P.beg) x = -INFINITY;
            const bool after = x < prev_v || (x == prev_v && i > prev_i);
            if (after && (!have || x > best.v || (x == best.v && i < best.i))) { best = MaxIdx{x, i}; have = true; }
        }
        // reduce (a thread without a candidate carries id 0x7fffffff, which loses every tie)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MaxIdx a{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)}; best = better(best, a);
        }
        __syncthreads();                                            // rv / ri of the previous pass have been read by everybody
        if (lane == 0) { rv[0][warp] = best.v; ri[0][warp] = best.i; }
        __syncthreads();
        if (tid == 0) {
            MaxIdx t{rv[0][0], ri[0][0]};
            for (int w = 1; w < kSampleWarps; w++) t = better(t, MaxIdx{rv[0][w], ri[0][w]});
            pick = t;
            TokData tk;
            tk.id = t.i == 0x7fffffff ? 0 : t.i;
            tk.plog = t.i == 0x7fffffff ? -INFINITY : t.v - lse;
            tk.p = tk.plog > -INFINITY ? expf(tk.plog) : 0.f;
            tk.tid = tid_best; tk.pt = p_ts_max / (p_ts_sum + 1e-10f); tk.ptsum = p_ts_sum;
            if (tk.id >= P.beg) { tk.tid = tk.id; tk.pt = tk.p; }
            cand[(size_t)b * 8 + k] = tk;
        }
        __syncthreads();
        prev_v = pick.v; prev_i = pick.i;
    }
    if (tid == 0) { ctl->done = 1; atomicAdd(P.n_done, 1); }
}

bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("SS_BATCH_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
template <typename... KArgs, typename... Args>
void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args &&...args) {
    launch_pdl(pdl_enabled(), kernel, grid, block, 0, st, std::forward<Args>(args)...);
}

template <int WR, int WK, int EPI>
void launch_gemm(const BatchParams &P, int n_tiles, const __half *W, const float *bias, const __half *X, int N, int K, int il, cudaStream_t st) {
    const dim3 grid(ceil_div(N, 16 * WR)), block(kGemmThreads);
    if (n_tiles == 1) launch(bd_gemm_kernel<WR, WK, 1, EPI>, grid, block, st, P, W, bias, X, N, K, il);
    else if (n_tiles == 2) launch(bd_gemm_kernel<WR, WK, 2, EPI>, grid, block, st, P, W, bias, X, N, K, il);
    else launch(bd_gemm_kernel<WR, WK, 4, EPI>, grid, block, st, P, W, bias, X, N, K, il);
}

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace

// ================================================================================================
size_t decode_batch_scratch_bytes(int d, int n_vocab, int H, int xsplit_max) {
    const size_t R = kMaxBatch;
    return align256(R * d * 4) + 3 * align256(R * d * 2) + align256(R * 4 * d * 2) + align256(R * (size_t)n_vocab * 4) +
           align256(R * H * (size_t)xsplit_max * 66 * 4) + 256;
}

void decode_batch_bind(BatchParams &P, void *scratch) {
    const size_t R = kMaxBatch, d = P.d;
    uint8_t *p = static_cast<uint8_t *>(scratch);
    P.n_done = reinterpret_cast<int *>(p); p += 256;
    P.x = reinterpret_cast<float *>(p); p += align256(R * d * 4);
    P.xn = reinterpret_cast<__half *>(p); p += align256(R * d * 2);
    P.q = reinterpret_cast<__half *>(p); p += align256(R * d * 2);
    P.att = reinterpret_cast<__half *>(p); p += align256(R * d * 2);
    P.hid = reinterpret_cast<__half *>(p); p += align256(R * 4 * d * 2);
    P.logits = reinterpret_cast<float *>(p); p += align256(R * (size_t)P.n_vocab * 4);
    P.part = reinterpret_cast<float *>(p);
}

// key splits of the cross-attention so that at least ~2 CTAs per SM are in flight when the batch is small
int decode_batch_xsplit(int B, int H, int sms) {
    const int want = 2 * sms;
    return std::max(1, std::min(8, ceil_div(want, std::max(1, B * H))));
}

void decode_batch_step_enqueue(const BatchParams &P, const MegaParams &w, bool need_logits, int xsplit, cudaStream_t st, int *launches,
                               const BeamStep *beam) {
    const int B = P.B, d = P.d, nt = B <= 8 ? 1 : B <= 16 ? 2 : 4;
    if (B < 1 || B > kMaxBatch) SS_THROW(-1, "decode_batch: batch %d out of range", B);
    if ((d & 63) || d > kLnPer * kLnThreads || d != P.H * 64 || P.ctx > 512 || P.T > 1536) SS_THROW(-1, "decode_batch: unsupported decoder shape");
    int n = 0;
    const dim3 ln_block(kLnThreads);
    for (int il = 0; il < P.L; il++) {
        const MegaLayer &L = w.layer[il];
        launch(bd_ln_kernel, dim3(B), ln_block, st, P, L.lnw[0], L.lnb[0], il == 0 ? 1 : 0); n++;
        launch_gemm<2, 4, EPI_QKV>(P, nt, L.w[0], L.b[0], P.xn, 3 * d, d, il, st); n++;
        launch(bd_self_attn_kernel, dim3(P.H, B), dim3(kSelfWarps * 32), st, P, il); n++;
        launch_gemm<1, 8, EPI_RES>(P, nt, L.w[1], L.b[1], P.att, d, d, il, st); n++;
        launch(bd_ln_kernel, dim3(B), ln_block, st, P, L.lnw[1], L.lnb[1], 0); n++;
        launch_gemm<1, 8, EPI_Q>(P, nt, L.w[2], L.b[2], P.xn, d, d, il, st); n++;
        launch(bd_cross_attn_kernel, dim3(P.H, B, xsplit), dim3(kCrossWarps * 32), st, P, il, xsplit); n++;
        if (xsplit > 1) { launch(bd_cross_fold_kernel, dim3(P.H, B), dim3(64), st, P, xsplit); n++; }
        launch_gemm<1, 8, EPI_RES>(P, nt, L.w[3], L.b[3], P.att, d, d, il, st); n++;
        launch(bd_ln_kernel, dim3(B), ln_block, st, P, L.lnw[2], L.lnb[2], 0); n++;
        launch_gemm<2, 4, EPI_GELU>(P, nt, L.w[4], L.b[4], P.xn, 4 * d, d, il, st); n++;
        launch_gemm<1, 8, EPI_RES>(P, nt, L.w[5], L.b[5], P.hid, d, 4 * d, il, st); n++;
    }
    if (need_logits) {
        launch(bd_ln_kernel, dim3(B), ln_block, st, P, P.lnf_w, P.lnf_b, 0); n++;
        launch_gemm<8, 1, EPI_LOGITS>(P, nt, P.tok_emb, (const float *)nullptr, P.xn, P.n_vocab, d, 0, st); n++;
    }
    if (beam) {      // beam search: k candidates per live beam instead of one greedy token
        if (beam->k < 1 || beam->k > 8 || !beam->cand) SS_THROW(-1, "decode_batch: bad beam step");
        launch(bd_topk_kernel, dim3(B), dim3(kSampleWarps * 32), st, P, beam->temperature, beam->k, beam->cand); n++;
    } else {
        launch(bd_sample_kernel, dim3(B), dim3(kSampleWarps * 32), st, P); n++;
    }
    CUDA_CHECK(cudaGetLastError());
    if (launches) *launches += n;
}

// ------------------------------------------------------------------------------------------------
static bool batch_graph_enabled() {
    static const bool on = [] { const char *e = getenv("SS_BATCH_GRAPH"); return !(e && e[0] == '0'); }();
    return on;
}
BatchStepGraph::~BatchStepGraph() {
    for (Slot &s : slot) if (s.exec) cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(s.exec));
}
void decode_batch_step_graph(BatchStepGraph &G, const BatchParams &P, const MegaParams &w, bool need_logits, int xsplit, cudaStream_t st,
                             int *launches, const BeamStep *beam) {
    if (!batch_graph_enabled()) { decode_batch_step_enqueue(P, w, need_logits, xsplit, st, launches, beam); return; }
    BatchStepGraph::Slot &s = G.slot[need_logits ? 1 : 0];
    const bool same = s.valid && memcmp(&s.P, &P, sizeof(BatchParams)) == 0 && s.xsplit == xsplit && s.has_beam == (beam != nullptr) &&
                      (!beam || (s.beam.temperature == beam->temperature && s.beam.k == beam->k && s.beam.cand == beam->cand));
    if (!same) {
        if (s.exec) { cudaGraphExecDestroy(static_cast<cudaGraphExec_t>(s.exec)); s.exec = nullptr; }
        s.valid = false;
        cudaGraph_t graph = nullptr;
        int n = 0;
        CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try { decode_batch_step_enqueue(P, w, need_logits, xsplit, st, &n, beam); }
        catch (...) { cudaStreamEndCapture(st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
        CUDA_CHECK(cudaStreamEndCapture(st, &graph));
        cudaGraphExec_t exec = nullptr;
        const cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) SS_THROW(-4, "cudaGraphInstantiate of the batched decoder step failed: %s", cudaGetErrorString(e));
        s.exec = exec; s.n_launch = n; s.P = P; s.xsplit = xsplit; s.has_beam = beam != nullptr; if (beam) s.beam = *beam;
        s.valid = true;
    }
    CUDA_CHECK(cudaGraphLaunch(static_cast<cudaGraphExec_t>(s.exec), st));
    if (launches) *launches += s.n_launch;
}

}  // namespace ss
