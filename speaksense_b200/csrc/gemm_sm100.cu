// gemm_sm100.cu - TMA-fed tcgen05 GEMM for sm_100a: D = A . B^T with f16 operands, f32 accumulation in
// TMEM and a fused epilogue (bias / scale / GELU / positional add / residual / f16 or f32 store, row- or
// head-major).  This is the encoder's workhorse (BASELINE.json north_star stage 2; SURVEY.md §2.4 rows
// mul_mm / im2col / add / gelu / cpy): conv stem as implicit GEMM over an overlapping-row TMA view,
// QKV / out / MLP projections, QK^T, PV and the cross-KV projection all go through this kernel.
//
// Persistent CTAs of 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> global).  128 x BN output tiles (BN = 128 or 256), BLOCK_K = 64
// (one 128-byte swizzle atom), 4-stage mbarrier ring, two accumulators in TMEM so that a tile's epilogue
// overlaps the next tile's main loop.
#include <cuda.h>

#include "kernels.h"
#include "sm100_ptx.cuh"

namespace ss {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
enum : int { EPI_F16_BIAS = 0, EPI_F16_BIAS_GELU, EPI_F32_RESID, EPI_F32_GELU_POS, EPI_F16_HEADMAJOR, EPI_F32_PLAIN };

using namespace ptx;

struct GemmDev {
    int M, N, K;
    int nb0, nbatch;            // inner batch count (z = b1 * nb0 + b0), total batches
    int a_bcast;                // A is shared by every batch (cross-KV: one activation, 32 layers of weights)
    long bias_stride0;          // bias elements between inner batches
    const float *bias; float alpha; int alpha_cols; int gelu;
    const float *pos; int pos_rows; int residual; int out_f16;
    void *out; long out_ld, out_stride0, out_stride1; int head_major; long head_rows; int out_row_offset;
};

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float gelu_ggml(float x) {
    const float xh = r16(x);
    return r16(0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh * (1.0f + 0.044715f * xh * xh))));
}

// Epilogue of one accumulator tile for one warp: TMEM lanes [32 q, 32 q + 32) (= output rows m_base + 0..31), columns
// [chalf * BN / 2, (chalf + 1) * BN / 2) of the BN-wide tile at `tmem_acc`.  tcgen05.ld hands every THREAD one row x 32 columns;
// storing from there puts 32 different rows into one warp store (32 partly written sectors per instruction - measured: the stores
// were 1/3 of a GEMM's time).  The 32 x 32 block is therefore transposed through `stage` (2.5 KB of shared memory per warp) and
// written as whole row segments: one warp instruction = 4 rows x 128 contiguous bytes (f32) or 8 rows x 64 bytes (f16).
// Column-only math (bias, scale, GELU) happens in registers before the transpose, row-dependent math (positional add, residual
// read-modify-write, head-major swizzle) on the way out.
constexpr int kStageFloatsPerWarp = 640;      // 32 rows x 80 B (f16 blocks) >= 16 rows x 144 B (f32 half blocks)
template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmDev &p, uint32_t tmem_acc, int q, int chalf, int m_base, int n0, long zoff,
                                              const float *bias, float *stage) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
        uint32_t rr[32];
        tmem_ld_32x32b_x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c, rr);
        const int nb = n0 + c;
        if (m_base >= p.M || nb >= p.N) continue;      // (warp-uniform)
        const int nvalid = min(32, p.N - nb);
        float v[32];
        if (bias && nvalid == 32) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + nb + j));
                v[j] = __uint_as_float(rr[j]) + b4.x; v[j + 1] = __uint_as_float(rr[j + 1]) + b4.y;
                v[j + 2] = __uint_as_float(rr[j + 2]) + b4.z; v[j + 3] = __uint_as_float(rr[j + 3]) + b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = __uint_as_float(rr[j]) + ((bias && j < nvalid) ? __ldg(bias + nb + j) : 0.f);
        }
        if (EPI == EPI_F16_HEADMAJOR || EPI == EPI_F32_PLAIN) {
#pragma unroll
            for (int j = 0; j < 32; j++) if (nb + j < p.alpha_cols) v[j] *= p.alpha;
        }
        if (EPI == EPI_F16_BIAS_GELU || EPI == EPI_F32_GELU_POS) {
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = gelu_ggml(v[j]);
        }
#ifdef SS_GEMM_EXP_NOSTORE      // (timing experiment: everything but the global stores of the epilogue)
        if (p.M > 0) { float z = 0.f; for (int j = 0; j < 32; j++) z += v[j]; if (z == 1.2345e33f) reinterpret_cast<float *>(p.out)[0] = z; continue; }
#endif
        // ---- transpose through shared memory, then whole row segments per warp instruction
        constexpr bool kF16 = EPI == EPI_F16_BIAS || EPI == EPI_F16_BIAS_GELU || EPI == EPI_F16_HEADMAJOR;
        const int rows = min(32, p.M - m_base);
        if (kF16) {
            // every thread parks its row (32 halfs = 64 B) as 4 x 16 B at row stride 80 B (conflict-free); then lane = (row l / 4 + 8 k,
            // 16-byte chunk l % 4): one instruction writes 8 rows x 64 contiguous bytes
            uint4 *sw = reinterpret_cast<uint4 *>(stage);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const __half2 h0 = __floats2half2_rn(v[8 * k], v[8 * k + 1]), h1 = __floats2half2_rn(v[8 * k + 2], v[8 * k + 3]);
                const __half2 h2 = __floats2half2_rn(v[8 * k + 4], v[8 * k + 5]), h3 = __floats2half2_rn(v[8 * k + 6], v[8 * k + 7]);
                uint4 u;
                u.x = *reinterpret_cast<const uint32_t *>(&h0); u.y = *reinterpret_cast<const uint32_t *>(&h1);
                u.z = *reinterpret_cast<const uint32_t *>(&h2); u.w = *reinterpret_cast<const uint32_t *>(&h3);
                sw[lane * 5 + k] = u;
            }
            __syncwarp();
            __half *ob = reinterpret_cast<__half *>(p.out);
            const int ck = lane & 3;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int r = (lane >> 2) + 8 * k;
                if (r >= rows) continue;
                const int m = m_base + r;
                const uint4 u = sw[r * 5 + ck];
                long at;
                if (EPI == EPI_F16_HEADMAJOR) {      // cross-KV cache row: chunk c of row m stored at c ^ (m & 7) (see kernels.h)
                    const int c8 = ((nb & 63) >> 3) + ck;
                    at = ((long)(nb >> 6) * p.head_rows + m) * 64 + zoff + ((c8 ^ (m & 7)) << 3);
                } else at = (long)(m + p.out_row_offset) * p.out_ld + nb + 8 * ck + zoff;
                if (8 * ck + 7 < nvalid && (at & 7) == 0) *reinterpret_cast<uint4 *>(ob + at) = u;
                else {
                    const __half *hv = reinterpret_cast<const __half *>(&u);
                    for (int j = 0; j < 8; j++) if (8 * ck + j < nvalid) ob[at + j] = hv[j];      // (row-major only: head-major blocks are always full)
                }
            }
        } else {
            // f32: 16 rows at a time (row stride 144 B, conflict-free); lane = (row l / 8 + 4 k, float4 chunk l % 8): one instruction
            // reads / writes 4 rows x 128 contiguous bytes; all residual (positional) loads of a half are in flight before the stores
            float4 *sw = reinterpret_cast<float4 *>(stage);
            float *ob = reinterpret_cast<float *>(p.out);
            const int ck = lane & 7;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if ((lane >> 4) == h) {
#pragma unroll
                    for (int k = 0; k < 8; k++) sw[(lane & 15) * 9 + k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                }
                __syncwarp();
                float4 val[4], old[4];
                long at[4];
                bool on[4], vec[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int r = 16 * h + (lane >> 3) + 4 * k, m = m_base + r;
                    on[k] = r < rows && 4 * ck < nvalid;
                    at[k] = (long)(m + p.out_row_offset) * p.out_ld + nb + 4 * ck + zoff;
                    vec[k] = on[k] && 4 * ck + 3 < nvalid && (at[k] & 3) == 0;
                    val[k] = sw[((lane >> 3) + 4 * k) * 9 + ck];
                    old[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (EPI == EPI_F32_RESID && vec[k]) old[k] = *reinterpret_cast<const float4 *>(ob + at[k]);
                    if (EPI == EPI_F32_GELU_POS && vec[k]) old[k] = __ldg(reinterpret_cast<const float4 *>(p.pos + (size_t)(m % p.pos_rows) * p.N + nb + 4 * ck));
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (vec[k]) *reinterpret_cast<float4 *>(ob + at[k]) = make_float4(val[k].x + old[k].x, val[k].y + old[k].y, val[k].z + old[k].z, val[k].w + old[k].w);
                    else if (on[k]) {      // ragged right edge / unaligned rows: element by element
                        const int m = m_base + 16 * h + (lane >> 3) + 4 * k;
                        const float e[4] = {val[k].x, val[k].y, val[k].z, val[k].w};
                        for (int j = 0; j < 4; j++) {
                            if (4 * ck + j >= nvalid) break;
                            float x = e[j];
                            if (EPI == EPI_F32_GELU_POS) x += __ldg(p.pos + (size_t)(m % p.pos_rows) * p.N + nb + 4 * ck + j);
                            if (EPI == EPI_F32_RESID) x += ob[at[k] + j];
                            ob[at[k] + j] = x;
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();      // the block is out before the next one overwrites the staging rows
    }
}

// Persistent, warp-specialised: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, + gridDim.x, ...
// The smem ring (TMA <-> MMA) keeps running across tile boundaries and the accumulator is double-buffered in
// TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i + 1.
template <int BN, int kStages, int EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                   const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr uint32_t kABytes = BM * BK * 2;            // 16 KB
    constexpr uint32_t kBBytes = BN * BK * 2;
    constexpr uint32_t kTmemCols = 2 * BN;               // 256 or 512: two accumulators
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kStages * kABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;       // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *stage_all = reinterpret_cast<float *>(tmem_slot + 4);      // 8 epilogue warps x 2.5 KB (transposing stores), 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = (p.K + BK - 1) / BK;
    const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
    const int tiles_mn = tiles_n * tiles_m;
    const int total = tiles_mn * p.nbatch;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // programmatic dependent launch: the prologue above (barriers, tensor memory, descriptor prefetch) ran while the predecessor was
    // still finishing on other SMs; from here on the kernel reads and overwrites what the predecessor produced / still reads
    pdl_trigger();
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;     // running k-block counter across tiles (ring position)
            for (int t = blockIdx.x; t < total; t += gridDim.x) {
                const int z = t / tiles_mn, r = t - z * tiles_mn;
                const int m0 = (r / tiles_n) * BM, n0 = (r % tiles_n) * BN;
                const int b0 = z % p.nb0, b1 = z / p.nb0;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
#ifdef SS_GEMM_EXP_NOLOAD       // (timing experiment, garbage results: no operand traffic at all)
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
#else
                    mbar_expect_tx(&full[s], kABytes + kBBytes);
                    tma_load_4d(&tmA, &full[s], sA + s * kABytes, kb * BK, m0, p.a_bcast ? 0 : b0, p.a_bcast ? 0 : b1);
                    tma_load_4d(&tmB, &full[s], sB + s * kBBytes, kb * BK, n0, b0, b1);
#endif
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1<<4), a/b F16 (0),
            // both K-major, N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, i = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, i++) {
                const uint32_t acc = i & 1;
                mbar_wait(&tmem_empty[acc], ((i >> 1) & 1) ^ 1);      // epilogue drained this accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % kStages;
#ifndef SS_GEMM_EXP_NOWAIT      // (timing experiment, garbage results: the MMAs do not wait for their operands)
                    mbar_wait(&full[s], (it / kStages) & 1);
#endif
                    tcgen05_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * kABytes));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)     // +32 B per UMMA_K inside the 128-byte swizzle atom
                        tcgen05_mma_f16(tmem_base + acc * BN, adesc + (uint64_t)((k * 32) >> 4), bdesc + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
                    tcgen05_commit(&empty[s]);
                }
                tcgen05_commit(&tmem_full[acc]);
            }
        }
    } else {
        // ---------------- epilogue: warps w and w+4 own TMEM lanes [32q, 32q+32), q = w & 3, and split the columns ----------------
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;                  // 0: columns [0, BN/2), 1: [BN/2, BN)
        uint32_t i = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, i++) {
            const int z = t / tiles_mn, r = t - z * tiles_mn;
            const int m0 = (r / tiles_n) * BM, n0 = (r % tiles_n) * BN;
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const uint32_t acc = i & 1;
            mbar_wait(&tmem_full[acc], (i >> 1) & 1);
            tcgen05_fence_after();
            const long zoff = (long)b0 * p.out_stride0 + (long)b1 * p.out_stride1;
            const float *bias = p.bias ? p.bias + (long)b0 * p.bias_stride0 : nullptr;
            epilogue_tile<BN, EPI>(p, tmem_base + acc * BN, q, chalf, m0 + q * 32, n0, zoff, bias, stage_all + (warp - 2) * kStageFloatsPerWarp);
            tcgen05_fence_before();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[acc])) : "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (round 2): tcgen05.mma.cta_group::2 on 256 x BN tiles.  The 1-CTA kernel above pulls 48 KB per 64-wide k-step
// through one SM's L2 port for a 128 x 256 tile and is bound by that operand delivery (DESIGN.md §5 "Encoder: what bounds it").
// Here the two CTAs of a (2,1,1) cluster share one MMA of M = 256: CTA r loads A rows [128 r, 128 r + 128) and B rows (output
// columns) [BN / 2 * r, BN / 2 * (r + 1)) of the tile - 32 KB per CTA and k-step for the same work per CTA (BN = 256) - and ends
// up with accumulator rows [128 r, +128) x all BN columns in its own TMEM.  Only the leader (cluster rank 0) issues the MMAs:
//   full[s]  (leader's)   : expects the bytes of BOTH CTAs' loads; the peer's TMA completes on it through its cluster address
//   empty[s] (both CTAs') : tcgen05.commit multicast - the stage is free in both CTAs once the MMAs that read it are done
//   tmem_full[a] (both)   : tcgen05.commit multicast - the accumulator is complete, both epilogues start
//   tmem_empty[a] (leader's): 16 arrivals = the 8 epilogue warps of each CTA (the peer's arrive remotely)
// ------------------------------------------------------------------------------------------------
template <int BN, int kStages, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int BH = BN / 2;                           // B rows (output columns) this CTA loads
    constexpr uint32_t kABytes = BM * BK * 2;            // 16 KB
    constexpr uint32_t kBBytes = BH * BK * 2;
    constexpr uint32_t kTmemCols = 2 * BN;               // two accumulators of BN columns (128 lanes each, per CTA)
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + kStages * kABytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(sB + kStages * kBBytes);
    uint64_t *empty = full + kStages;
    uint64_t *tmem_full = empty + kStages;      // [2]
    uint64_t *tmem_empty = tmem_full + 2;       // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float *stage_all = reinterpret_cast<float *>(tmem_slot + 4);      // 8 epilogue warps x 2.5 KB (transposing stores), 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int nkb = (p.K + BK - 1) / BK;
    const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + 2 * BM - 1) / (2 * BM);
    const int tiles_mn = tiles_n * tiles_m;
    const int total = tiles_mn * p.nbatch;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {      // the same warp of both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = pair; t < total; t += n_pairs) {
                const int z = t / tiles_mn, r = t - z * tiles_mn;
                const int m0 = (r / tiles_n) * 2 * BM + (int)rank * BM, n0 = (r % tiles_n) * BN + (int)rank * BH;
                const int b0 = z % p.nb0, b1 = z / p.nb0;
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % kStages;
                    mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
                    if (leader) mbar_expect_tx(&full[s], 2 * (kABytes + kBBytes));
                    const uint32_t lf = mapa_u32(smem_u32(&full[s]), 0);
                    tma_load_4d_2sm(&tmA, lf, sA + s * kABytes, kb * BK, m0, p.a_bcast ? 0 : b0, p.a_bcast ? 0 : b1);
                    tma_load_4d_2sm(&tmB, lf, sB + s * kBBytes, kb * BK, n0, b0, b1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);      // M = 256
            uint32_t it = 0, i = 0;
            for (int t = pair; t < total; t += n_pairs, i++) {
                const uint32_t acc = i & 1;
                mbar_wait(&tmem_empty[acc], ((i >> 1) & 1) ^ 1);      // both epilogues drained this accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < nkb; kb++, it++) {
                    const int s = it % kStages;
                    mbar_wait(&full[s], (it / kStages) & 1);
                    tcgen05_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * kABytes));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        tcgen05_mma_f16_2cta(tmem_base + acc * BN, adesc + (uint64_t)((k * 32) >> 4), bdesc + (uint64_t)((k * 32) >> 4), idesc, (kb | k) ? 1u : 0u);
                    tcgen05_commit_2cta(&empty[s], 3);
                }
                tcgen05_commit_2cta(&tmem_full[acc], 3);
            }
        }
    } else {
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const uint32_t leader_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
        uint32_t i = 0;
        for (int t = pair; t < total; t += n_pairs, i++) {
            const int z = t / tiles_mn, r = t - z * tiles_mn;
            const int m0 = (r / tiles_n) * 2 * BM + (int)rank * BM, n0 = (r % tiles_n) * BN;
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const uint32_t acc = i & 1;
            mbar_wait(&tmem_full[acc], (i >> 1) & 1);
            tcgen05_fence_after();
            const long zoff = (long)b0 * p.out_stride0 + (long)b1 * p.out_stride1;
            const float *bias = p.bias ? p.bias + (long)b0 * p.bias_stride0 : nullptr;
            epilogue_tile<BN, EPI>(p, tmem_base + acc * BN, q, chalf, m0 + q * 32, n0, zoff, bias, stage_all + (warp - 2) * kStageFloatsPerWarp);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_empty0 + acc * 8);
        }
    }
    __syncthreads();
    cluster_sync_all();          // nobody frees tensor memory (or exits) while the pair's MMAs / remote arrivals may still target it
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

// 128 x 128 tiles: 32 KB per stage - six stages keep 192 KB of operands in flight (a stage's TMA round trip is ~1.4 us under load; with
// four stages the 128 x 128 main loop ran at 0.35 us per k-step against 0.14 us of tcgen05.mma work)
#ifndef SS_GEMM_STAGES128
#define SS_GEMM_STAGES128 6
#endif
constexpr int kStages128 = SS_GEMM_STAGES128;
template <int BN, int kStages>
constexpr size_t smem_bytes() { return 1024 + (size_t)kStages * (BM * BK * 2 + BN * BK * 2) + (2 * kStages + 4) * 8 + 16 + 8 * 640 * 4; }
// CTA-pair kernel: per CTA 128 A rows + BN / 2 B rows per stage
template <int BN> constexpr int stages_2cta() { return BN == 256 ? 6 : 8; }
template <int BN>
constexpr size_t smem_bytes_2cta() { return 1024 + (size_t)stages_2cta<BN>() * (BM * BK * 2 + (BN / 2) * BK * 2) + (2 * stages_2cta<BN>() + 4) * 8 + 16 + 8 * 640 * 4; }

void make_map(CUtensorMap *map, const GemmOperand &op, long inner, long rows, int box_inner, int box_rows) {
    cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)op.batch0, (cuuint64_t)op.batch1};
    auto fix = [&](long s, long fallback) { long v = s > 0 ? s : fallback; return (cuuint64_t)v * 2; };
    const long natural = op.ld * rows;
    cuuint64_t strides[3] = {(cuuint64_t)op.ld * 2, fix(op.stride0, natural), fix(op.stride1, natural * op.batch0)};
    cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < 3; i++)
        if (strides[i] % 16) SS_THROW(-9, "GEMM operand stride %llu B is not a multiple of 16", (unsigned long long)strides[i]);
    if (reinterpret_cast<uintptr_t>(op.ptr) % 16) SS_THROW(-9, "GEMM operand pointer is not 16-byte aligned");
    CUresult rc = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half *>(op.ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) SS_THROW(-4, "cuTensorMapEncodeTiled failed with %d (inner %ld rows %ld ld %ld)", (int)rc, inner, rows, op.ld);
}

// per device (one process may own several GPUs): SM count, and the > 48 KB dynamic shared memory opt-in of every instantiation
int g_sms_tab[64] = {0};
PerDeviceOnce g_gemm_ready;
thread_local int t_sms = 0;      // SM count of the device gemm_enqueue is launching on

template <int BN, int kStages, int EPI>
void configure_one() {
    CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, kStages, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<BN, kStages>()));
}
template <int BN, int EPI>
void configure_one_2cta() {
    CUDA_CHECK(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<BN, stages_2cta<BN>(), EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_2cta<BN>()));
}
template <int EPI>
void configure_epi() { configure_one<128, kStages128, EPI>(); configure_one<256, 4, EPI>(); configure_one_2cta<128, EPI>(); configure_one_2cta<256, EPI>(); }

template <int BN, int kStages, int EPI>
void launch(const GemmOperand &A, const GemmOperand &B, const GemmDev &p, cudaStream_t st) {
    const int g_sms = t_sms;
    CUtensorMap ta, tb;
    make_map(&ta, A, p.K, A.rows, BK, BM);
    make_map(&tb, B, p.K, B.rows, BK, BN);
    const int total = ceil_div(p.N, BN) * ceil_div(p.M, BM) * p.nbatch;
    launch_pdl(encoder_pdl_enabled(), gemm_tcgen05_kernel<BN, kStages, EPI>, dim3(std::min(total, g_sms)), dim3(kThreads), smem_bytes<BN, kStages>(), st, ta, tb, p);
    CUDA_CHECK(cudaGetLastError());
}
template <int BN, int EPI>
void launch_2cta(const GemmOperand &A, const GemmOperand &B, const GemmDev &p, cudaStream_t st) {
    const int g_sms = t_sms;
    CUtensorMap ta, tb;
    make_map(&ta, A, p.K, A.rows, BK, BM);
    make_map(&tb, B, p.K, B.rows, BK, BN / 2);
    const int total = ceil_div(p.N, BN) * ceil_div(p.M, 2 * BM) * p.nbatch;
    const int pairs = std::min(total, g_sms / 2);
    gemm_tcgen05_2cta_kernel<BN, stages_2cta<BN>(), EPI><<<2 * pairs, kThreads, smem_bytes_2cta<BN>(), st>>>(ta, tb, p);
    CUDA_CHECK(cudaGetLastError());
}
// mode: 0 = 1-CTA 128 x 128, 1 = 1-CTA 128 x 256, 2 = CTA pair 256 x 128, 3 = CTA pair 256 x 256
template <int EPI>
void launch_bn(int mode, const GemmOperand &A, const GemmOperand &B, const GemmDev &p, cudaStream_t st) {
    if (mode == 3) launch_2cta<256, EPI>(A, B, p, st);
    else if (mode == 2) launch_2cta<128, EPI>(A, B, p, st);
    else if (mode == 1) launch<256, 4, EPI>(A, B, p, st);
    else launch<128, kStages128, EPI>(A, B, p, st);
}

}  // namespace

void make_tensor_map_4d(void *map, const GemmOperand &op, long inner, long rows, int box_inner, int box_rows) {
    if (!g_encode) gemm_init();
    make_map(reinterpret_cast<CUtensorMap *>(map), op, inner, rows, box_inner, box_rows);
}

void gemm_init() {
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) SS_THROW(-4, "cuTensorMapEncodeTiled is not available in this driver");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (!g_gemm_ready.need(dev)) return;
    int sms = 0;
    CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configure_epi<EPI_F16_HEADMAJOR>(); configure_epi<EPI_F16_BIAS_GELU>(); configure_epi<EPI_F16_BIAS>();
    configure_epi<EPI_F32_RESID>(); configure_epi<EPI_F32_GELU_POS>(); configure_epi<EPI_F32_PLAIN>();
    if (dev >= 0 && dev < 64) g_sms_tab[dev] = sms;
    g_gemm_ready.done(dev);
}

void gemm_enqueue(const GemmOperand &A, const GemmOperand &B, int M, int N, int K, bool b_mn_major, const GemmEpilogue &ep,
                  cudaStream_t st, int *launches) {
    GemmDev p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = ep.bias; p.alpha = ep.alpha; p.alpha_cols = ep.alpha_cols; p.gelu = ep.gelu; p.pos = ep.pos; p.pos_rows = ep.pos_rows > 0 ? ep.pos_rows : 1;
    p.residual = ep.residual; p.out_f16 = ep.out_type == GEMM_OUT_F16; p.out = ep.out; p.out_ld = ep.out_ld;
    p.out_stride0 = ep.out_stride0; p.out_stride1 = ep.out_stride1; p.head_major = ep.head_major; p.head_rows = ep.head_rows;
    p.out_row_offset = ep.out_row_offset;
    if (p.residual && p.out_f16) SS_THROW(-9, "residual epilogue needs an f32 output");
    p.a_bcast = ep.a_broadcast; p.bias_stride0 = ep.bias_stride0;
    if (!p.a_bcast && (A.batch0 != B.batch0 || A.batch1 != B.batch1)) SS_THROW(-9, "GEMM batch mismatch");
    p.nb0 = (int)B.batch0; p.nbatch = (int)(B.batch0 * B.batch1);
    if (b_mn_major) SS_THROW(-1, "MN-major B is only supported inside the fused attention kernel");
    const int dev = current_device();
    if (dev < 0 || dev >= 64) SS_THROW(-3, "GEMM on CUDA device %d: only devices 0..63 are supported", dev);
    if (g_gemm_ready.need(dev)) gemm_init();
    const int g_sms = t_sms = g_sms_tab[dev];
    // tile width: rounds of the persistent grid x relative tile cost (a 128x256 tile costs ~1.6x a 128x128 one)
    const long t128 = (long)ceil_div(N, 128) * ceil_div(M, BM) * p.nbatch, t256 = (long)ceil_div(N, 256) * ceil_div(M, BM) * p.nbatch;
    const double c128 = (double)ceil_div<long>(t128, g_sms), c256 = 1.6 * (double)ceil_div<long>(t256, g_sms);
    int wide = (N >= 256 && c256 < c128) ? 1 : 0;
    // CTA pairs (SS_GEMM_2CTA, on by default): rounds of the persistent grid of sms / 2 pairs x relative tile cost per CTA
    // (a 256 x 128 pair tile costs each CTA what a 128 x 128 tile does, with 24 instead of 32 KB of operands per k-step;
    //  256 x 256: the work of a 128 x 256 tile with 32 instead of 48 KB)
    const char *pair_env = getenv("SS_GEMM_2CTA");      // "0": never, "2": whenever the shape allows (tests), else: by the rule below
    const bool pair_on = !(pair_env && pair_env[0] == '0'), pair_force = pair_env && pair_env[0] == '2';
    // measured (large-v3 encoder): one clip (M = 1500, <= 240 tiles): 5.32 ms per window with 1-CTA tiles, 5.45 ms with pairs; the
    // batched pass of 32 clips (M = 48000): 103 ms vs 100 ms - pairs pay once a GEMM has several waves of tiles
    if (pair_on && g_sms >= 2 && M > BM && (pair_force || t256 >= 4l * g_sms)) {
        const int np = g_sms / 2;
        const long p128 = (long)ceil_div(N, 128) * ceil_div(M, 2 * BM) * p.nbatch, p256 = (long)ceil_div(N, 256) * ceil_div(M, 2 * BM) * p.nbatch;
        const double d128 = (double)ceil_div<long>(p128, np), d256 = 1.6 * (double)ceil_div<long>(p256, np);
        wide = (N >= 256 && d256 < d128) ? 3 : 2;
    }
    if (ep.head_major) {
        if (!p.out_f16 || ep.gelu || ep.pos || ep.residual) SS_THROW(-9, "unsupported epilogue combination");
        launch_bn<EPI_F16_HEADMAJOR>(wide, A, B, p, st);
    } else if (p.out_f16) {
        if (ep.pos || ep.residual || ep.alpha_cols) SS_THROW(-9, "unsupported epilogue combination");
        if (ep.gelu) launch_bn<EPI_F16_BIAS_GELU>(wide, A, B, p, st); else launch_bn<EPI_F16_BIAS>(wide, A, B, p, st);
    } else if (ep.residual) {
        if (ep.gelu || ep.pos || ep.alpha_cols) SS_THROW(-9, "unsupported epilogue combination");
        launch_bn<EPI_F32_RESID>(wide, A, B, p, st);
    } else if (ep.gelu && ep.pos) {
        launch_bn<EPI_F32_GELU_POS>(wide, A, B, p, st);
    } else {
        if (ep.gelu || ep.pos) SS_THROW(-9, "unsupported epilogue combination");
        launch_bn<EPI_F32_PLAIN>(wide, A, B, p, st);
    }
    (*launches)++;
}

}  // namespace ss
