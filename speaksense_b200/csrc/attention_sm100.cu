// attention_sm100.cu - fused non-causal attention of the audio encoder on tcgen05 tensor cores
// (BASELINE.json north_star stage 2: "self-attention ... as tcgen05 tensor-core GEMMs fed by TMA with
// warp-specialised softmax"; SURVEY.md §2.4 rows mul_mm / soft_max / cpy: "fused into attention kernels
// (online softmax); never materialise KQ").
//
// One CTA = one (clip, head, 128-query tile); keys/values stream in blocks of 128 (T = 1500 -> 12 blocks).
//   warp 0     : TMA producer  - Q tile once, then K_j / V_j blocks into a 2-stage ring
//   warp 1     : TMEM allocator + single-thread MMA issuer
//                  S_j  = Q . K_j^T         (M128 x N128 x K64,  A,B K-major)   -> TMEM S[j & 1]
//                  PV_j = P_j . V_j         (M128 x N64  x K128, B = V MN-major) -> TMEM O[j & 1]
//   warps 2..  : softmax / correction, NS = 2 or 4 threads per query row (the NS warps w, w + 4, .. share a TMEM lane quarter: part
//                p takes keys [p * 128 / NS, ..) and output channels [p * 64 / NS, ..) of every block; the block maximum is
//                exchanged through shared memory + a named barrier of the row's NS warps): tcgen05.ld S_j, running
//                max and sum (exp2), P_j (f16) written into shared memory in the 128-byte-swizzled K-major layout the
//                MMA reads, and the rescaled accumulation of PV_{j-1} in registers.  (One thread per row - 4 softmax
//                warps, one per scheduler - left the kernel exp / issue bound at 7 % tensor-pipe activity.)
// S and P never leave the SM: per layer this removes 0.27 GB of HBM traffic and three launches per head
// batch compared with the unfused path of round-1 v0 (S GEMM + softmax + PV GEMM = 340 us per layer).
//
// Arithmetic: f16 Q/K/V, f32 scores and running statistics, P rounded to f16 before P.V (ggml rounds the
// normalised probabilities to f16; here the un-normalised ones - same relative rounding), f32 accumulate.
#include <cstdlib>

#include "kernels.h"
#include "sm100_ptx.cuh"

namespace ss {

namespace {

using namespace ptx;

constexpr int kBQ = 128;          // queries per CTA
constexpr int kBK = 128;          // keys per block
constexpr int kD = 64;            // head dim
constexpr int attn_threads(int ns) { return 64 + ns * 128; }      // TMA warp, MMA warp, NS x 4 softmax warps
constexpr uint32_t kQBytes = kBQ * kD * 2;        // 16 KB
constexpr uint32_t kKBytes = kBK * kD * 2;        // 16 KB
constexpr uint32_t kVBytes = kBK * kD * 2;        // 16 KB
constexpr uint32_t kPBytes = kBQ * kBK * 2;       // 32 KB (two 64-wide K halves)
constexpr uint32_t kTmemCols = 512;               // S[2] at 0 / 128, O[2] at 256 / 320
#ifndef SS_ATTN_SPLIT_DEFAULT
#define SS_ATTN_SPLIT_DEFAULT 2
#endif

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct AttnDev {
    int T;                 // keys == queries per (clip, head)
    int H;
    float scale_log2;      // softmax scale * log2(e)
    __half *out;           // [clip][T][H*64]
    long out_ld, out_clip_stride;
};

template <int NS>
__global__ void __launch_bounds__(attn_threads(NS), 1) attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                             const __grid_constant__ CUtensorMap tmK,
                                                                             const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem;
    uint8_t *sK = sQ + kQBytes;               // 2 stages
    uint8_t *sV = sK + 2 * kKBytes;           // 2 stages
    uint8_t *sP = sV + 2 * kVBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + kPBytes);
    uint64_t *q_full = bars, *kv_full = bars + 1, *kv_empty = bars + 3, *s_full = bars + 5, *p_full = bars + 7, *o_full = bars + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 10);
    float *xmax = reinterpret_cast<float *>(bars + 16);      // [2 blocks in flight][NS parts][128 rows] block maxima
    float *xsum = xmax + 2 * NS * kBQ;                       // [NS parts][128 rows] final partial row sums
    constexpr int KP = kBK / NS, CP = kD / NS;               // keys per block / output channels of one softmax thread

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kBQ, h = blockIdx.y, clip = blockIdx.z;
    const int nb = (p.T + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; s++) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); mbar_init(&s_full[s], 1); mbar_init(&o_full[s], 1); }
        mbar_init(p_full, 128 * NS);
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
    pdl_trigger();      // (programmatic dependent launch: the prologue overlaps the tail of the QKV GEMM; see gemm_sm100.cu)
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, kQBytes);
            tma_load_4d(&tmQ, q_full, sQ, 0, q0, h, clip);
            for (int j = 0; j < nb; j++) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], kKBytes + kVBytes);
                tma_load_4d(&tmK, &kv_full[s], sK + s * kKBytes, 0, j * kBK, h, clip);
                tma_load_4d(&tmV, &kv_full[s], sV + s * kVBytes, 0, j * kBK, h, clip);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptors: F32 accumulate, F16 operands; S: N=128, both K-major; PV: N=64, B MN-major
            const uint32_t idesc_s = (1u << 4) | ((uint32_t)(kBK >> 3) << 17) | ((uint32_t)(kBQ >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 16) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kBQ >> 4) << 24);
            const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
            const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
            mbar_wait(q_full, 0);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&kv_full[s], (j >> 1) & 1);
                tcgen05_fence_after();
                const uint64_t kdesc = umma_desc_sw128(smem_u32(sK + s * kKBytes));
#pragma unroll
                for (int k = 0; k < kD / 16; k++)
                    tcgen05_mma_f16(tmem_base + (uint32_t)(s * 128), qdesc + (uint64_t)((k * 32) >> 4), kdesc + (uint64_t)((k * 32) >> 4), idesc_s, k ? 1u : 0u);
                tcgen05_commit(&s_full[s]);
            };
            issue_s(0);
            for (int j = 0; j < nb; j++) {
                if (j + 1 < nb) issue_s(j + 1);
                const int s = j & 1;
                mbar_wait(p_full, j & 1);
                tcgen05_fence_after();
                const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + s * kVBytes));
#pragma unroll
                for (int k = 0; k < kBK / 16; k++) {
                    // A = P: K-major, two 64-wide halves of [128 rows][128 B]; B = V: MN-major, 16 keys = 2048 B per step
                    const uint64_t ad = pdesc + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4);
                    const uint64_t bd = vdesc + (uint64_t)((k * 2048) >> 4);
                    tcgen05_mma_f16(tmem_base + 256u + (uint32_t)(s * 64), ad, bd, idesc_pv, k ? 1u : 0u);
                }
                tcgen05_commit(&o_full[s]);      // PV_j done: O[s] readable, P and the K/V stage reusable
                tcgen05_commit(&kv_empty[s]);
            }
        }
    } else {
        // ---------------- softmax / correction: NS threads per query row ----------------
        const int q = warp & 3, part = (warp - 2) >> 2;     // TMEM lane quarter; which KP keys / CP channels of the row
        const int row = q * 32 + lane;                      // row inside the tile == TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t row_bar = 1u + (uint32_t)q;          // named barrier of the NS warps of this lane quarter
        const int k0 = part * KP;
        float m = -INFINITY, l = 0.f;
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; c++) acc[c] = 0.f;
        auto load_o = [&](int s, float (&a)[CP]) {          // += this thread's CP channels of O[s]
            if constexpr (CP == 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(lane_base + 256u + (uint32_t)(s * 64 + part * CP), r);
#pragma unroll
                for (int i = 0; i < 32; i++) a[i] += __uint_as_float(r[i]);
            } else {
                uint32_t r[16];
                tmem_ld_32x32b_x16(lane_base + 256u + (uint32_t)(s * 64 + part * CP), r);
#pragma unroll
                for (int i = 0; i < 16; i++) a[i] += __uint_as_float(r[i]);
            }
        };
        for (int j = 0; j < nb; j++) {
            const int s = j & 1;
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tcgen05_fence_after();
            // pass 1: maximum of this thread's KP keys (scores scaled into the exp2 domain), then the other parts'.
            // The scores stay in registers for pass 2 (one TMEM read per block instead of two).
            float mj = -INFINITY;
            const int kvalid = min(kBK, p.T - j * kBK);
            uint32_t sr[KP / 32][32];
#pragma unroll
            for (int u = 0; u < KP / 32; u++) tmem_ld_32x32b_x32(lane_base + (uint32_t)(s * 128 + k0 + 32 * u), sr[u]);
            const bool full = kvalid == kBK;      // every block but the last: no per-key predicates (the softmax warps are issue-bound)
            if (full) {
                float mr = -INFINITY;             // maximum of the raw scores; the (positive) scale is applied once
#pragma unroll
                for (int u = 0; u < KP / 32; u++) {
#pragma unroll
                    for (int i = 0; i < 32; i++) mr = fmaxf(mr, __uint_as_float(sr[u][i]));
                }
                mj = mr * p.scale_log2;
            } else {
#pragma unroll
                for (int u = 0; u < KP / 32; u++) {
#pragma unroll
                    for (int i = 0; i < 32; i++) if (k0 + 32 * u + i < kvalid) mj = fmaxf(mj, __uint_as_float(sr[u][i]) * p.scale_log2);
                }
            }
            xmax[(s * NS + part) * kBQ + row] = mj;
            asm volatile("bar.sync %0, %1;" ::"r"(row_bar), "n"(32 * NS) : "memory");
#pragma unroll
            for (int o = 1; o < NS; o++) mj = fmaxf(mj, xmax[(s * NS + ((part + o) % NS)) * kBQ + row]);
            const float m_new = fmaxf(m, mj);
            const float alpha = exp2f(m - m_new);          // 0 on the first block (m = -inf)
            // fold the previous block's P.V (this thread's CP channels) into the accumulator, then rescale to the new maximum
            if (j > 0) {
                mbar_wait(&o_full[s ^ 1], ((j - 1) >> 1) & 1);
                tcgen05_fence_after();
                load_o(s ^ 1, acc);
            }
#pragma unroll
            for (int c = 0; c < CP; c++) acc[c] *= alpha;
            // pass 2: probabilities -> shared memory (f16, swizzled K-major A operand), partial row sum
            float lsum = 0.f;
            auto store_p = [&](int c, const uint32_t (&pk)[16]) {
#pragma unroll
                for (int g = 0; g < 4; g++) {          // four 16-byte chunks of 8 keys
                    const int cc = (c >> 3) + g;       // chunk index 0..15 inside the 128-key row
                    uint8_t *dst = sP + (cc >> 3) * 16384 + row * 128 + (((cc & 7) ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
            };
            if (full) {
                const float nm = -m_new;
#pragma unroll
                for (int u = 0; u < KP / 32; u++) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = ex2_approx(fmaf(__uint_as_float(sr[u][i]), p.scale_log2, nm));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(sr[u][i + 1]), p.scale_log2, nm));
                        __half2 hp = __floats2half2_rn(p0, p1);
                        lsum += __low2float(hp) + __high2float(hp);      // sum what the MMA will actually see
                        pk[i >> 1] = *reinterpret_cast<uint32_t *>(&hp);
                    }
                    store_p(k0 + 32 * u, pk);
                }
            } else {
#pragma unroll
                for (int u = 0; u < KP / 32; u++) {
                    const int c = k0 + 32 * u;
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = (c + i < kvalid) ? exp2f(__uint_as_float(sr[u][i]) * p.scale_log2 - m_new) : 0.f;
                        const float p1 = (c + i + 1 < kvalid) ? exp2f(__uint_as_float(sr[u][i + 1]) * p.scale_log2 - m_new) : 0.f;
                        __half2 hp = __floats2half2_rn(p0, p1);
                        lsum += __low2float(hp) + __high2float(hp);
                        pk[i >> 1] = *reinterpret_cast<uint32_t *>(&hp);
                    }
                    store_p(c, pk);
                }
            }
            l = l * alpha + lsum;
            m = m_new;
            tcgen05_fence_before();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes of P -> tensor-core reads
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(p_full)) : "memory");
        }
        {   // last block's P.V
            const int s = (nb - 1) & 1;
            mbar_wait(&o_full[s], ((nb - 1) >> 1) & 1);
            tcgen05_fence_after();
            load_o(s, acc);
        }
        xsum[part * kBQ + row] = l;
        asm volatile("bar.sync %0, %1;" ::"r"(row_bar), "n"(32 * NS) : "memory");
#pragma unroll
        for (int o = 1; o < NS; o++) l += xsum[((part + o) % NS) * kBQ + row];
        const int t = q0 + row;
        if (t < p.T) {
            const float inv = 1.0f / l;
            __half *o = p.out + (size_t)clip * p.out_clip_stride + (size_t)t * p.out_ld + h * kD + part * CP;
#pragma unroll
            for (int c = 0; c < CP; c += 8) {
                __half2 h0 = __floats2half2_rn(acc[c] * inv, acc[c + 1] * inv), h1 = __floats2half2_rn(acc[c + 2] * inv, acc[c + 3] * inv);
                __half2 h2 = __floats2half2_rn(acc[c + 4] * inv, acc[c + 5] * inv), h3 = __floats2half2_rn(acc[c + 6] * inv, acc[c + 7] * inv);
                uint4 u;
                u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
                u.z = *reinterpret_cast<uint32_t *>(&h2); u.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(o + c) = u;
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

constexpr size_t attn_smem(int ns) { return 1024 + kQBytes + 2 * kKBytes + 2 * kVBytes + kPBytes + 16 * 8 + (size_t)(2 * ns + ns) * kBQ * 4; }
// threads per query row: 2 (8 softmax warps) or 4 (16); SS_ATTN_SPLIT overrides.  Measured (one clip, encoder stage): 4.60 ms with 2,
// 4.59 ms with 4 - twice the softmax warps change nothing, and neither does summing the unrounded exponentials (64 conversions
// less per thread and block): a block's time is not instruction issue, it is the chain softmax -> P in shared memory ->
// mbarrier -> P.V issue -> commit -> mbarrier -> softmax (3.7 k cycles per 128 x 128 block against 1 k of MUFU work); the way out is a
// second query tile per CTA in ping-pong (TMEM: 2 x (S 128 + O 2 x 64) = 512 columns), not more threads.
int attn_split() {
    static const int ns = [] { const char *e = getenv("SS_ATTN_SPLIT"); return e && e[0] == '4' ? 4 : e && e[0] == '2' ? 2 : SS_ATTN_SPLIT_DEFAULT; }();
    return ns;
}

}  // namespace

void attention_init() {
    CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem(2)));
    CUDA_CHECK(cudaFuncSetAttribute(attention_tcgen05_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem(4)));
}

// qkv: [clips][T][3*H*64] f16 (Q | K | V); out: [clips][T][H*64] f16
void attention_enqueue(const __half *qkv, __half *out, int clips, int T, int H, float scale, cudaStream_t st, int *launches) {
    const long d = (long)H * kD, ld = 3 * d;
    CUtensorMap tq, tk, tv;
    GemmOperand Q; Q.ptr = qkv; Q.rows = T; Q.ld = ld; Q.batch0 = H; Q.stride0 = kD; Q.batch1 = clips; Q.stride1 = (long)T * ld;
    GemmOperand K = Q; K.ptr = qkv + d;
    GemmOperand V = Q; V.ptr = qkv + 2 * d;
    make_tensor_map_4d(&tq, Q, kD, T, kD, kBQ);
    make_tensor_map_4d(&tk, K, kD, T, kD, kBK);
    make_tensor_map_4d(&tv, V, kD, T, kD, kBK);      // [keys][64 dv], dv contiguous: MN-major B operand of P.V
    AttnDev p{};
    p.T = T; p.H = H; p.scale_log2 = scale * 1.4426950408889634f; p.out = out; p.out_ld = d; p.out_clip_stride = (long)T * d;
    dim3 grid(ceil_div(T, kBQ), H, clips);
    if (attn_split() == 4) launch_pdl(encoder_pdl_enabled(), attention_tcgen05_kernel<4>, grid, dim3(attn_threads(4)), attn_smem(4), st, tq, tk, tv, p);
    else launch_pdl(encoder_pdl_enabled(), attention_tcgen05_kernel<2>, grid, dim3(attn_threads(2)), attn_smem(2), st, tq, tk, tv, p);
    CUDA_CHECK(cudaGetLastError());
    (*launches)++;
}

}  // namespace ss
