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
                    mbar_expect_tx(&full[s], kABytes + kBBytes);
                    tma_load_4d(&tmA, &full[s], sA + s * kABytes, kb * BK, m0, p.a_bcast ? 0 : b0, p.a_bcast ? 0 : b1);
                    tma_load_4d(&tmB, &full[s], sB + s * kBBytes, kb * BK, n0, b0, b1);
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
                    mbar_wait(&full[s], (it / kStages) & 1);
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
            const int m = m0 + q * 32 + lane;
            mbar_wait(&tmem_full[acc], (i >> 1) & 1);
            tcgen05_fence_after();
            const long zoff = (long)b0 * p.out_stride0 + (long)b1 * p.out_stride1;
            const float *bias = p.bias ? p.bias + (long)b0 * p.bias_stride0 : nullptr;
            const bool row_ok = m < p.M;
#pragma unroll 1
            for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
                uint32_t rr[32];
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), rr);
                const int nb = n0 + c;
                if (!row_ok || nb >= p.N) continue;
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
                if (EPI == EPI_F32_GELU_POS) {
                    const float *pr = p.pos + (size_t)(m % p.pos_rows) * p.N + nb;
#pragma unroll
                    for (int j = 0; j < 32; j++) if (j < nvalid) v[j] += __ldg(pr + j);
                }
                long idx;
                if (EPI == EPI_F16_HEADMAJOR) idx = ((long)(nb >> 6) * p.head_rows + m) * 64 + zoff;      // start of row m of head nb / 64
                else idx = (long)(m + p.out_row_offset) * p.out_ld + nb + zoff;
                if (EPI == EPI_F16_HEADMAJOR) {
                    // cross-KV cache row: 64 halfs = eight 16-byte chunks, chunk c stored at position c ^ (m & 7) - the layout the
                    // decoders' ldmatrix reads want (8 consecutive rows of one chunk column fall into 8 different bank groups)
                    __half *o = reinterpret_cast<__half *>(p.out) + idx;
                    const int sw = m & 7, c0 = (nb & 63) >> 3;
                    if (nvalid == 32 && (nb & 7) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                            __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
                            u.z = *reinterpret_cast<uint32_t *>(&h2); u.w = *reinterpret_cast<uint32_t *>(&h3);
                            *reinterpret_cast<uint4 *>(o + (((c0 + (j >> 3)) ^ sw) << 3)) = u;
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) { const int e = (nb & 63) + j; o[(((e >> 3) ^ sw) << 3) | (e & 7)] = __float2half_rn(v[j]); }
                    }
                } else if (EPI == EPI_F16_BIAS || EPI == EPI_F16_BIAS_GELU) {
                    __half *o = reinterpret_cast<__half *>(p.out) + idx;
                    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            __half2 h0 = __floats2half2_rn(v[j], v[j + 1]), h1 = __floats2half2_rn(v[j + 2], v[j + 3]);
                            __half2 h2 = __floats2half2_rn(v[j + 4], v[j + 5]), h3 = __floats2half2_rn(v[j + 6], v[j + 7]);
                            uint4 u;
                            u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
                            u.z = *reinterpret_cast<uint32_t *>(&h2); u.w = *reinterpret_cast<uint32_t *>(&h3);
                            *reinterpret_cast<uint4 *>(o + j) = u;
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) o[j] = __float2half_rn(v[j]);
                    }
                } else {
                    float *o = reinterpret_cast<float *>(p.out) + idx;
                    if (nvalid == 32 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
                        if (EPI == EPI_F32_RESID) {
                            float4 g[8];
#pragma unroll
                            for (int j = 0; j < 8; j++) g[j] = *reinterpret_cast<const float4 *>(o + 4 * j);     // all residual loads in flight
#pragma unroll
                            for (int j = 0; j < 8; j++) *reinterpret_cast<float4 *>(o + 4 * j) = make_float4(v[4 * j] + g[j].x, v[4 * j + 1] + g[j].y, v[4 * j + 2] + g[j].z, v[4 * j + 3] + g[j].w);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    } else {
                        for (int j = 0; j < nvalid; j++) o[j] = EPI == EPI_F32_RESID ? o[j] + v[j] : v[j];
                    }
                }
            }
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

template <int BN, int kStages>
constexpr size_t smem_bytes() { return 1024 + (size_t)kStages * (BM * BK * 2 + BN * BK * 2) + (2 * kStages + 4) * 8 + 16; }

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
template <int EPI>
void configure_epi() { configure_one<128, 4, EPI>(); configure_one<256, 4, EPI>(); }

template <int BN, int kStages, int EPI>
void launch(const GemmOperand &A, const GemmOperand &B, const GemmDev &p, cudaStream_t st) {
    const int g_sms = t_sms;
    CUtensorMap ta, tb;
    make_map(&ta, A, p.K, A.rows, BK, BM);
    make_map(&tb, B, p.K, B.rows, BK, BN);
    const int total = ceil_div(p.N, BN) * ceil_div(p.M, BM) * p.nbatch;
    gemm_tcgen05_kernel<BN, kStages, EPI><<<std::min(total, g_sms), kThreads, smem_bytes<BN, kStages>(), st>>>(ta, tb, p);
    CUDA_CHECK(cudaGetLastError());
}
template <int EPI>
void launch_bn(bool wide, const GemmOperand &A, const GemmOperand &B, const GemmDev &p, cudaStream_t st) {
    if (wide) launch<256, 4, EPI>(A, B, p, st); else launch<128, 4, EPI>(A, B, p, st);
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
    const bool wide = N >= 256 && c256 < c128;
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
