// decoder_mega.cu - the batch-1 autoregressive decoder as ONE persistent cooperative kernel
// (BASELINE.json north_star stage 3: "HBM-bandwidth-bound vectorised kernels with a persistent
// KV-cache"; SURVEY.md §8a rows a8/a9, §7.2 "batch-1 decode latency").
//
// Why one kernel: a decoder step is ~260 dependent mat-vec phases of 0.5-2 us of HBM traffic each.
// As separate launches the dependency latency of every launch dominates (6.8 us per launch measured
// in round-1 v0 = 16 % of the HBM roofline).  Here one CTA per SM stays resident for the whole
// decode loop:
//   * warp 8 (producer) streams this CTA's static slice of the weights and of the cross-attention K/V
//     cache through a 4 x 40 KB shared-memory ring with cp.async.bulk + mbarriers (one bulk copy per
//     16-row chunk).  Its schedule does not depend on activations, so it runs ahead across phase, layer
//     and token boundaries and keeps the HBM pipe busy while the consumers wait for each other.
//   * warps 0..7 (consumers) run the mat-vecs on the tensor cores (ldmatrix + mma.m16n8k16): a weight
//     row is 8 interleaved K slices riding the 8 MMA columns, so a warp owns its row pairs outright -
//     no K split across warps, no cross-warp fold; the finishing lanes publish the rows themselves.
//     The first chunk's A fragments are fetched before the phase's input vector has arrived.
//   * CTAs exchange activations (a few KB per phase) through L2 with a flag-in-data protocol: every
//     value travels in a 64-bit word {epoch, f32} or {epoch, 2 x f16}, written with a single 8-byte
//     store and polled by the readers until the epoch matches.  There are no grid barriers, fences or
//     atomics on the critical path (an all-to-all exchange still costs ~1.5 us on 148 SMs: DESIGN.md §5).
//   * logits filter, greedy sampling and the decoder-state update (whisper_process_logits /
//     whisper_sample_token / whisper_full bookkeeping) are folded in: every CTA
//     reduces its slice of the vocabulary, all CTAs combine the per-CTA records redundantly, so the
//     next token is known everywhere and the loop never returns to the host.
//
// Arithmetic is the oracle's (oracle/whisper_oracle.c wo_decode / process_logits): f16 weights,
// activations rounded to f16 in front of every mat-vec, f32 accumulation (LayerNorm statistics in one
// pass, E[x^2] - mean^2, instead of ggml's two passes: within the logits tolerance).
#include "kernels.h"

namespace ss {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kMegaThreads = kConsumerThreads + 32;
constexpr int kChunkRows = 16;        // weight rows (of d halfs) per ring chunk: one row pair (= one m16 MMA tile) per consumer warp
constexpr int kSlots = 4;
constexpr int kChunkBytes = kChunkRows * 1280 * 2;   // 40960
constexpr int kMaxXs = 5120;          // largest mat-vec input (4 * d, d <= 1280)
constexpr int kMaxRowsPerCta = 512;   // per phase (LM head: n_vocab / grid)
constexpr int kMaxScores = 512;
constexpr int kMaxJ = 5;              // flagged pairs per thread and poll batch
#ifndef SS_MAX_INFLIGHT
#define SS_MAX_INFLIGHT 0      // > 0: the producer keeps at most this many bulk copies outstanding (the ring still buffers kSlots chunks)
#endif
#ifndef SS_NO_STREAM
#define SS_NO_STREAM 0         // 1 (timing experiment only, results are garbage): the producer copies 16 bytes per chunk instead of the chunk
#endif
#ifndef SS_SELF_ONLINE
#define SS_SELF_ONLINE 0     // self-attention as one online (max, sum, P.V) reduction with two CTA barriers; 0: the five-barrier version (faster)
#endif
#ifndef SS_XEXP_INLINE
#define SS_XEXP_INLINE 0     // cross-attention: exp(s - m) inside the P.V loop instead of a separate pass + barrier (slower)
#endif
#ifndef SS_XMMA
#define SS_XMMA 1            // cross-attention scores and P.V on the tensor cores (ldmatrix from the swizzled cross-KV rows)
#endif
#ifndef SS_XATTN8
#define SS_XATTN8 0          // cross-attention: 8 lanes per key row (scores and P.V), 0: one thread per key / one lane per channel pair
#endif
#ifndef SS_NA
#define SS_NA 2
#endif
#ifndef SS_KG
#define SS_KG 2
#endif
// chunks of a phase that go through the tensor cores together (their slice reductions and epilogues overlap), per phase kind.  On 148
// SMs with d = 1280 a CTA owns 1 chunk of O / CQ / CO, 2 of QKV, 3 of FC1 / FC2 and 22 of the LM head.  Measured (same-box A/B, ms per
// token): groups of 2 everywhere 0.679; 1 for the one-chunk phases 0.670; then, one kind at a time: 3 for FC2 0.660 (its three chunks in
// one group), 1 for FC1 0.656 (3: 0.681 - the GELU epilogue of a chunk overlaps the next chunk's tiles), 1 for QKV 0.678 (slower), 3 / 4 for
// the LM head 0.666 (slower), three accumulator chains per tile instead of two: no change.
// An 18-row FC1 chunk (a ninth row pair as a second tile for warp 0, so that 35 rows are 2 chunks): 0.710 - measured, removed.
#ifndef SS_KG_SMALL
#define SS_KG_SMALL 1
#endif
#ifndef SS_KG_FC1
#define SS_KG_FC1 1
#endif
#ifndef SS_KG_FC2
#define SS_KG_FC2 3
#endif
#ifndef SS_KG_LM
#define SS_KG_LM SS_KG
#endif
#ifndef SS_RED3
#define SS_RED3 1            // slice reductions in 3 shuffle levels over the 8 diagonal lanes instead of 5 over the warp (-0.5 %)
#endif
typedef unsigned long long u64;
constexpr int kProfN = 96;
#ifndef SS_MEGA_PROFILE
#define SS_MEGA_PROFILE 0      // 1: per-phase / per-stage cycle counters (tools/mega_prof.py; costs ~13 % of a step)
#endif
constexpr bool kProf = SS_MEGA_PROFILE != 0;
#ifndef SS_MEGA_TRACE
#define SS_MEGA_TRACE 0        // 1: thread 0 of every CTA timestamps (globaltimer) "input complete" / "outputs published" of every phase of ONE step
#endif                         //    into P.prof[cta][kTraceN] (SS_MEGA_TRACE=<file> dumps it; tools/mega_trace.py rebuilds the critical path)
constexpr int kTraceN = 1024, kTraceStep = 40;
enum TracePhase : int { TP_QKV = 0, TP_SELF, TP_O, TP_CQ, TP_CROSS, TP_FOLD, TP_CO, TP_FC1, TP_FC2, TP_COUNT };

enum SegKind : int { SEG_QKV = 0, SEG_O, SEG_CQ, SEG_XK, SEG_XV, SEG_CO, SEG_FC1, SEG_FC2, SEG_LM, SEG_COUNT };

struct DecState {   // replicated per CTA (thread-uniform, lives in shared memory)
    int pos, token, n_sampled, has_ts, seek_delta, result_len, last_id, penult_id, n_kept, failed, completed, done;
};
struct SegTab { int row0, rows, row_bytes, rows_per_chunk, n_chunks, prows; };   // prows: rows of d halfs (FC2: 4 per output row)

struct __align__(128) MegaSmem {
    uint8_t ring[kSlots][kChunkBytes];
    alignas(16) __half xin[2][kMaxXs];   // the f16 mat-vec operand (after LayerNorm where the phase has one), double-buffered by phase parity
    float xs[1280];                 // scratch of the attention / sampling phases (partials, per-CTA records)
    alignas(16) MegaParams P;       // descriptor copy: no pointer chasing through L2 on the critical path
    alignas(16) float acc[kMaxRowsPerCta];      // LM-head rows of this CTA
    alignas(16) float p4[256];                  // FC2: the four quarter-row partial sums of every output row
    alignas(16) float sc[kMaxScores];
    alignas(16) float red[kConsumerWarps][64];
    alignas(16) float red1[32];
    alignas(16) float2 red2[kConsumerWarps];    // LayerNorm {sum, sum of squares} per warp
    alignas(16) float qkv[192];                 // q / current k / current v of this CTA's head
    int redi[32];
    alignas(16) uint64_t full[kSlots];
    uint64_t empty[kSlots];
    SegTab seg[SEG_COUNT];
    DecState st;
    volatile int stop_req;      // consumers -> producer: stop issuing
    volatile int prod_done;     // producer -> consumers: `issued` is final
    volatile uint32_t issued;
    long long trace_last;
    int trace_on, trace_layer;  // (SS_MEGA_TRACE builds) the traced step is running / its current layer
    long long prof[kProfN];     // thread-0 cycle counters: [0..23] totals / per phase kind, [24 + kind * 8 + stage] per-stage breakdown
};

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
__device__ __forceinline__ void trace_event(int phase, int which) {
#if SS_MEGA_TRACE
    MegaSmem &sm = *reinterpret_cast<MegaSmem *>(mega_smem_raw);
    if (threadIdx.x == 0 && sm.trace_on && sm.P.prof != nullptr) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sm.P.prof[(size_t)blockIdx.x * kTraceN + (sm.trace_layer * TP_COUNT + phase) * 2 + which] = (long long)t;
    }
#endif
}
// sub-stage durations of the traced step, summed over the layers: prof[cta][600 + k] += time since the previous mark / event
__device__ __forceinline__ void trace_mark(int k) {
#if SS_MEGA_TRACE
    MegaSmem &sm = *reinterpret_cast<MegaSmem *>(mega_smem_raw);
    if (threadIdx.x == 0 && sm.trace_on && sm.P.prof != nullptr) {
        const long long t = clock64();      // (cycles: %globaltimer costs ~0.3 us per read)
        if (k >= 0) sm.prof[k] += t - sm.trace_last;      // (shared memory; flushed to P.prof[cta][600 + k] when the kernel ends)
        sm.trace_last = t;
    }
#endif
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

// ---- flag-in-data exchange through L2: one 64-bit word {epoch << 32 | float bits} per value -------
__device__ __forceinline__ void ll_store(u64 *p, float v, uint32_t epoch) {
    const u64 w = ((u64)epoch << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load2(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
// f16-valued exchanges (attention outputs, the MLP hidden vector) travel two per word: {epoch << 32 | half2 bits}
__device__ __forceinline__ void ll_store_h2(u64 *p, float lo, float hi, uint32_t epoch) {
    const __half2 h = __floats2half2_rn(lo, hi);
    const u64 w = ((u64)epoch << 32) | (u64)(*reinterpret_cast<const uint32_t *>(&h));
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// poll n_words (even) packed words into dst (shared, 32 bits = one half2 per word); every thread spins only on its own words
__device__ __noinline__ void poll_packed(const u64 *buf, int n_words, uint32_t epoch, uint32_t *dst) {
    const int n2 = n_words >> 1, tid = threadIdx.x;
    const bool prof_on = kProf && SM.P.prof != nullptr && tid == 0;
    const long long t0 = prof_on ? clock64() : 0;
    for (int base = 0; base < n2; base += kConsumerThreads * kMaxJ) {
        ulonglong2 v[kMaxJ];
        bool all;
        do {
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) { const int i = base + tid + j * kConsumerThreads; if (i < n2) v[j] = ll_load2(buf + 2 * i); }
            all = true;
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) {
                const int i = base + tid + j * kConsumerThreads;
                if (i < n2 && ((uint32_t)(v[j].x >> 32) != epoch || (uint32_t)(v[j].y >> 32) != epoch)) all = false;
            }
        } while (!all);
#pragma unroll
        for (int j = 0; j < kMaxJ; j++) {
            const int i = base + tid + j * kConsumerThreads;
            if (i < n2) reinterpret_cast<uint2 *>(dst)[i] = make_uint2((uint32_t)v[j].x, (uint32_t)v[j].y);
        }
    }
    if (prof_on) SM.prof[0] += clock64() - t0;
}
// poll n (even) flagged f32 words into dst (shared; TO_HALF: rounded to f16); every thread spins only on its own words.
// Returns this thread's {sum, sum of squares} of what it fetched (LayerNorm statistics for free).
template <bool TO_HALF>
__device__ __noinline__ float2 poll_vec(const u64 *buf, int n, uint32_t epoch, void *dst) {
    const int n2 = n >> 1, tid = threadIdx.x;
    float s = 0.f, s2 = 0.f;
    const bool prof_on = kProf && SM.P.prof != nullptr && tid == 0;
    const long long t0 = prof_on ? clock64() : 0;
    for (int base = 0; base < n2; base += kConsumerThreads * kMaxJ) {
        ulonglong2 v[kMaxJ];
        bool all;
        do {
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) { const int i = base + tid + j * kConsumerThreads; if (i < n2) v[j] = ll_load2(buf + 2 * i); }
            all = true;
#pragma unroll
            for (int j = 0; j < kMaxJ; j++) {
                const int i = base + tid + j * kConsumerThreads;
                if (i < n2 && ((uint32_t)(v[j].x >> 32) != epoch || (uint32_t)(v[j].y >> 32) != epoch)) all = false;
            }
        } while (!all);
#pragma unroll
        for (int j = 0; j < kMaxJ; j++) {
            const int i = base + tid + j * kConsumerThreads;
            if (i < n2) {
                const float a = __uint_as_float((uint32_t)v[j].x), b = __uint_as_float((uint32_t)v[j].y);
                if (TO_HALF) reinterpret_cast<__half2 *>(dst)[i] = __floats2half2_rn(a, b);
                else { reinterpret_cast<float2 *>(dst)[i] = make_float2(a, b); s += a + b; s2 += a * a + b * b; }
            }
        }
    }
    if (prof_on) SM.prof[0] += clock64() - t0;
    return make_float2(s, s2);
}

// block-wide (256 consumer threads) reductions; result broadcast.  One CTA barrier each: the scratch row alternates
// (`which`), and two uses of the same row are always separated by a barrier of the other one.
__device__ __forceinline__ float consumer_sum(float v, int which = 0) {
    MegaSmem &sm = SM;
    v = warp_sum(v);
    float *r = sm.red1 + 8 * which;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    consumer_sync();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) t += r[i];
    return t;
}
__device__ __forceinline__ float consumer_max(float v, int which = 1) {
    MegaSmem &sm = SM;
    v = warp_max(v);
    float *r = sm.red1 + 8 * which;
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    consumer_sync();
    float t = r[0];
#pragma unroll
    for (int i = 1; i < kConsumerWarps; i++) t = fmaxf(t, r[i]);
    return t;
}

__device__ __forceinline__ float ll_value(const u64 *p) {      // a flagged word this CTA has already seen valid
    u64 w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return __uint_as_float((uint32_t)w);
}

// index, inside the CTA's slice, of the row (FC2: quarter row) that finishing lane `l01` of `warp` publishes in chunk `ch`.
// FC2: warp w works on quarter w & 3 of the rows {2 * (w >> 2), 2 * (w >> 2) + 1} of every 4-row chunk.
// Surplus threads of a per-thread batch re-read a word from the start of the vector (instead of predicating the load off, which
// would push the register arrays to local memory).  They must NOT all pick the same word: 148 CTAs x 4 warps asking one L2
// sector for the same 16 bytes are served one request per clock (measured: tools/xchg_bench2.cu, d = 256 case).
__device__ __forceinline__ int wrap_idx(int i, int n) { return i >= n ? i - n : i; }

template <int KIND>
__device__ __forceinline__ int tile_row(int ch, int warp, int l01) {
    return KIND == SEG_FC2 ? kChunkRows * ch + 4 * (2 * (warp >> 2) + l01) + (warp & 3) : kChunkRows * ch + 2 * warp + l01;
}

// One mat-vec phase, start to finish, specialised on the model width (KS = d / 128 k-steps) and the phase kind:
//   * poll the input vector (every thread a few flagged pairs); LayerNorm where the phase has one (statistics: one shuffle
//     reduction + one CTA barrier; every thread normalises the values it polled); f16 operand to shared memory; one CTA barrier.
//   * the tensor-core part.  A chunk of the ring holds 16 weight rows (of d halfs); consumer warp w owns rows 2w, 2w+1 of
//     every chunk and needs nobody else: a row is viewed as 8 interleaved K slices, so that one m16n8k16 tile =
//     {2 rows} x {8 slices} and column n of the B operand carries the x values of slice n - the wanted products are the
//     diagonal C[8 * row + slice][slice].  Per k-step (128 halfs = 256 B of a row) ldmatrix reads two contiguous 128-byte
//     segments per row (conflict-free, no padding) and the warp's B fragments are 128 consecutive x values (2 per lane,
//     twice), held in registers for the whole phase.  No cross-warp reduction and no CTA barrier: the two finishing lanes
//     of the warp run the epilogue and publish the rows themselves (FC2: the four quarters of a row meet in shared memory).
//   ep_in : epoch the input carries; outputs are published with ep_out;  ph: running phase count (operand buffer parity)
template <int KS, int KIND>
__device__ __noinline__ uint32_t gemv_phase(uint32_t cons, int il, uint32_t ep_in, uint32_t ep_out, uint32_t ph) {
    constexpr int D = KS * 128;
    constexpr int kG = KIND == SEG_QKV ? SS_KG : KIND == SEG_FC1 ? SS_KG_FC1 : KIND == SEG_FC2 ? SS_KG_FC2 : KIND == SEG_LM ? SS_KG_LM : SS_KG_SMALL;
    constexpr bool has_ln = KIND == SEG_QKV || KIND == SEG_CQ || KIND == SEG_FC1 || KIND == SEG_LM;
    constexpr int widx = KIND == SEG_QKV ? 0 : KIND == SEG_O ? 1 : KIND == SEG_CQ ? 2 : KIND == SEG_CO ? 3 : KIND == SEG_FC1 ? 4 : 5;
    constexpr int lidx = KIND == SEG_QKV ? 0 : KIND == SEG_CQ ? 1 : 2;
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = sm.seg[KIND].row0, prows = sm.seg[KIND].prows, n_chunks = sm.seg[KIND].n_chunks;
    __half *buf = sm.xin[ph & 1];
    const bool prof_on = kProf && P.prof != nullptr && tid == 0;
    long long tq = prof_on ? clock64() : 0;
#define SS_STAGE(k) if (prof_on) { const long long tn = clock64(); sm.prof[24 + KIND * 8 + (k)] += tn - tq; tq = tn; }
    // ---- static / already-valid operands of the epilogue, fetched now so that their L2 latency hides behind the poll
    // epilogue role of this lane inside a group of kG chunks: lanes 2g, 2g+1 finish rows 2w, 2w+1 of the group's chunk g
    // (SS_RED3: the slice sums are folded over the 8 diagonal lanes only - xor masks 4, 9, 18 - and diagonal lane number e finishes)
    const int e_idx = SS_RED3 ? ((lane & 3) == (lane >> 3) ? (lane >> 2) : 64) : lane;
    const int eg = e_idx >> 1, el = e_idx & 1;
    float pb = 0.f, pr = 0.f;      // bias / residual of the row this lane publishes in the phase's first group
    float fb = 0.f, fr = 0.f;      // FC2: bias and residual of output row `tid` (folded after the tiles)
    const int tok = sm.st.token, pos = sm.st.pos;
    if (KIND == SEG_FC2) {
        if (tid < sm.seg[KIND].rows) { fb = __ldg(P.layer[il].b[5] + row0 + tid); fr = ll_value(P.xC + row0 + tid); }
    } else if (KIND != SEG_LM && e_idx < 2 * kG) {
        const int R = tile_row<KIND>(eg, warp, el);
        if (R < prows) {
            pb = __ldg(P.layer[il].b[widx] + row0 + R);
            if (KIND == SEG_O) pr = il == 0 ? __half2float(__ldg(P.tok_emb + (size_t)tok * D + row0 + R)) + __ldg(P.d_pos + (size_t)pos * D + row0 + R) : ll_value(P.xA + row0 + R);
            else if (KIND == SEG_CO) pr = ll_value(P.xB + row0 + R);
        }
    }
    // ---- A fragments of the first chunk: the weights are static and the producer runs ahead, so they are fetched
    //      before the input vector has even arrived (ldmatrix.x4 row addresses: lanes 8m..8m+7 feed matrix
    //      m = (row of the pair: m & 1, k half: m >> 1), slice = lane & 7)
    const uint32_t a_off = (uint32_t)tile_row<KIND>(0, warp, (lane >> 3) & 1) * (uint32_t)(2 * D) + 128u * (uint32_t)(lane >> 4) + 16u * (uint32_t)(lane & 7);
    if (n_chunks == 0) return cons;      // (CTA-uniform) no rows of this matrix here: nothing to compute, nothing to publish
    uint32_t af[KS][4];
    {
        const int slot = cons % kSlots;
        if (prof_on) { const long long tw0 = clock64(); mbar_wait(&sm.full[slot], (cons / kSlots) & 1); sm.prof[13] += clock64() - tw0; }
        else mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const uint32_t abase = smem_u32(sm.ring[slot]) + a_off;
#pragma unroll
        for (int j = 0; j < KS; j++)
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(af[j][0]), "=r"(af[j][1]), "=r"(af[j][2]), "=r"(af[j][3]) : "r"(abase + 256u * j));
        // the fragments are in registers: hand the slot back to the producer now (it refills it while the input is polled).
        // Consuming every fragment register first guarantees that the ldmatrix reads have completed.
        uint32_t sink = 0;
#pragma unroll
        for (int j = 0; j < KS; j++) sink ^= af[j][0] ^ af[j][1] ^ af[j][2] ^ af[j][3];
        asm volatile("" ::"r"(sink) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
    }
    // ---- input vector -> f16 operand in shared memory
    if (has_ln) {
#ifdef SS_EXP_HALF_POLL      // timing experiment (garbage results): the LayerNorm phases poll only half of their input words
        constexpr int n2 = D >> 2;
#else
        constexpr int n2 = D >> 1;                                          // flagged pairs
#endif
        constexpr int NK = (n2 + kConsumerThreads - 1) / kConsumerThreads;   // per thread (<= 3)
        const float *lw = KIND == SEG_LM ? P.lnf_w : P.layer[il].lnw[lidx], *lb = KIND == SEG_LM ? P.lnf_b : P.layer[il].lnb[lidx];
        float2 w2[NK], b2[NK], xv[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) {      // (indices clamped instead of predicated: the arrays stay in registers)
            const int i = wrap_idx(tid + k * kConsumerThreads, n2);
            w2[k] = __ldg(reinterpret_cast<const float2 *>(lw) + i); b2[k] = __ldg(reinterpret_cast<const float2 *>(lb) + i);
        }
        if (KIND == SEG_QKV && il == 0) {           // token embedding + positional embedding
            const __half2 *e = reinterpret_cast<const __half2 *>(P.tok_emb + (size_t)tok * D);
            const float2 *pe = reinterpret_cast<const float2 *>(P.d_pos + (size_t)pos * D);
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const int i = wrap_idx(tid + k * kConsumerThreads, n2);
                const float2 ev = __half22float2(__ldg(e + i)), pv = __ldg(pe + i); xv[k] = make_float2(ev.x + pv.x, ev.y + pv.y);
            }
        } else {
            const u64 *src = KIND == SEG_QKV || KIND == SEG_LM ? P.xA : KIND == SEG_CQ ? P.xB : P.xC;
            const long long t0 = prof_on ? clock64() : 0;
            ulonglong2 v[NK];
            bool all;
            do {
#pragma unroll
                for (int k = 0; k < NK; k++) v[k] = ll_load2(src + 2 * wrap_idx(tid + k * kConsumerThreads, n2));
                all = true;
#pragma unroll
                for (int k = 0; k < NK; k++)
                    if ((uint32_t)(v[k].x >> 32) != ep_in || (uint32_t)(v[k].y >> 32) != ep_in) all = false;
            } while (!all);
#pragma unroll
            for (int k = 0; k < NK; k++) xv[k] = make_float2(__uint_as_float((uint32_t)v[k].x), __uint_as_float((uint32_t)v[k].y));
            if (prof_on) sm.prof[0] += clock64() - t0;
        }
        SS_STAGE(0)
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NK; k++) if (tid + k * kConsumerThreads < n2) { s1 += xv[k].x + xv[k].y; s2 += xv[k].x * xv[k].x + xv[k].y * xv[k].y; }
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (lane == 0) sm.red2[warp] = make_float2(s1, s2);
        consumer_sync();
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) { const float2 t = sm.red2[w]; t1 += t.x; t2 += t.y; }
        const float mean = t1 / D, rstd = rsqrtf(fmaxf(t2 / D - mean * mean, 0.f) + 1e-5f);
        SS_STAGE(1)
#pragma unroll
        for (int k = 0; k < NK; k++) {
            const int i = tid + k * kConsumerThreads;
            if (i < n2) reinterpret_cast<__half2 *>(buf)[i] = __floats2half2_rn((xv[k].x - mean) * rstd * w2[k].x + b2[k].x, (xv[k].y - mean) * rstd * w2[k].y + b2[k].y);
        }
    } else {
        const u64 *src = KIND == SEG_O ? P.att1 : KIND == SEG_CO ? P.att2 : P.hbuf;
        poll_packed(src, KIND == SEG_FC2 ? 2 * D : D / 2, ep_in, reinterpret_cast<uint32_t *>(buf));
        SS_STAGE(0)
    }
    consumer_sync();
    trace_event(KIND == SEG_QKV ? TP_QKV : KIND == SEG_O ? TP_O : KIND == SEG_CQ ? TP_CQ : KIND == SEG_CO ? TP_CO : KIND == SEG_FC1 ? TP_FC1 : TP_FC2, 0);
    trace_mark(-1);
    SS_STAGE(2)
    // ---- B fragments: the x side of every k-step, in registers for the whole phase
    uint32_t bf[2 * KS];
    {
        const __half *xq = buf + (KIND == SEG_FC2 ? (warp & 3) * D : 0) + 2 * lane;
#pragma unroll
        for (int j = 0; j < KS; j++) {
            bf[2 * j] = *reinterpret_cast<const uint32_t *>(xq + 128 * j);
            bf[2 * j + 1] = *reinterpret_cast<const uint32_t *>(xq + 128 * j + 64);
        }
    }
    const int g4 = lane >> 2;
    const bool diag = (lane & 3) == (g4 >> 1), odd = (g4 & 1) != 0;
    // where this finishing lane publishes (kinds with one flagged output buffer)
    u64 *outp = (KIND == SEG_O ? P.xB : KIND == SEG_CO ? P.xC : KIND == SEG_CQ ? P.q2 : P.hbuf) + row0 + tile_row<KIND>(0, warp, el);
    const float s4 = P.s4;
    // The chunks of a phase go through the tensor cores in groups of kG: per chunk ldmatrix (chunk 0: done above) -> KS mma -> slot
    // handed back; the slice reductions (5 shuffle levels per row) and the epilogues of the whole group then run interleaved, so a
    // phase pays the shuffle / activation / store latency chain once per group instead of once per chunk.
#pragma unroll 1
    for (int c0 = 0; c0 < n_chunks; c0 += kG) {
        float v0[kG], v1[kG];
#pragma unroll
        for (int g = 0; g < kG; g++) {
            v0[g] = 0.f; v1[g] = 0.f;
            const int ch = c0 + g;
            if (ch < n_chunks) {        // (CTA-uniform)
                if (ch > 0) {
                    const int slot = cons % kSlots;
                    if (prof_on) { const long long tw0 = clock64(); mbar_wait(&sm.full[slot], (cons / kSlots) & 1); sm.prof[13] += clock64() - tw0; }
                    else mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
                    const uint32_t abase = smem_u32(sm.ring[slot]) + a_off;
#pragma unroll
                    for (int j = 0; j < KS; j++)
                        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(af[j][0]), "=r"(af[j][1]), "=r"(af[j][2]), "=r"(af[j][3]) : "r"(abase + 256u * j));
                }
                constexpr int NA = KS >= SS_NA ? SS_NA : KS;     // independent accumulator sets
                float cc[NA][4];
#pragma unroll
                for (int i = 0; i < NA; i++) { cc[i][0] = 0.f; cc[i][1] = 0.f; cc[i][2] = 0.f; cc[i][3] = 0.f; }
#pragma unroll
                for (int j = 0; j < KS; j++)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                                 : "+f"(cc[j % NA][0]), "+f"(cc[j % NA][1]), "+f"(cc[j % NA][2]), "+f"(cc[j % NA][3])
                                 : "r"(af[j][0]), "r"(af[j][1]), "r"(af[j][2]), "r"(af[j][3]), "r"(bf[2 * j]), "r"(bf[2 * j + 1]));
                if (ch > 0) {      // (chunk 0 was released in the prologue)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.empty[cons % kSlots]);      // every ldmatrix of this warp has been consumed by an issued mma
                }
                cons++;
#pragma unroll
                for (int i = 1; i < NA; i++) { cc[0][0] += cc[i][0]; cc[0][1] += cc[i][1]; cc[0][2] += cc[i][2]; cc[0][3] += cc[i][3]; }
                v0[g] = diag ? (odd ? cc[0][1] : cc[0][0]) : 0.f;      // row 2w   : its 8 slices sit on the diagonal
                v1[g] = diag ? (odd ? cc[0][3] : cc[0][2]) : 0.f;      // row 2w+1
            }
        }
        trace_mark(KIND * 8 + 0);      // B fragments, ldmatrix, mma of the group
        if (SS_RED3) {      // diagonal lane of slice g = (a b c): lane bits (a b c a b): flipping c / b / a = xor 4 / 9 / 18
#pragma unroll
            for (int lv = 0; lv < 3; lv++) {
                const int o = lv == 0 ? 4 : lv == 1 ? 9 : 18;
#pragma unroll
                for (int g = 0; g < kG; g++) { v0[g] += __shfl_xor_sync(0xffffffffu, v0[g], o); v1[g] += __shfl_xor_sync(0xffffffffu, v1[g], o); }
            }
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int g = 0; g < kG; g++) { v0[g] += __shfl_xor_sync(0xffffffffu, v0[g], o); v1[g] += __shfl_xor_sync(0xffffffffu, v1[g], o); }
            }
        }
        if (SS_MEGA_TRACE) { float z = 0.f; for (int g = 0; g < kG; g++) z += v0[g] + v1[g]; asm volatile("" ::"f"(z)); }
        trace_mark(KIND * 8 + 1);      // slice reductions
        // every (SS_RED3: every diagonal) lane now holds every row sum of the group: finishing lanes 2g, 2g+1 take chunk c0 + g
        float val = el ? v1[0] : v0[0];
#pragma unroll
        for (int g = 1; g < kG; g++) if (eg == g) val = el ? v1[g] : v0[g];
        const int ch = c0 + eg;
        const int R = tile_row<KIND>(ch, warp, el);
        const bool mine = e_idx < 2 * kG && ch < n_chunks && R < prows;
        float b = pb, r = pr;
        if (c0 > 0 && mine && KIND != SEG_FC2 && KIND != SEG_LM) {      // more chunks per phase than one group (fewer SMs than the design point)
            b = __ldg(P.layer[il].b[widx] + row0 + R); r = 0.f;
            if (KIND == SEG_O) r = il == 0 ? __half2float(__ldg(P.tok_emb + (size_t)tok * D + row0 + R)) + __ldg(P.d_pos + (size_t)pos * D + row0 + R) : ll_value(P.xA + row0 + R);
            else if (KIND == SEG_CO) r = ll_value(P.xB + row0 + R);
        }
        float hid = 0.f, hid_hi = 0.f;
        if (KIND == SEG_FC1) {       // hidden units leave in pairs (this CTA's slice starts and ends on even rows)
            hid = gelu16(val + b);
            hid_hi = __shfl_down_sync(0xffffffffu, hid, SS_RED3 ? 4 : 1);      // the finishing lane of the pair's odd row
        }
        if (mine) {
            if (KIND == SEG_FC2) sm.p4[R] = val;
            else if (KIND == SEG_LM) sm.acc[R] = val;
            else {
                const float v = val + b;
                if (KIND == SEG_QKV) {
                    const int row = row0 + R;
                    if (row < D) ll_store(P.q1 + row, r16(v * s4), ep_out);
                    else if (row < 2 * D) ll_store(P.kcur + (row - D), r16(v * s4), ep_out);          // (the head's CTA appends k / v to the self-KV
                    else ll_store(P.vcur + (row - 2 * D), r16(v), ep_out);                             //  cache itself: see self_attn)
                } else if (KIND == SEG_O || KIND == SEG_CO) ll_store(outp + kChunkRows * ch, r + v, ep_out);
                else if (KIND == SEG_CQ) ll_store(outp + kChunkRows * ch, r16(v * s4), ep_out);
                else if (el == 0) ll_store_h2(P.hbuf + ((row0 + R) >> 1), hid, hid_hi, ep_out);
            }
        }
        trace_mark(KIND * 8 + 2);      // epilogue + stores
    }
    if (prof_on) sm.prof[1] += clock64() - tq;
    SS_STAGE(3)
    if (KIND == SEG_FC2) {      // fold the four quarters of every output row
        consumer_sync();
        if (tid < sm.seg[KIND].rows) {
            const float4 q = *reinterpret_cast<const float4 *>(&sm.p4[4 * tid]);
            ll_store(P.xA + row0 + tid, fr + (((q.x + q.y) + q.z) + q.w) + fb, ep_out);
        }
        SS_STAGE(4)
    } else if (KIND == SEG_LM) consumer_sync();     // lm_epilogue reads other warps' rows
    if (KIND != SEG_LM) trace_event(KIND == SEG_QKV ? TP_QKV : KIND == SEG_O ? TP_O : KIND == SEG_CQ ? TP_CQ : KIND == SEG_CO ? TP_CO : KIND == SEG_FC1 ? TP_FC1 : TP_FC2, 1);
#undef SS_STAGE
    return cons;
}

// scores of one query against n key rows (128 B each) at `K`: 8 lanes per row, 4 rows per warp per step.
// FROM_RING: rows come from a ring slot (shared memory), else from L2 with 4 rows per thread in flight.
template <bool FROM_RING>
__device__ __forceinline__ float attn_scores(const uint8_t *K, int n, int sc_base, const float4 &qa, const float4 &qb, float lmax, int row0 = -1) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    for (int jb = 0; jb < n; jb += 128) {
        uint4 kv[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) kv[u] = FROM_RING ? *reinterpret_cast<const uint4 *>(K + (size_t)j * 128 + ((row0 >= 0 ? l8 ^ ((row0 + j) & 7) : l8) << 4)) : __ldcg(reinterpret_cast<const uint4 *>(K + (size_t)j * 128 + l8 * 16));
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
// ---- cross-attention out of a ring slot: one thread per key for the scores (its 128-byte row is read as eight 16-byte
// pieces in an order rotated by the key index: the 32 lanes of a load touch every bank group equally), one lane per
// channel pair and one warp per key residue for P.V
__device__ __forceinline__ float xattn_scores(const uint8_t *K, int n, int sc_base, float lmax, int row0) {
    MegaSmem &sm = SM;
    for (int j = threadIdx.x; j < n; j += kConsumerThreads) {
        const uint8_t *row = K + (size_t)j * 128;
        const int sw = (row0 + j) & 7;      // chunk c of key m sits at position c ^ (m & 7) (cross-KV cache layout)
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            const int c0 = (c + j) & 7, c1 = (c + 1 + j) & 7;
            const uint4 k0 = *reinterpret_cast<const uint4 *>(row + (c0 ^ sw) * 16), k1 = *reinterpret_cast<const uint4 *>(row + (c1 ^ sw) * 16);
            a0 = dot8(k0, *reinterpret_cast<const float4 *>(sm.qkv + c0 * 8), *reinterpret_cast<const float4 *>(sm.qkv + c0 * 8 + 4), a0);
            a1 = dot8(k1, *reinterpret_cast<const float4 *>(sm.qkv + c1 * 8), *reinterpret_cast<const float4 *>(sm.qkv + c1 * 8 + 4), a1);
        }
        const float ds = a0 + a1;
        sm.sc[sc_base + j] = ds;
        lmax = fmaxf(lmax, ds);
    }
    return lmax;
}
__device__ __forceinline__ void xattn_pv_p(const uint8_t *V, int n, int sc_base, float &o0, float &o1, int row0) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cch = lane >> 2, cin = (lane & 3) * 4;      // this lane's channel pair: chunk cch, byte cin inside it
    auto vptr = [&](int j) { return V + (size_t)j * 128 + ((cch ^ ((row0 + j) & 7)) << 4) + cin; };
    float b0 = 0.f, b1 = 0.f;
    int j = warp;
    for (; j + 8 < n; j += 16) {
        const float2 va = h2f(*reinterpret_cast<const uint32_t *>(vptr(j))), vb = h2f(*reinterpret_cast<const uint32_t *>(vptr(j + 8)));
        const float pa = sm.sc[sc_base + j], pb = sm.sc[sc_base + j + 8];
        o0 = fmaf(pa, va.x, o0); o1 = fmaf(pa, va.y, o1); b0 = fmaf(pb, vb.x, b0); b1 = fmaf(pb, vb.y, b1);
    }
    if (j < n) { const float2 va = h2f(*reinterpret_cast<const uint32_t *>(vptr(j))); const float pa = sm.sc[sc_base + j]; o0 = fmaf(pa, va.x, o0); o1 = fmaf(pa, va.y, o1); }
    o0 += b0; o1 += b1;
}

// P.V with the soft-max numerator computed in place: every lane of a warp turns the warp's raw scores into exp(s - m) itself
// (one MUFU per key and lane), so no separate exp pass / barrier; `lsum` accumulates the warp's share of the denominator
__device__ __forceinline__ void xattn_pv(const uint8_t *V, int n, int sc_base, float m, float &o0, float &o1, float &lsum, int row0) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cch = lane >> 2, cin = (lane & 3) * 4;      // this lane's channel pair: chunk cch, byte cin inside it
    auto vptr = [&](int j) { return V + (size_t)j * 128 + ((cch ^ ((row0 + j) & 7)) << 4) + cin; };
    float b0 = 0.f, b1 = 0.f, l1 = 0.f;
    int j = warp;
    for (; j + 8 < n; j += 16) {
        const float2 va = h2f(*reinterpret_cast<const uint32_t *>(vptr(j))), vb = h2f(*reinterpret_cast<const uint32_t *>(vptr(j + 8)));
        const float pa = __expf(sm.sc[sc_base + j] - m), pb = __expf(sm.sc[sc_base + j + 8] - m);
        lsum += pa; l1 += pb;
        o0 = fmaf(pa, va.x, o0); o1 = fmaf(pa, va.y, o1); b0 = fmaf(pb, vb.x, b0); b1 = fmaf(pb, vb.y, b1);
    }
    if (j < n) { const float2 va = h2f(*reinterpret_cast<const uint32_t *>(vptr(j))); const float pa = __expf(sm.sc[sc_base + j] - m); lsum += pa; o0 = fmaf(pa, va.x, o0); o1 = fmaf(pa, va.y, o1); }
    o0 += b0; o1 += b1; lsum += l1;
}

// 8 lanes per key row, 4 rows per warp and step (as attn_scores); exp(s - m) in place; every lane of an 8-lane group adds the
// same numerators, so `lsum` is the group's share of the denominator
__device__ __forceinline__ void xattn_pv8(const uint8_t *V, int n, int sc_base, float m, float (&acc)[8], float &lsum, int row0) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    for (int jb = 0; jb < n; jb += 128) {
        uint4 vv[4]; float pp[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = jb + u * 32 + warp * 4 + sub;
            if (j < n) { vv[u] = *reinterpret_cast<const uint4 *>(V + (size_t)j * 128 + ((l8 ^ ((row0 + j) & 7)) << 4)); pp[u] = __expf(sm.sc[sc_base + j] - m); }
            else { vv[u] = make_uint4(0, 0, 0, 0); pp[u] = 0.f; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { lsum += pp[u]; axpy8(vv[u], pp[u], acc); }
    }
}
// ---- cross-attention on the tensor cores.  A slot holds n key (value) rows of 128 bytes in cache layout: chunk c of key m at
// position c ^ (m & 7), so the 8 rows of an ldmatrix 8x8 tile hit 8 different 16-byte bank groups.
// scores: m16n8k16 tiles = {16 keys} x {q replicated on the 8 columns}; warp w takes key tiles w, w + 8, ...; thread (g, t) ends
// up with the score of key g (c0) and key g + 8 (c2) of the tile.
__device__ __forceinline__ float xmma_scores(const uint8_t *K, int n, int sc_base, float lmax, int row0, const uint32_t (&qb)[8]) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t kb = smem_u32(K);
    const int lr = (lane & 7) + ((lane >> 3) & 1) * 8, lc = lane >> 4;      // ldmatrix.x4: this lane addresses row lr, chunk 2s + lc
    for (int i = warp; i * 16 < n; i += kConsumerWarps) {
        const int r = 16 * i + lr, sw = (row0 + r) & 7;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t a[4][4];
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++)      // all four loads in flight before the first mma
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(a[s4][0]), "=r"(a[s4][1]), "=r"(a[s4][2]), "=r"(a[s4][3]) : "r"(kb + (uint32_t)r * 128u + (uint32_t)(((2 * s4 + lc) ^ sw) << 4)));
#pragma unroll
        for (int s4 = 0; s4 < 4; s4 += 2) {   // two accumulator chains
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[s4][0]), "r"(a[s4][1]), "r"(a[s4][2]), "r"(a[s4][3]), "r"(qb[2 * s4]), "r"(qb[2 * s4 + 1]));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c2[0]), "+f"(c2[1]), "+f"(c2[2]), "+f"(c2[3]) : "r"(a[s4 + 1][0]), "r"(a[s4 + 1][1]), "r"(a[s4 + 1][2]), "r"(a[s4 + 1][3]), "r"(qb[2 * s4 + 2]), "r"(qb[2 * s4 + 3]));
        }
        c[0] += c2[0]; c[2] += c2[2];
        if (t == 0) {
            const int j0 = 16 * i + g, j1 = j0 + 8;
            if (j0 < n) { sm.sc[sc_base + j0] = c[0]; lmax = fmaxf(lmax, c[0]); }
            if (j1 < n) { sm.sc[sc_base + j1] = c[2]; lmax = fmaxf(lmax, c[2]); }
        }
    }
    return lmax;
}
// P.V: m16n8k16 with A = the probabilities (row 0 of the tile, f16 like ggml's mat-mul operand; rows 1..15 zero), B = 16 value rows x
// 8 channels through ldmatrix.trans; warp w takes key blocks w, w + 8, ...; lanes 0..3 end up with channels 8 nt + 2 t, + 1.
__device__ __forceinline__ void xmma_pv(const uint8_t *V, int n, int sc_base, int row0, float (&o)[8][4]) {
    MegaSmem &sm = SM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t vb = smem_u32(V);
    const int lr = (lane & 7) + ((lane >> 3) & 1) * 8, lc = lane >> 4;
    for (int ks = warp; ks * 16 < n; ks += kConsumerWarps) {
        const int k0 = 16 * ks + 2 * t;
        uint32_t a0 = 0u, a2 = 0u;
        if (g == 0) {
            const float p0 = k0 < n ? sm.sc[sc_base + k0] : 0.f, p1 = k0 + 1 < n ? sm.sc[sc_base + k0 + 1] : 0.f;
            const float p2 = k0 + 8 < n ? sm.sc[sc_base + k0 + 8] : 0.f, p3 = k0 + 9 < n ? sm.sc[sc_base + k0 + 9] : 0.f;
            const __half2 h0 = __floats2half2_rn(p0, p1), h2 = __floats2half2_rn(p2, p3);
            a0 = *reinterpret_cast<const uint32_t *>(&h0); a2 = *reinterpret_cast<const uint32_t *>(&h2);
        }
        // rows past the end of the slice hold whatever the ring slot held before (possibly NaN bit patterns): 0 * NaN must not happen
        const uint32_t m0 = (k0 < n ? 0x0000ffffu : 0u) | (k0 + 1 < n ? 0xffff0000u : 0u), m1 = (k0 + 8 < n ? 0x0000ffffu : 0u) | (k0 + 9 < n ? 0xffff0000u : 0u);
        const int r = 16 * ks + lr, sw = (row0 + r) & 7;
        uint32_t b[4][4];
#pragma unroll
        for (int np = 0; np < 4; np++)      // all four loads in flight before the first mma
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(b[np][0]), "=r"(b[np][1]), "=r"(b[np][2]), "=r"(b[np][3]) : "r"(vb + (uint32_t)r * 128u + (uint32_t)(((2 * np + lc) ^ sw) << 4)));
#pragma unroll
        for (int np = 0; np < 4; np++) {
            b[np][0] &= m0; b[np][1] &= m1; b[np][2] &= m0; b[np][3] &= m1;
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(o[2 * np][0]), "+f"(o[2 * np][1]), "+f"(o[2 * np][2]), "+f"(o[2 * np][3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b[np][0]), "r"(b[np][1]));
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(o[2 * np + 1][0]), "+f"(o[2 * np + 1][1]), "+f"(o[2 * np + 1][2]), "+f"(o[2 * np + 1][3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b[np][2]), "r"(b[np][3]));
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

// (v1: separate max / sum / P.V reductions, five CTA barriers - measured 1.5 % faster than the online variant below)
// self-attention, one CTA per head: past keys/values from the KV cache in L2 (the first 128 positions are fetched into
// registers BEFORE the poll: their addresses do not depend on the current token, so the L2 latency hides behind the wait
// for q), the current token's q/k/v from the flagged exchange buffers
__device__ __noinline__ void self_attn_v1(int il, uint32_t ep) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int h = blockIdx.x, tid = threadIdx.x;
    if (h >= P.H) return;
    const int n_past = sm.st.pos, d = P.d, l8 = tid & 7, r8 = tid >> 3;
    const bool prof_on = kProf && P.prof != nullptr && tid == 0;
    long long tq = prof_on ? clock64() : 0;
#undef SS_STAGE
#define SS_STAGE(kd, k) if (prof_on) { const long long tn = clock64(); sm.prof[24 + (kd) * 8 + (k)] += tn - tq; tq = tn; }
    const uint8_t *Kh = reinterpret_cast<const uint8_t *>(P.self_k + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    const uint8_t *Vh = reinterpret_cast<const uint8_t *>(P.self_v + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    uint4 kr[4], vr[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int j = min(u * 32 + r8, max(n_past - 1, 0));      // clamped: rows past the end are fetched but not used
        kr[u] = __ldcg(reinterpret_cast<const uint4 *>(Kh + (size_t)j * 128 + l8 * 16));
        vr[u] = __ldcg(reinterpret_cast<const uint4 *>(Vh + (size_t)j * 128 + l8 * 16));
    }
    if (tid < 96) {   // 3 x 64 flagged floats: q, k, v of this head
        const int which = tid >> 5, i2 = tid & 31;
        const u64 *src = (which == 0 ? P.q1 : which == 1 ? P.kcur : P.vcur) + h * 64 + 2 * i2;
        ulonglong2 v;
        do { v = ll_load2(src); } while ((uint32_t)(v.x >> 32) != ep || (uint32_t)(v.y >> 32) != ep);
        sm.qkv[which * 64 + 2 * i2] = __uint_as_float((uint32_t)v.x); sm.qkv[which * 64 + 2 * i2 + 1] = __uint_as_float((uint32_t)v.y);
    }
    consumer_sync();
    trace_event(TP_SELF, 0);
    SS_STAGE(SEG_XV, 0)
    const float4 qa = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8), qb = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8 + 4);
    float lmax = -INFINITY;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int j = u * 32 + r8;
        float ds = dot8(kr[u], qa, qb, 0.f);
        ds += __shfl_xor_sync(0xffffffffu, ds, 1);
        ds += __shfl_xor_sync(0xffffffffu, ds, 2);
        ds += __shfl_xor_sync(0xffffffffu, ds, 4);
        if (j < n_past) { if (l8 == 0) sm.sc[j] = ds; lmax = fmaxf(lmax, ds); }
    }
    if (n_past > 128) lmax = attn_scores<false>(Kh + 128 * 128, n_past - 128, 128, qa, qb, lmax);
    if (tid < 32) {   // the current token's own key
        float ds = sm.qkv[tid] * sm.qkv[64 + tid] + sm.qkv[32 + tid] * sm.qkv[96 + tid];
        ds = warp_sum(ds);
        if (tid == 0) sm.sc[n_past] = ds;
        lmax = fmaxf(lmax, ds);
    }
    SS_STAGE(SEG_XV, 1)
    const float m = consumer_max(lmax);
    float lsum = 0.f;
    for (int j = tid; j <= n_past; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; lsum += e; }
    const float l = consumer_sum(lsum);
    SS_STAGE(SEG_XV, 2)
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int u = 0; u < 4; u++) { const int j = u * 32 + r8; if (j < n_past) axpy8(vr[u], sm.sc[j], acc); }
    if (n_past > 128) attn_pv<false>(Vh + 128 * 128, n_past - 128, 128, acc);
    SS_STAGE(SEG_XV, 3)
    float o = attn_fold(acc);
    if (tid < 64) {
        o = (o + sm.sc[n_past] * sm.qkv[128 + tid]) / l;
        const float o_hi = __shfl_down_sync(0xffffffffu, o, 1);
        if ((tid & 1) == 0) ll_store_h2(P.att1 + h * 32 + (tid >> 1), o, o_hi, ep);
    }
    trace_event(TP_SELF, 1);
    // KV-cache append by the CTA that will read it: the rows of head h are written and (from the next token on) read by this CTA
    // only, ordered by program order + CTA barriers - no cross-CTA visibility assumption (the flagged k / v words are f16 values)
    if (tid < 64) {
        const int which = tid >> 5, i2 = tid & 31;      // 0: k, 1: v
        __half *row = (which == 0 ? P.self_k : P.self_v) + (size_t)il * P.ctx * d + ((size_t)h * P.ctx + n_past) * 64;
        reinterpret_cast<__half2 *>(row)[i2] = __floats2half2_rn(sm.qkv[64 + which * 64 + 2 * i2], sm.qkv[64 + which * 64 + 2 * i2 + 1]);
    }
    SS_STAGE(SEG_XV, 4)
}

// self-attention, one CTA per head: past keys/values from the KV cache in L2 (the first 128 positions are fetched into
// registers BEFORE the poll: their addresses do not depend on the current token, so the L2 latency hides behind the wait
// for q), the current token's q/k/v from the flagged exchange buffers.  Soft-max and P.V run as ONE online reduction
// (every 8-lane group keeps {max, sum, 8 channels} of its keys; groups merge by shuffles, warps through shared memory):
// two CTA barriers per call instead of five.
__device__ __forceinline__ void osm_merge(float &m, float &l, float (&acc)[8], float m2, float l2, const float (&a2)[8]) {
    const float mn = fmaxf(m, m2);
    const float s1 = m > -INFINITY ? __expf(m - mn) : 0.f, s2 = m2 > -INFINITY ? __expf(m2 - mn) : 0.f;
    l = l * s1 + l2 * s2;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = acc[i] * s1 + a2[i] * s2;
    m = mn;
}
__device__ __noinline__ void self_attn_online(int il, uint32_t ep) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int h = blockIdx.x, tid = threadIdx.x;
    if (h >= P.H) return;
    const int n_past = sm.st.pos, d = P.d, l8 = tid & 7, r8 = tid >> 3, warp = tid >> 5, lane = tid & 31;
    const bool prof_on = kProf && P.prof != nullptr && tid == 0;
    long long tq = prof_on ? clock64() : 0;
#undef SS_STAGE
#define SS_STAGE(kd, k) if (prof_on) { const long long tn = clock64(); sm.prof[24 + (kd) * 8 + (k)] += tn - tq; tq = tn; }
    const uint8_t *Kh = reinterpret_cast<const uint8_t *>(P.self_k + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    const uint8_t *Vh = reinterpret_cast<const uint8_t *>(P.self_v + (size_t)il * P.ctx * d + (size_t)h * P.ctx * 64);
    uint4 kr[4], vr[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int j = min(u * 32 + r8, max(n_past - 1, 0));      // clamped: rows past the end are fetched but not used
        kr[u] = __ldcg(reinterpret_cast<const uint4 *>(Kh + (size_t)j * 128 + l8 * 16));
        vr[u] = __ldcg(reinterpret_cast<const uint4 *>(Vh + (size_t)j * 128 + l8 * 16));
    }
    if (tid < 96) {   // 3 x 64 flagged floats: q, k, v of this head
        const int which = tid >> 5, i2 = tid & 31;
        const u64 *src = (which == 0 ? P.q1 : which == 1 ? P.kcur : P.vcur) + h * 64 + 2 * i2;
        ulonglong2 v;
        do { v = ll_load2(src); } while ((uint32_t)(v.x >> 32) != ep || (uint32_t)(v.y >> 32) != ep);
        sm.qkv[which * 64 + 2 * i2] = __uint_as_float((uint32_t)v.x); sm.qkv[which * 64 + 2 * i2 + 1] = __uint_as_float((uint32_t)v.y);
    }
    consumer_sync();
    trace_event(TP_SELF, 0);
    trace_mark(-1);
    SS_STAGE(SEG_XV, 0)
    const float4 qa = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8), qb = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8 + 4);
    float m = -INFINITY, l = 0.f, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int jb = 0; jb < n_past; jb += 128) {      // (one trip unless the window has more than 128 tokens)
        if (jb > 0) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int j = min(jb + u * 32 + r8, n_past - 1);
                kr[u] = __ldcg(reinterpret_cast<const uint4 *>(Kh + (size_t)j * 128 + l8 * 16));
                vr[u] = __ldcg(reinterpret_cast<const uint4 *>(Vh + (size_t)j * 128 + l8 * 16));
            }
        }
        float sv[4], bm = -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float ds = dot8(kr[u], qa, qb, 0.f);
            ds += __shfl_xor_sync(0xffffffffu, ds, 1);
            ds += __shfl_xor_sync(0xffffffffu, ds, 2);
            ds += __shfl_xor_sync(0xffffffffu, ds, 4);
            sv[u] = jb + u * 32 + r8 < n_past ? ds : -INFINITY;
            bm = fmaxf(bm, sv[u]);
        }
        if (bm > -INFINITY) {       // (uniform inside an 8-lane group)
            float bl = 0.f, ba[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 4; u++) { const float pu = sv[u] > -INFINITY ? __expf(sv[u] - bm) : 0.f; bl += pu; axpy8(vr[u], pu, ba); }
            osm_merge(m, l, acc, bm, bl, ba);
        }
    }
    trace_mark(90);      // scores + online soft-max of the group
    // the four 8-lane groups of a warp
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
        float a2[8];
#pragma unroll
        for (int i = 0; i < 8; i++) a2[i] = __shfl_xor_sync(0xffffffffu, acc[i], o);
        osm_merge(m, l, acc, m2, l2, a2);
    }
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 8; i++) sm.red[warp][l8 * 8 + i] = acc[i];
        if (lane == 0) sm.red2[warp] = make_float2(m, l);
    }
    // the current token's own key (warps 0 and 1 each need it for their channels)
    float s_self = 0.f;
    if (tid < 64) { s_self = sm.qkv[lane] * sm.qkv[64 + lane] + sm.qkv[32 + lane] * sm.qkv[96 + lane]; s_self = warp_sum(s_self); }
    SS_STAGE(SEG_XV, 1)
    trace_mark(91);      // group merges, own key
    consumer_sync();
    trace_mark(92);      // barrier
    SS_STAGE(SEG_XV, 2)
    if (tid < 64) {
        float M = s_self;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) M = fmaxf(M, sm.red2[w].x);
        const float e0 = __expf(s_self - M);
        float L = e0, o = e0 * sm.qkv[128 + tid];
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) {
            const float2 ml = sm.red2[w];
            if (ml.x > -INFINITY) { const float e = __expf(ml.x - M); L += ml.y * e; o += sm.red[w][tid] * e; }
        }
        o /= L;
        const float o_hi = __shfl_down_sync(0xffffffffu, o, 1);
        if ((tid & 1) == 0) ll_store_h2(P.att1 + h * 32 + (tid >> 1), o, o_hi, ep);
    }
    trace_mark(93);      // final merge + store
    trace_event(TP_SELF, 1);
    // KV-cache append by the CTA that will read it: the rows of head h are written and (from the next token on) read by this CTA
    // only, ordered by program order + CTA barriers - no cross-CTA visibility assumption (the flagged k / v words are f16 values)
    if (tid < 64) {
        const int which = tid >> 5, i2 = tid & 31;      // 0: k, 1: v
        __half *row = (which == 0 ? P.self_k : P.self_v) + (size_t)il * P.ctx * d + ((size_t)h * P.ctx + n_past) * 64;
        reinterpret_cast<__half2 *>(row)[i2] = __floats2half2_rn(sm.qkv[64 + which * 64 + 2 * i2], sm.qkv[64 + which * 64 + 2 * i2 + 1]);
    }
    SS_STAGE(SEG_XV, 4)
}

__device__ __forceinline__ void self_attn(int il, uint32_t ep) { if (SS_SELF_ONLINE) self_attn_online(il, ep); else self_attn_v1(il, ep); }

// cross-attention; K then V of this CTA's (head, key split) streamed through the ring; split 0 of every head
// folds the head's partials into the attention vector
__device__ __noinline__ uint32_t cross_attn(uint32_t cons, uint32_t ep) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const SegTab sk = sm.seg[SEG_XK];
    if (sk.n_chunks == 0) return cons;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ns = P.xsplit, h = blockIdx.x / ns, sp = blockIdx.x % ns, n = sk.rows;
    const bool prof_on = kProf && P.prof != nullptr && tid == 0;
    long long tq = prof_on ? clock64() : 0;
    if (tid < 32) {
        const u64 *src = P.q2 + h * 64 + 2 * tid;
        ulonglong2 v;
        do { v = ll_load2(src); } while ((uint32_t)(v.x >> 32) != ep || (uint32_t)(v.y >> 32) != ep);
        sm.qkv[2 * tid] = __uint_as_float((uint32_t)v.x); sm.qkv[2 * tid + 1] = __uint_as_float((uint32_t)v.y);
    }
    consumer_sync();
    trace_event(TP_CROSS, 0);
    trace_mark(-1);
    SS_STAGE(SEG_XK, 0)
    float lmax = -INFINITY;
    uint32_t qfrag[8];      // (SS_XMMA) B fragments of q: k-step s4 covers dims 16 s4 .. 16 s4 + 15; q values are f16-valued
    {
        const int t = lane & 3;
#pragma unroll
        for (int s4 = 0; s4 < 4; s4++) {
            const __half2 h0 = __floats2half2_rn(sm.qkv[16 * s4 + 2 * t], sm.qkv[16 * s4 + 2 * t + 1]);
            const __half2 h1 = __floats2half2_rn(sm.qkv[16 * s4 + 2 * t + 8], sm.qkv[16 * s4 + 2 * t + 9]);
            qfrag[2 * s4] = *reinterpret_cast<const uint32_t *>(&h0); qfrag[2 * s4 + 1] = *reinterpret_cast<const uint32_t *>(&h1);
        }
    }
    for (int ch = 0; ch < sk.n_chunks; ch++) {
        const int slot = cons % kSlots;
        if (prof_on) { const long long tw0 = clock64(); mbar_wait(&sm.full[slot], (cons / kSlots) & 1); sm.prof[14] += clock64() - tw0; }
        else mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const int kbase = ch * sk.rows_per_chunk, nk = min(sk.rows_per_chunk, n - kbase);
        if (SS_XMMA) lmax = xmma_scores(sm.ring[slot], nk, kbase, lmax, sk.row0 + kbase, qfrag);
        else if (SS_XATTN8) {
            const int l8 = lane & 7;
            const float4 qa = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8), qb = *reinterpret_cast<const float4 *>(sm.qkv + l8 * 8 + 4);
            lmax = attn_scores<true>(sm.ring[slot], nk, kbase, qa, qb, lmax, sk.row0 + kbase);
        } else lmax = xattn_scores(sm.ring[slot], nk, kbase, lmax, sk.row0 + kbase);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
    SS_STAGE(SEG_XK, 1)
    trace_mark(80);      // scores
    const float m = consumer_max(lmax);        // its barrier also publishes the raw scores
    trace_mark(81);      // max
    float l_cta = 0.f;
    if (SS_XMMA || !SS_XEXP_INLINE) {
        float ls = 0.f;
        for (int j = tid; j < n; j += kConsumerThreads) { const float e = __expf(sm.sc[j] - m); sm.sc[j] = e; ls += e; }
        l_cta = consumer_sum(ls);        // its barrier also publishes the probabilities
    }
    SS_STAGE(SEG_XK, 2)
    float o0 = 0.f, o1 = 0.f, lsum = 0.f, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float om[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) { om[i][0] = 0.f; om[i][1] = 0.f; om[i][2] = 0.f; om[i][3] = 0.f; }
    for (int ch = 0; ch < sk.n_chunks; ch++) {     // the V slice has the same chunking as the K slice
        const int slot = cons % kSlots;
        mbar_wait(&sm.full[slot], (cons / kSlots) & 1);
        const int kbase = ch * sk.rows_per_chunk, nk = min(sk.rows_per_chunk, n - kbase);
        if (SS_XMMA) xmma_pv(sm.ring[slot], nk, kbase, sk.row0 + kbase, om);
        else if (SS_XATTN8) xattn_pv8(sm.ring[slot], nk, kbase, m, acc, lsum, sk.row0 + kbase);
        else if (SS_XEXP_INLINE) xattn_pv(sm.ring[slot], nk, kbase, m, o0, o1, lsum, sk.row0 + kbase);
        else xattn_pv_p(sm.ring[slot], nk, kbase, o0, o1, sk.row0 + kbase);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[slot]);
        cons++;
    }
    SS_STAGE(SEG_XK, 3)
    trace_mark(82);      // P.V
    float o = 0.f;
    if (SS_XMMA) {
        if (lane < 4) {      // row 0 of the accumulator tiles: channels 8 nt + 2 lane, + 1
#pragma unroll
            for (int nt = 0; nt < 8; nt++) *reinterpret_cast<float2 *>(&sm.red[tid >> 5][8 * nt + 2 * lane]) = make_float2(om[nt][0], om[nt][1]);
        }
        consumer_sync();
        if (tid < 64) for (int w = 0; w < kConsumerWarps; w++) o += sm.red[w][tid];
    } else if (SS_XATTN8) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 8); lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
        if (lane == 0) sm.red2[tid >> 5].x = lsum;
        o = attn_fold(acc);
    } else {
        *reinterpret_cast<float2 *>(&sm.red[tid >> 5][2 * lane]) = make_float2(o0, o1);
        if (lane == 0) sm.red2[tid >> 5].x = lsum;      // (every lane of a warp holds the same sum)
        consumer_sync();
        if (tid < 64) for (int w = 0; w < kConsumerWarps; w++) o += sm.red[w][tid];
    }
    u64 *out = P.part + ((size_t)h * ns + sp) * 66;
    if (tid < 64) ll_store(out + 2 + tid, o, ep);
    if (tid == 0) {
        float l = l_cta;
        if (!SS_XMMA && (SS_XEXP_INLINE || SS_XATTN8)) { l = 0.f; for (int w = 0; w < kConsumerWarps; w++) l += sm.red2[w].x; }
        ll_store(out, m, ep); ll_store(out + 1, l, ep);
    }
    trace_mark(83);      // fold + stores
    trace_event(TP_CROSS, 1);
    SS_STAGE(SEG_XK, 4)
    if (sp != 0) return cons;
    // ---- split 0 folds all ns (<= 8) partial records of head h
    poll_vec<false>(P.part + (size_t)h * ns * 66, ns * 66, ep, sm.xs);
    consumer_sync();
    trace_event(TP_FOLD, 0);
    if (tid < 64) {      // (ns <= 8: all loads first, independent exponentials, no loop-carried latency chains)
        float pm[8], pl[8], po[8];
#pragma unroll
        for (int s2 = 0; s2 < 8; s2++) {
            const bool on = s2 < ns;
            pm[s2] = on ? sm.xs[s2 * 66] : -INFINITY; pl[s2] = on ? sm.xs[s2 * 66 + 1] : 0.f; po[s2] = on ? sm.xs[s2 * 66 + 2 + tid] : 0.f;
        }
        float M = pm[0];
#pragma unroll
        for (int s2 = 1; s2 < 8; s2++) M = fmaxf(M, pm[s2]);
        float Lsum = 0.f, oo = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < 8; s2++) { const float e = pm[s2] > -INFINITY ? __expf(pm[s2] - M) : 0.f; Lsum = fmaf(pl[s2], e, Lsum); oo = fmaf(po[s2], e, oo); }
        const float a = oo / Lsum, a_hi = __shfl_down_sync(0xffffffffu, a, 1);
        if ((tid & 1) == 0) ll_store_h2(P.att2 + h * 32 + (tid >> 1), a, a_hi, ep);
    }
    trace_event(TP_FOLD, 1);
    consumer_sync();
    SS_STAGE(SEG_XK, 5)
#undef SS_STAGE
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
__device__ __noinline__ void lm_epilogue(bool keep, bool sampling, uint32_t ep) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = sm.seg[SEG_LM].rows, row0 = sm.seg[SEG_LM].row0;
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
    sa = consumer_sum(sa, 0);
    sb = consumer_sum(sb, 1);
    if (tid < 8) {
        u64 *rec = P.stats + (size_t)blockIdx.x * 8;
        const float vals[8] = {mt.v, __int_as_float(mt.i), ms.v, __int_as_float(ms.i), sa, sb, 0.f, 0.f};
        ll_store(rec + tid, vals[tid], ep);
    }
}

// combine the per-CTA records (every CTA, identically), greedy sample (whisper_sample_token best=true) and
// the per-token decoder bookkeeping of whisper_full; updates sm.st
__device__ __noinline__ void sample_and_update(int seek, int seek_end, int n_max, uint32_t ep) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int tid = threadIdx.x, lane = tid & 31, ncta = gridDim.x;
    poll_vec<false>(P.stats, ncta * 8, ep, sm.xs);
    consumer_sync();
    if (tid < 32) {
        MaxIdx mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};
        for (int c = lane; c < ncta; c += 32) {
            const float *r = sm.xs + c * 8;
            mt = better(mt, MaxIdx{r[0], __float_as_int(r[1])}); ms = better(ms, MaxIdx{r[2], __float_as_int(r[3])});
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            MaxIdx a{__shfl_xor_sync(0xffffffffu, mt.v, o), __shfl_xor_sync(0xffffffffu, mt.i, o)}; mt = better(mt, a);
            MaxIdx b{__shfl_xor_sync(0xffffffffu, ms.v, o), __shfl_xor_sync(0xffffffffu, ms.i, o)}; ms = better(ms, b);
        }
        const float m_all = fmaxf(mt.v, ms.v);
        float sa = 0.f, sb = 0.f;
        for (int c = lane; c < ncta; c += 32) {
            const float *r = sm.xs + c * 8;
            const float cm = fmaxf(r[0], r[2]), cs = r[2];
            if (cm > -INFINITY) sa += r[4] * expf(cm - m_all);
            if (cs > -INFINITY) sb += r[5] * expf(cs - ms.v);
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

// per-CTA slice of every segment kind (computed once: the 64-bit divisions stay off the critical path)
__device__ void build_segtab() {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    const int cta = blockIdx.x, ncta = gridDim.x, d = P.d;
    for (int kind = threadIdx.x; kind < SEG_COUNT; kind += blockDim.x) {
        SegTab s;
        if (kind == SEG_XK || kind == SEG_XV) {
            s.row_bytes = 128;
            if (cta >= P.H * P.xsplit) { s.row0 = 0; s.rows = 0; }
            else {
                const int sp = cta % P.xsplit, per = (P.T + P.xsplit - 1) / P.xsplit;
                s.row0 = sp * per; s.rows = max(0, min(P.T, s.row0 + per) - s.row0);
            }
            s.prows = s.rows;
            s.rows_per_chunk = max(1, kChunkBytes / s.row_bytes);
        } else {
            const int N = kind == SEG_QKV ? 3 * d : kind == SEG_FC1 ? 4 * d : kind == SEG_LM ? P.n_vocab : d;
            const int kq = kind == SEG_FC2 ? 4 : 1;
            const int unit = kind == SEG_FC1 ? 2 : 1;     // FC1 publishes its rows in pairs
            s.row0 = unit * (int)((long)cta * (N / unit) / ncta); s.rows = unit * (int)((long)(cta + 1) * (N / unit) / ncta) - s.row0;
            s.row_bytes = d * 2; s.prows = s.rows * kq;      // rows of d halfs: FC2's 4d-long rows count as 4
            s.rows_per_chunk = kChunkRows;
        }
        s.n_chunks = (s.prows + s.rows_per_chunk - 1) / s.rows_per_chunk;
        sm.seg[kind] = s;
    }
}
__device__ __forceinline__ const uint8_t *seg_base(const MegaParams &P, const SegTab &s, int kind, int il) {
    if (kind == SEG_XK || kind == SEG_XV) {
        const int h = blockIdx.x / P.xsplit;
        return reinterpret_cast<const uint8_t *>((kind == SEG_XK ? P.cross_k : P.cross_v) + (size_t)il * 2 * P.T * P.d + ((size_t)h * P.T + s.row0) * 64);
    }
    const int widx = kind == SEG_QKV ? 0 : kind == SEG_O ? 1 : kind == SEG_CQ ? 2 : kind == SEG_CO ? 3 : kind == SEG_FC1 ? 4 : 5;
    const __half *w = kind == SEG_LM ? P.tok_emb : P.layer[il].w[widx];
    return reinterpret_cast<const uint8_t *>(w) + (size_t)s.row0 * (kind == SEG_FC2 ? 4 : 1) * s.row_bytes;
}

struct ConsArgs { int L, pos0, n_prompt, do_sample, seek, seek_end, n_max, keep_logits, all_logits, steps_left; };

// the consumer warps' decode loop, specialised on the model width (KS = d / 128)
template <int KS>
__device__ __noinline__ uint32_t consumer_main(const ConsArgs a) {
    MegaSmem &sm = SM;
    const MegaParams &P = sm.P;
    DecCtl *ctl = P.ctl;
    const int tid = threadIdx.x;
    const int L = a.L, pos0 = a.pos0, n_prompt = a.n_prompt, do_sample = a.do_sample, seek = a.seek, seek_end = a.seek_end, n_max = a.n_max,
              keep_logits = a.keep_logits, all_logits = a.all_logits, steps_left = a.steps_left;
    uint32_t cons = 0, ph = 0;
    const uint32_t ep_stride = (uint32_t)L + 1;

    for (int t = 0; t < steps_left; t++) {
        if (sm.st.done) break;
        const int jrel = sm.st.pos - pos0;
        const bool need_logits = all_logits || jrel >= n_prompt - 1;
        const uint32_t ep0 = 1u + (uint32_t)t * ep_stride;      // epoch of layer il: ep0 + il; LM head input: ep0 + L

#pragma unroll 1
        for (int il = 0; il < L; il++) {
            const uint32_t ep = ep0 + (uint32_t)il;
            if (SS_MEGA_TRACE && tid == 0) { sm.trace_on = t == kTraceStep; sm.trace_layer = il; }
            long long tp = kProf ? clock64() : 0;
#define SS_PROF_PHASE(k) if (kProf) { const long long tn = clock64(); if (tid == 0) sm.prof[4 + (k)] += tn - tp; tp = tn; }
            cons = gemv_phase<KS, SEG_QKV>(cons, il, ep, ep, ph++);       SS_PROF_PHASE(0)
            self_attn(il, ep);                                  SS_PROF_PHASE(1)
            cons = gemv_phase<KS, SEG_O>(cons, il, ep, ep, ph++);         SS_PROF_PHASE(2)
            cons = gemv_phase<KS, SEG_CQ>(cons, il, ep, ep, ph++);        SS_PROF_PHASE(3)
            cons = cross_attn(cons, ep);                        SS_PROF_PHASE(4)
            cons = gemv_phase<KS, SEG_CO>(cons, il, ep, ep, ph++);        SS_PROF_PHASE(5)
            cons = gemv_phase<KS, SEG_FC1>(cons, il, ep, ep, ph++);       SS_PROF_PHASE(6)
            cons = gemv_phase<KS, SEG_FC2>(cons, il, ep, ep + 1, ph++);   SS_PROF_PHASE(7)
        }
        long long tp = kProf ? clock64() : 0;

        // ---------------- final LN + LM head + per-CTA softmax statistics ----------------
        const uint32_t epL = ep0 + (uint32_t)L;
        const bool sampling = do_sample && jrel >= n_prompt - 1;
        if (need_logits) {
            cons = gemv_phase<KS, SEG_LM>(cons, 0, epL, epL, ph++);
            const bool keep = keep_logits && sm.st.n_kept < P.keep_cap;
            lm_epilogue(keep, sampling, epL);
            consumer_sync();
            if (keep && tid == 0) sm.st.n_kept = sm.st.n_kept + 1;
            consumer_sync();
        }
        if (jrel < n_prompt - 1) {   // prompt token: feed the next one
            consumer_sync();
            if (tid == 0) { sm.st.token = ctl->prompt[jrel + 1]; sm.st.pos = sm.st.pos + 1; if (kProf) sm.prof[12] += clock64() - tp; }
            consumer_sync();
            continue;
        }
        if (!do_sample) { consumer_sync(); if (tid == 0) { sm.st.done = 1; if (kProf) sm.prof[12] += clock64() - tp; } consumer_sync(); break; }
        sample_and_update(seek, seek_end, n_max, epL);
        SS_PROF_PHASE(8)
#undef SS_PROF_PHASE
    }

    return cons;
}

}  // namespace

// ================================================================================================
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaParams *__restrict__ Pp, int max_steps) {
    MegaSmem &sm = SM;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(Pp);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.P);
        for (int i = tid; i < (int)(sizeof(MegaParams) / 4); i += kMegaThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const MegaParams &P = sm.P;
    const int L = P.L;
    build_segtab();

    DecCtl *ctl = P.ctl;
    // snapshot of the control block (identical in every CTA)
    const int pos_start = ctl->pos, pos0 = ctl->pos0, n_prompt = ctl->n_prompt, do_sample = ctl->sample;
    const int seek = ctl->seek, seek_end = ctl->seek_end, n_max = ctl->n_max, keep_logits = ctl->keep_logits, all_logits = ctl->all_logits;
    const int already_done = ctl->done;
    const int steps_left = max_steps;

    if (tid == 0) {
        for (int s = 0; s < kSlots; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kConsumerWarps); }
        sm.stop_req = 0; sm.prod_done = 0; sm.issued = 0; sm.trace_on = 0; sm.trace_layer = 0;
        for (int i = 0; i < kProfN; i++) sm.prof[i] = 0;
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
        // Lane 0 tracks the ring and issues one bulk copy per chunk (16 weight rows, or a K / V slice of the cross-attention cache).
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
                const SegTab seg = sm.seg[kind];
                const uint8_t *base = seg_base(P, seg, kind, layer < L ? layer : 0);
                for (int ch = 0; ch < seg.n_chunks; ch++) {
                    const int slot = issued % kSlots;
                    const uint32_t par = ((issued / kSlots) & 1) ^ 1;
                    const int rbase = ch * seg.rows_per_chunk;
                    const int nrows = min(seg.rows_per_chunk, seg.prows - rbase);
                    if (lane == 0) {
                        while (!mbar_try_wait(&sm.empty[slot], par)) { if (sm.stop_req) { stopped = true; break; } __nanosleep(128); }   // do not out-prioritise the consumer warps of this scheduler
                        if (sm.stop_req) stopped = true;
                        if (SS_MAX_INFLIGHT > 0 && !stopped && issued >= (uint32_t)SS_MAX_INFLIGHT) {      // copy issued - SS_MAX_INFLIGHT must have landed
                            const uint32_t o = issued - (uint32_t)SS_MAX_INFLIGHT;
                            while (!mbar_try_wait(&sm.full[o % kSlots], (o / kSlots) & 1)) { if (sm.stop_req) { stopped = true; break; } }
                        }
                        if (!stopped) mbar_expect_tx(&sm.full[slot], SS_NO_STREAM ? 16u : (uint32_t)nrows * seg.row_bytes);
                    }
                    stopped = __shfl_sync(0xffffffffu, (int)stopped, 0) != 0;
                    __syncwarp();
                    if (stopped) break;
                    if (lane == 0) bulk_g2s(sm.ring[slot], base + (size_t)rbase * seg.row_bytes, SS_NO_STREAM ? 16u : (uint32_t)nrows * seg.row_bytes, &sm.full[slot], policy);
                    issued++;
                }
            }
        }
        // hand the issue count to the consumers so that they can drain in-flight copies before exit
        if (lane == 0) {
            sm.issued = issued;
            __threadfence_block();
            sm.prod_done = 1;
        }
        return;
    }

    // ======================= consumers =======================
    const long long t_begin = clock64();
    ConsArgs ca{L, pos0, n_prompt, do_sample, seek, seek_end, n_max, keep_logits, all_logits, steps_left};
    uint32_t cons = 0;
    switch (P.d >> 7) {      // host side guarantees d % 128 == 0 and d <= 1280 (Whisper widths: 384 .. 1280; 256: test models)
        case 2: cons = consumer_main<2>(ca); break;
        case 3: cons = consumer_main<3>(ca); break;
        case 4: cons = consumer_main<4>(ca); break;
        case 6: cons = consumer_main<6>(ca); break;
        case 8: cons = consumer_main<8>(ca); break;
        default: cons = consumer_main<10>(ca); break;
    }

    // ---------------- shutdown: stop the producer, drain copies still in flight, publish the state ----------------
    if (tid == 0) {
        sm.stop_req = 1;
        while (!sm.prod_done) {}
        __threadfence_block();
        const uint32_t issued = sm.issued;
        for (uint32_t c = cons; c < issued; c++) mbar_wait(&sm.full[c % kSlots], (c / kSlots) & 1);
        const DecState st = sm.st;
        if (P.prof && SS_MEGA_TRACE) for (int i = 0; i < kProfN; i++) P.prof[(size_t)cta * kTraceN + 600 + i] = sm.prof[i];
        if (P.prof && !SS_MEGA_TRACE) {
            sm.prof[3] = clock64() - t_begin;
            for (int i = 0; i < kProfN; i++) P.prof[(size_t)cta * kProfN + i] = sm.prof[i];
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

void decode_mega_launch(const MegaParams *d_params, void *d_ll, size_t ll_bytes, int max_steps, int grid, cudaStream_t st) {
    CUDA_CHECK(cudaMemsetAsync(d_ll, 0, ll_bytes, st));     // epochs restart at 1 on every launch
    void *args[] = {(void *)&d_params, (void *)&max_steps};
    CUDA_CHECK(cudaLaunchCooperativeKernel((const void *)decode_mega_kernel, dim3(grid), dim3(kMegaThreads), args, decode_mega_smem_bytes(), st));
}

}  // namespace ss
