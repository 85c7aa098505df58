// xchg_bench3.cu - round 2: what bounds one all-to-all activation exchange of the persistent decode kernel, and what a
// 2-level exchange (thread-block clusters, DSMEM forward with st.async + mbarrier complete_tx) buys.  Bounded spins everywhere.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench3 tools/xchg_bench3.cu
// Variants (template V):
//   0  flag-in-data, every CTA polls the whole vector (the kernel's protocol today)
//   1  every CTA polls only 1/S of the vector (S = `share`), nothing forwarded: lower bound of any S-way shared poll
//   2  cluster of S CTAs: each polls 1/S and forwards its words into every member's shared memory with st.async (complete_tx on the
//      member's mbarrier); a member waits for its mbarrier only
//   3  group-local exchange: groups of `share` CTAs exchange `d` words among themselves only (head-local q/k/v, partial folds)
//   5  like 0 with `share` replicas of the vector (writers store every replica, CTA c polls replica c % share)
//   6  like 0, but a thread re-polls only its stale pairs and surplus threads idle
//   4  like 0 with the poll done by TMA: one thread bulk-copies the vector to shared memory, all threads check the flags there
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define SPIN_LIMIT 400000

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ll_store(u64 *p, float v, unsigned ep) {
    const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load2(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_b64(uint32_t remote_addr, u64 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Smem {
    alignas(16) float xs[2][5120];
    alignas(16) u64 stage[5120];
    alignas(8) uint64_t bar[2];
};

template <int V>
__global__ void __launch_bounds__(256, 1) xk(u64 *buf, int d, int iters, int share, int work, int *err, long long *out) {
    extern __shared__ __align__(128) uint8_t raw[];
    Smem &sm = *reinterpret_cast<Smem *>(raw);
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
    int wcta = cta, wn = ncta;                    // writer index / count inside the exchange domain
    u64 *dom = buf;                               // exchange domain base (3 rotating buffers of d words)
    if (V == 3) { wcta = cta % share; wn = share; dom = buf + (size_t)(cta / share) * 3 * 8192; }
    const int row0 = (int)((long)wcta * d / wn), rows = (int)((long)(wcta + 1) * d / wn) - row0;
    const int n2 = d / 2;
    const uint32_t crank = (V == 2) ? cluster_rank() : 0;
    if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (V == 2) cluster_sync();
    float accv = 0.f;
    bool bad = false;
    const long long t0 = clock64();
    for (int it = 1; it <= iters && !bad; it++) {
        u64 *b = dom + (size_t)(it % 3) * d;
        float *xs = sm.xs[it & 1];
        if (V == 5) {
            b = buf + (size_t)(it % 3) * share * 8192;
            for (int t = tid; t < rows * share; t += 256) ll_store(b + (size_t)(t / rows) * 8192 + row0 + (t % rows), accv + (t % rows), (unsigned)it);
            b += (size_t)(cta % share) * 8192;
        } else
        for (int t = tid; t < rows; t += 256) ll_store(b + row0 + t, accv + t, (unsigned)it);
        if (V == 0 || V == 3 || V == 5) {
            for (int base = 0; base < n2; base += 256 * 3) {
                ulonglong2 v[3]; bool all; int spins = 0;
                do {
                    all = true;
#pragma unroll
                    for (int k = 0; k < 3; k++) { int i = base + tid + k * 256; if (i >= n2) i -= n2 * (i / n2); v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                    for (int k = 0; k < 3; k++) if ((unsigned)(v[k].x >> 32) != (unsigned)it || (unsigned)(v[k].y >> 32) != (unsigned)it) all = false;
                    if (++spins > SPIN_LIMIT) { *err = 1; bad = true; all = true; }
                } while (!all);
#pragma unroll
                for (int k = 0; k < 3; k++) { const int i = base + tid + k * 256; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
            }
        } else if (V == 6) {
            for (int base = 0; base < n2; base += 256 * 3) {
                ulonglong2 v[3]; bool ok[3] = {false, false, false}; bool all; int spins = 0;
                do {
                    all = true;
#pragma unroll
                    for (int k = 0; k < 3; k++) { const int i = base + tid + k * 256; if (i < n2 && !ok[k]) v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                    for (int k = 0; k < 3; k++) { const int i = base + tid + k * 256; if (i < n2 && !ok[k]) { ok[k] = (unsigned)(v[k].x >> 32) == (unsigned)it && (unsigned)(v[k].y >> 32) == (unsigned)it; if (!ok[k]) all = false; } }
                    if (++spins > SPIN_LIMIT) { *err = 1; bad = true; all = true; }
                } while (!all);
#pragma unroll
                for (int k = 0; k < 3; k++) { const int i = base + tid + k * 256; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
            }
        } else if (V == 1 || V == 2) {
            // my share of the pairs: [p0, p1)
            const int S = V == 2 ? (int)cluster_size() : share, me = V == 2 ? (int)crank : cta % share;
            const int p0 = (int)((long)me * n2 / S), p1 = (int)((long)(me + 1) * n2 / S);
            const uint32_t bar_l = smem_u32(&sm.bar[it & 1]);
            if (V == 2 && tid == 0) mbar_expect_tx(&sm.bar[it & 1], (uint32_t)n2 * 8u);
            for (int base = p0; base < p1; base += 256 * 2) {
                ulonglong2 v[2]; bool ok[2] = {false, false}; bool all; int spins = 0;
                do {
                    all = true;
#pragma unroll
                    for (int k = 0; k < 2; k++) { const int i = base + tid + k * 256; if (i < p1 && !ok[k]) v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                    for (int k = 0; k < 2; k++) { const int i = base + tid + k * 256; if (i < p1 && !ok[k]) { ok[k] = (unsigned)(v[k].x >> 32) == (unsigned)it && (unsigned)(v[k].y >> 32) == (unsigned)it; if (!ok[k]) all = false; } }
                    if (++spins > SPIN_LIMIT) { *err = 1; bad = true; all = true; }
                } while (!all);
#pragma unroll
                for (int k = 0; k < 2; k++) {
                    const int i = base + tid + k * 256;
                    if (i < p1) {
                        if (V == 1) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); }
                        else {
                            const u64 pay = ((u64)(unsigned)v[k].y << 32) | (u64)(unsigned)v[k].x;      // two f32 payloads
                            const uint32_t dst_l = smem_u32(&xs[2 * i]);
                            for (int r = 0; r < S; r++) st_async_b64(mapa(dst_l, (uint32_t)r), pay, mapa(bar_l, (uint32_t)r));
                        }
                    }
                }
            }
            if (V == 2) {
                int spins = 0;
                while (!mbar_try_wait_cluster(&sm.bar[it & 1], ((unsigned)(it - 1) >> 1) & 1)) { if (++spins > SPIN_LIMIT) { *err = 2; bad = true; break; } }
            }
        } else if (V == 4) {
            int spins = 0;
            uint32_t tma_ph = 0;
            for (;;) {
                if (tid == 0) { mbar_expect_tx(&sm.bar[0], (uint32_t)d * 8u); bulk_g2s(sm.stage, b, (uint32_t)d * 8u, &sm.bar[0]); }
                while (!mbar_try_wait(&sm.bar[0], tma_ph)) { if (++spins > SPIN_LIMIT) { *err = 3; bad = true; break; } }
                tma_ph ^= 1;
                bool all = true;
                for (int i = tid; i < d; i += 256) if ((unsigned)(sm.stage[i] >> 32) != (unsigned)it) all = false;
                const int ok = __syncthreads_and(all ? 1 : 0);
                if (ok || bad) break;
                if (++spins > SPIN_LIMIT) { *err = 3; bad = true; break; }
            }
            // an odd number of TMA rounds leaves the barrier's phase flipped: keep it in step with tma_ph across iterations
            if (tma_ph) { if (tid == 0) { mbar_expect_tx(&sm.bar[0], 16u); bulk_g2s(sm.stage, b, 16u, &sm.bar[0]); } while (!mbar_try_wait(&sm.bar[0], 1)) { if (++spins > SPIN_LIMIT) { bad = true; break; } } }
            for (int i = tid; i < d; i += 256) xs[i] = __uint_as_float((unsigned)sm.stage[i]);
        }
        bad = __syncthreads_or(bad ? 1 : 0) != 0;
        accv = xs[(tid * 7 + it) % (V == 1 ? 64 : d)] * 0.5f;
        if (work > 0) { const long long w0 = clock64(); while (clock64() - w0 < work) {} }
        __syncthreads();
        if ((it & 63) == 0 && *(volatile int *)err) bad = true;
    }
    if (tid == 0) out[cta] = clock64() - t0;
    if (accv == 12345.678f) out[0] = 0;
    if (V == 2) cluster_sync();      // nobody leaves while a peer may still st.async into it
}

static u64 *g_buf; static int *g_err; static long long *g_out;

template <int V>
static void run(const char *name, int grid, int d, int share, int work, int cluster) {
    const int iters = 2000;
    cudaMemset(g_buf, 0, (size_t)80 * 3 * 8192 * 8); cudaMemset(g_err, 0, 4);
    cudaFuncSetAttribute(xk<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (cluster > 8) cudaFuncSetAttribute(xk<V>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sizeof(Smem); cfg.stream = 0;
    cudaLaunchAttribute at[2]; int na = 0;
    if (cluster > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cluster; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; na++; }
    at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++;
    cfg.attrs = at; cfg.numAttrs = na;
    if (cluster > 1) {
        int ncl = 0; cudaOccupancyMaxActiveClusters(&ncl, xk<V>, &cfg);
        if (ncl * cluster < grid) { printf("%-28s grid=%3d cluster=%d: only %d clusters can be co-resident - skipped\n", name, grid, cluster, ncl); return; }
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, xk<V>, g_buf, d, iters, share, work, g_err, g_out);
    if (e != cudaSuccess) {      // cooperative + cluster may be refused: all CTAs are co-resident anyway (grid <= SMs, 1 CTA / SM, idle GPU)
        printf("   (cooperative launch refused: %s; retrying without)\n", cudaGetErrorString(e)); cudaGetLastError();
        cfg.numAttrs = na - 1;
        e = cudaLaunchKernelEx(&cfg, xk<V>, g_buf, d, iters, share, work, g_err, g_out);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    int he = 0; cudaMemcpy(&he, g_err, 4, cudaMemcpyDeviceToHost);
    static long long h[256];
    cudaMemcpy(h, g_out, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-28s grid=%3d d=%4d share=%d work=%4d cluster=%d: %6.0f cycles per exchange%s %s\n", name, grid, d, share, work, cluster,
           (double)mx / iters - work, he ? "  [SPIN LIMIT HIT]" : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
    fflush(stdout);
    if (e != cudaSuccess) { cudaGetLastError(); }
}

int main() {
    cudaMalloc(&g_buf, (size_t)80 * 3 * 8192 * 8); cudaMalloc(&g_err, 4); cudaMalloc(&g_out, 256 * 8);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs: %d\n", sms);
    for (int rep = 0; rep < 2; rep++) {
        for (int d : {640, 1280, 2560}) run<0>("all poll all", sms, d, 1, 0, 1);
        for (int d : {640, 1280, 2560}) run<6>("all poll all, stale-only", sms, d, 1, 0, 1);
        run<6>("all poll all, stale-only", 132, 1280, 1, 0, 1);
        run<6>("all poll all, stale-only", 2, 1280, 1, 0, 1);
        for (int s : {2, 4, 8}) for (int d : {640, 1280, 2560}) run<1>("poll 1/S only (bound)", sms, d, s, 0, 1);
        for (int d : {640, 1280, 2560}) run<2>("cluster forward st.async", sms, d, 2, 0, 2);
        for (int d : {640, 1280, 2560}) run<2>("cluster forward st.async", 132, d, 4, 0, 4);
        for (int d : {640, 1280, 2560}) run<2>("cluster forward st.async", 144, d, 8, 0, 8);
        run<2>("cluster forward st.async", 128, 1280, 16, 0, 16);
        run<2>("cluster fwd + work", sms, 1280, 2, 1000, 2);
        run<2>("cluster fwd + work", 132, 1280, 4, 1000, 4);
    }
    return 0;
}
