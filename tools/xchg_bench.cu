// xchg_bench.cu - cost of ONE all-to-all activation exchange between the CTAs of the persistent decode kernel, in
// isolation (no compute): every CTA publishes its ~d/grid values, every CTA needs all d.  Variants of the protocol are
// timed back to back on the same GPU.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xchg_bench xchg_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
typedef unsigned long long u64;

__device__ __forceinline__ void ll_store(u64 *p, float v, unsigned ep) {
    const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_store_flavour(u64 *p, float v, unsigned ep, int fl) {
    const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
    if (fl == 1) atomicExch(p, w);
    else if (fl == 2) asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    else if (fl == 3) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    else if (fl == 4) asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    else if (fl == 5) asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load2(const u64 *p) {
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ll_load1(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// mode 0: flag-in-data, every thread polls its own pairs (the kernel's protocol)
// mode 1: flag-in-data, only warp 0 polls (20 pairs per lane), the others wait at the CTA barrier
// mode 2: data + per-CTA counter: producers st data, fence, red.add; one thread polls the counter, then everybody loads
// mode 3: flag-in-data, threads poll one canary pair each first (64 canaries), then the full vector
// mode 4: like 0 but `work` cycles of fake compute between poll and publish (shows skew sensitivity)
__global__ void __launch_bounds__(256, 1) xchg_kernel(u64 *buf, unsigned *counter, int d, int iters, int mode, int work, long long *out) {
    const int cta = blockIdx.x, ncta = gridDim.x, tid = threadIdx.x;
    const int row0 = (int)((long)cta * d / ncta), rows = (int)((long)(cta + 1) * d / ncta) - row0;
    const int n2 = d / 2;
    __shared__ float xs[8192];
    float accv = 0.f;
    cooperative_groups::this_grid().sync();
    const long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        u64 *b = buf + (size_t)(it % 3) * d;
        // publish
        const int R = mode >= 30 && mode < 40 ? (1 << (mode - 30)) : 1;
        if (mode >= 30 && mode < 40) {      // R replicas of the vector; CTA c reads replica c % R
            u64 *bb = buf + (size_t)(it % 3) * d * R;
            for (int t = tid; t < rows * R; t += 256) ll_store(bb + (size_t)(t / rows) * d + row0 + (t % rows), accv + (t % rows), (unsigned)it);
            b = bb + (size_t)(cta % R) * d;
        } else if (mode == 2) {
            if (tid < rows) b[row0 + tid] = (u64)__float_as_uint(accv + tid);
            __syncthreads();
            if (tid == 0) { __threadfence(); atomicAdd(counter + (it % 3) * 32, 1u); }
        } else if (mode >= 10 && mode < 20) { if (tid < rows) ll_store_flavour(b + row0 + tid, accv + tid, (unsigned)it, mode - 10); }
        else if (mode == 21) { if (tid < rows) ll_store(b + row0 + tid, accv + tid, (unsigned)it); }
        else for (int t = tid; t < rows; t += 256) ll_store(b + row0 + t, accv + t, (unsigned)it);
        // gather
        if (mode == 0 || mode == 4 || (mode >= 10 && mode < 20) || (mode >= 30 && mode < 40)) {
            ulonglong2 v[3]; bool all;
            do {
                all = true;
#pragma unroll
                for (int k = 0; k < 3; k++) { const int i = min(tid + k * 256, n2 - 1); v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                for (int k = 0; k < 3; k++) if ((unsigned)(v[k].x >> 32) != (unsigned)it || (unsigned)(v[k].y >> 32) != (unsigned)it) all = false;
            } while (!all);
#pragma unroll
            for (int k = 0; k < 3; k++) { const int i = tid + k * 256; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
        } else if (mode == 20) {
            ulonglong2 v; const int i = min(tid, n2 - 1);
            do { v = ll_load2(b + 2 * i); } while ((unsigned)(v.x >> 32) != (unsigned)it || (unsigned)(v.y >> 32) != (unsigned)it);
            xs[2 * i] = __uint_as_float((unsigned)v.x);
        } else if (mode == 21) {      // nanosleep between polls
            ulonglong2 v[3]; bool all;
            do {
                all = true;
#pragma unroll
                for (int k = 0; k < 3; k++) { const int i = min(tid + k * 256, n2 - 1); v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                for (int k = 0; k < 3; k++) if ((unsigned)(v[k].x >> 32) != (unsigned)it || (unsigned)(v[k].y >> 32) != (unsigned)it) all = false;
                if (!all) __nanosleep(100);
            } while (!all);
#pragma unroll
            for (int k = 0; k < 3; k++) { const int i = tid + k * 256; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
        } else if (mode == 1) {
            if (tid < 32) {
                for (int base = 0; base < n2; base += 32 * 10) {
                    ulonglong2 v[10]; bool all;
                    do {
                        all = true;
#pragma unroll
                        for (int k = 0; k < 10; k++) { const int i = min(base + tid + k * 32, n2 - 1); v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                        for (int k = 0; k < 10; k++) if ((unsigned)(v[k].x >> 32) != (unsigned)it || (unsigned)(v[k].y >> 32) != (unsigned)it) all = false;
                    } while (!all);
#pragma unroll
                    for (int k = 0; k < 10; k++) { const int i = base + tid + k * 32; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
                }
            }
        } else if (mode == 2) {
            if (tid == 0) { const unsigned want = (unsigned)ncta * ((it + 2) / 3); while (*(volatile unsigned *)(counter + (it % 3) * 32) < want) {} __threadfence(); }
            __syncthreads();
            for (int i = tid; i < n2; i += 256) { const ulonglong2 v = ll_load2(b + 2 * i); xs[2 * i] = __uint_as_float((unsigned)v.x); xs[2 * i + 1] = __uint_as_float((unsigned)v.y); }
        } else if (mode == 3) {
            if (tid < 64) { const int i = min(tid * (n2 / 64), n2 - 1); ulonglong2 c; do { c = ll_load2(b + 2 * i); } while ((unsigned)(c.x >> 32) != (unsigned)it || (unsigned)(c.y >> 32) != (unsigned)it); }
            __syncthreads();
            ulonglong2 v[3]; bool all;
            do {
                all = true;
#pragma unroll
                for (int k = 0; k < 3; k++) { const int i = min(tid + k * 256, n2 - 1); v[k] = ll_load2(b + 2 * i); }
#pragma unroll
                for (int k = 0; k < 3; k++) if ((unsigned)(v[k].x >> 32) != (unsigned)it || (unsigned)(v[k].y >> 32) != (unsigned)it) all = false;
            } while (!all);
#pragma unroll
            for (int k = 0; k < 3; k++) { const int i = tid + k * 256; if (i < n2) { xs[2 * i] = __uint_as_float((unsigned)v[k].x); xs[2 * i + 1] = __uint_as_float((unsigned)v[k].y); } }
        }
        __syncthreads();
        accv = xs[(tid * 7 + it) % d] * 0.5f;
        if (mode == 4 && work > 0) { const long long w0 = clock64(); while (clock64() - w0 < work) {} }
        __syncthreads();
    }
    if (tid == 0) out[cta] = clock64() - t0;
    if (accv == 12345.678f) out[0] = 0;
}

__global__ void pingpong_kernel(u64 *buf, int iters, long long *out) {
    const int cta = blockIdx.x;
    if (cta > 1 || threadIdx.x != 0) return;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        if (cta == 0) { ll_store(buf, 1.f, (unsigned)it); while ((unsigned)(ll_load1(buf + 16) >> 32) != (unsigned)it) {} }
        else { while ((unsigned)(ll_load1(buf) >> 32) != (unsigned)it) {} ll_store(buf + 16, 2.f, (unsigned)it); }
    }
    out[cta] = clock64() - t0;
}

int main(int argc, char **argv) {
    const int d = argc > 1 ? atoi(argv[1]) : 1280, iters = argc > 2 ? atoi(argv[2]) : 2000;
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    u64 *buf; unsigned *counter; long long *out;
    cudaMalloc(&buf, (size_t)3 * 8192 * 8 * 16); cudaMalloc(&counter, 3 * 32 * 4); cudaMalloc(&out, sms * 8);
    long long *h = (long long *)malloc(sms * 8);
    {
        cudaMemset(buf, 0, 4096);
        for (int pair = 1; pair < sms; pair += 37) {
            // CTA placement is up to the hardware: launch `sms` CTAs, only 0 and 1 play; different grids give different SM pairs
            pingpong_kernel<<<2, 32>>>(buf, 2000, out); cudaDeviceSynchronize();
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("ping-pong: %.0f cycles per round trip (one flagged store + one polled load each way)\n", (double)h[0] / 2000);
            cudaMemset(buf, 0, 4096);
        }
    }
    const int modes[] = {0, 1, 2, 3, 4, 21, 30, 31, 32, 33, 34};      // (mode 20 polls only a prefix of the vector: late CTAs can run ahead and deadlock - never in a default run)
    // the scaling with the number of CTAs is measured by tools/xchg_bench2.cu (bounded spins); here: the full grid only
    for (int rep = 0; rep < 2; rep++) {
        for (int mi = 0; mi < (int)(sizeof(modes) / sizeof(int)); mi++) {
            const int mode = modes[mi];
            for (int work = 0; work <= (mode == 4 ? 2000 : 0); work += 1000) {
                if (mode == 4 && work == 0) continue;
                cudaMemset(buf, 0, (size_t)3 * 8192 * 8 * 16); cudaMemset(counter, 0, 3 * 32 * 4);
                int dd = d, it = iters, md = mode, wk = work;
                void *args[] = {&buf, &counter, &dd, &it, &md, &wk, &out};
                cudaError_t e = cudaLaunchCooperativeKernel((const void *)xchg_kernel, dim3(sms), dim3(256), args, 0, 0);
                if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return 1; }
                e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sync: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, out, sms * 8, cudaMemcpyDeviceToHost);
                long long mx = 0; for (int i = 0; i < sms; i++) mx = h[i] > mx ? h[i] : mx;
                printf("d=%d mode=%d work=%d: %.0f cycles per exchange (%d CTAs)\n", d, mode, work, (double)mx / iters - work, sms);
            }
        }
    }
    return 0;
}
