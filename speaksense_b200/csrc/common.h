// common.h - shared host/device helpers for the speaksense_b200 CUDA engine.
#pragma once
#include <atomic>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace ss {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define SS_THROW(code, ...)                                        \
    do {                                                           \
        char _b[512];                                              \
        snprintf(_b, sizeof _b, __VA_ARGS__);                      \
        throw ::ss::Error((code), _b);                             \
    } while (0)

#define CUDA_CHECK(expr)                                                                        \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            SS_THROW(-4, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                     cudaGetErrorString(_e));                                                   \
    } while (0)

#define SS_STR2(x) #x
#define SS_STR(x) SS_STR2(x)

constexpr int kSampleRate = 16000;
constexpr int kNFft = 400;
constexpr int kHop = 160;
constexpr int kNBins = 201;
constexpr int kChunkSec = 30;

template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Per-device one-time setup (function attributes, constant tables are per device / context, not per process): `need(dev)` is true
// until `done(dev)` has been called for that device.  Two threads racing on the first use both run the (idempotent) setup.
struct PerDeviceOnce {
    std::atomic<unsigned long long> mask{0};      // devices 0..63
    bool need(int dev) const { return dev < 0 || dev >= 64 || !((mask.load(std::memory_order_acquire) >> dev) & 1ull); }
    void done(int dev) { if (dev >= 0 && dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release); }
};
static inline int current_device() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess) dev = -1; return dev; }

}  // namespace ss
