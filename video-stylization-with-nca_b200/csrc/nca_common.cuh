// Shared device/host helpers for the NCA step kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nca_b200.h"

// ---- host-side error / launch accounting (nca_api.cu owns the storage) ---------------------
void nca_set_error(const char* fmt, ...);
void nca_count_launch(int n = 1);
// SM count / ordinal of the CURRENT device (cached per device: one process may drive several GPUs, from several threads)
int nca_sm_count();
int nca_device_ordinal();
#define NCA_MAX_DEVICES 64
#define NCA_CHECK_ARG(cond, ...)                     \
    do {                                             \
        if (!(cond)) {                               \
            nca_set_error(__VA_ARGS__);              \
            return NCA_ERR_ARG;                      \
        }                                            \
    } while (0)
#define NCA_CUDA_OK(expr)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            nca_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return NCA_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)
#define NCA_LAUNCH_OK()                                                                       \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        nca_count_launch();                                                                   \
        if (e__ != cudaSuccess) {                                                             \
            nca_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return NCA_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

static inline size_t nca_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- padding index map (F.pad semantics, ExtraChannels/models/dynca.py:81) -------------------
// returns the image index read for padded coordinate r in [-n, 2n), or -1 for "zero" (constant pad)
__host__ __device__ __forceinline__ int nca_padmap(int r, int n, int mode) {
    if (r >= 0 && r < n) return r;
    if (mode == NCA_PAD_CONSTANT) return -1;
    if (mode == NCA_PAD_CIRCULAR) { r %= n; return r < 0 ? r + n : r; }
    if (mode == NCA_PAD_REPLICATE) return r < 0 ? 0 : n - 1;
    /* reflect */ r = r < 0 ? -r : 2 * (n - 1) - r;
    return r < 0 ? 0 : (r >= n ? n - 1 : r);
}

// ---- Philox4x32-10 fire mask (restated in oracle/philox.py for the tests) --------------------
// ctr = (quad, b, t, 'NCA1'), key = seed; pixel p = y*W+x uses word p&3 of quad p>>2.
#define NCA_PHILOX_STREAM 0x4E434131u
__device__ __forceinline__ uint4 nca_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t nca_philox_word(uint32_t p, uint32_t b, uint32_t t, uint32_t k0, uint32_t k1) {
    uint4 r = nca_philox4x32_10(p >> 2, b, t, NCA_PHILOX_STREAM, k0, k1);
    uint32_t l = p & 3u;
    return l == 0 ? r.x : (l == 1 ? r.y : (l == 2 ? r.z : r.w));
}
// thr is a 33-bit threshold (0 .. 2^32) carried in 64 bits
static inline uint64_t nca_fire_threshold(float rate, int enc) {
    double v = enc ? (double)rate * 4294967296.0 : (1.0 - (double)rate) * 4294967296.0;
    double c = v < 0 ? 0 : (v > 4294967296.0 ? 4294967296.0 : v);
    uint64_t u = (uint64_t)c;
    if ((double)u < c) ++u;   // ceil
    return u;
}
__device__ __forceinline__ float nca_fire(uint32_t r, uint64_t thr, int enc) {
    return enc ? ((uint64_t)r < thr ? 1.f : 0.f) : ((uint64_t)r >= thr ? 1.f : 0.f);
}
