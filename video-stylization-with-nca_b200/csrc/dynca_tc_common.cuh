// Pieces shared by the tensor-core DyNCA kernels (dynca_bf16.cu: 4x32 tiles + cp.async staging, any shape;
// dynca_tc2.cu: 8x16 tiles + TMA staging): tcgen05 / mbarrier PTX wrappers, UMMA descriptors, the cond chunk of A1
// and the geometry of the bf16 operand images.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "dynca_tile.cuh"
#include "nca_internal.h"

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x4000u)      // suspend-time hint: sleep instead of spinning on issue slots
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T ; one elected thread issues
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30),
// SBO [32,46), version=1 at bit 46, layout_type [61,64) = 0)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor for kind::f16: D=f32, A=B=bf16, both K-major (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_mn(int M, int N) {   // both operands MN-major
    return umma_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}


// the cond chunk of A1 for one cell: [cond_0..cond_{cc-1} (bf16 hi), 1, 1, cond_0..cond_{cc-1} (lo residuals), 0..];
// the lo slots exist when 2*cc + 2 <= 8 and make the cond inputs ~fp32-accurate (the matching W1 columns are
// simply repeated); the two constant-1 slots carry b1 as bf16 hi + lo
__device__ __forceinline__ bool dynca_cond_split(int cc) { return 2 * cc + 2 <= 8; }
__device__ __forceinline__ uint4 dynca_cond_chunk(const DyncaGeom& g, const float* __restrict__ cond, int b, int gy, int gx, bool inimg) {
    float cv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (inimg) {
        const bool split = dynca_cond_split(g.cc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int src = i < g.cc ? i : ((split && i >= g.cc + 2 && i < 2 * g.cc + 2) ? i - g.cc - 2 : -1);
            if (src >= 0) {
                float raw;
                if (g.cond_kind == NCA_COND_CPE) raw = (src == 0) ? dynca_cpe(gy, g.H, g.cpe_oh) : dynca_cpe(gx, g.W, g.cpe_ow);
                else raw = __ldg(cond + ((size_t)(b * g.cc + src) * g.H + gy) * g.W + gx);
                const float hi = __bfloat162float(__float2bfloat16_rn(raw));
                cv[i] = i < g.cc ? hi : raw - hi;
            } else if (i == g.cc || i == g.cc + 1) cv[i] = 1.0f;
        }
    }
    uint4 v;
    v.x = pack_bf16(cv[0], cv[1]); v.y = pack_bf16(cv[2], cv[3]); v.z = pack_bf16(cv[4], cv[5]); v.w = pack_bf16(cv[6], cv[7]);
    return v;
}
// split precision (NCA_PREC_F16X3): v = hi + lo with hi = fp16(v), lo = fp16(v - hi), i.e. ~22 significant bits per operand
// (bf16 pairs give 16, which measurably misses the 1e-5 bar); a product A.B is then evaluated as Ah.Bh + Al.Bh + Ah.Bl with fp32
// accumulation (the Al.Bl term, 2^-22 relative, is dropped).  fp16 has 5 exponent bits, so the operands are kept in range by
// exact power-of-two scales: weight images x 2^8 (NCA_X3_WSCALE; |w| < 255), gradient operands x a per-launch scale that brings
// max|g| to [32, 64); perception / hidden values are used as they are (|v| < 65504: beyond it the step yields NaN, loudly).
#define NCA_X3_WSCALE 256.0f
#define NCA_X3_WINV (1.0f / 256.0f)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// instruction descriptors for kind::f16 with fp16 operands (formats 0), K-major / MN-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(int M, int N) { return umma_idesc_f16(M, N) | (1u << 15) | (1u << 16); }
// gradient-operand scale of a launch from max|g| (device scalar written by dynca_absmax_kernel): 2^k with max * 2^k in [32, 64)
__device__ __forceinline__ float nca_x3_gscale(float gmax) {
    if (!(gmax > 0.0f) || !(gmax < 3.0e38f)) return 1.0f;
    int e;
    frexpf(gmax, &e);                       // gmax = f * 2^e, f in [0.5, 1)
    e = 6 - e;
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    return exp2f((float)e);
}
// the cond chunk as hi / lo images: same slots as dynca_cond_chunk; without room for the lo slots (cc > 3) the raw value goes
// into the hi slot and its residual into the lo image
__device__ __forceinline__ void dynca_cond_chunk_x3(const DyncaGeom& g, const float* __restrict__ cond, int b, int gy, int gx, bool inimg,
                                                    uint4& hi, uint4& lo) {
    float cv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (inimg) {
        const bool split = dynca_cond_split(g.cc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int src = i < g.cc ? i : ((split && i >= g.cc + 2 && i < 2 * g.cc + 2) ? i - g.cc - 2 : -1);
            if (src >= 0) {
                float raw;
                if (g.cond_kind == NCA_COND_CPE) raw = (src == 0) ? dynca_cpe(gy, g.H, g.cpe_oh) : dynca_cpe(gx, g.W, g.cpe_ow);
                else raw = __ldg(cond + ((size_t)(b * g.cc + src) * g.H + gy) * g.W + gx);
                const float h = __bfloat162float(__float2bfloat16_rn(raw));
                cv[i] = i < g.cc ? (split ? h : raw) : raw - h;
            } else if (i == g.cc || i == g.cc + 1) cv[i] = 1.0f;
        }
    }
    split_f16x2(cv[0], cv[1], hi.x, lo.x); split_f16x2(cv[2], cv[3], hi.y, lo.y);
    split_f16x2(cv[4], cv[5], hi.z, lo.z); split_f16x2(cv[6], cv[7], hi.w, lo.w);
}
// reference W1 column (or -1 = none, -2 = b1 hi, -3 = b1 lo) feeding slot s of the cond chunk
__host__ __device__ __forceinline__ int dynca_cond_slot_src(int cc, int s) {
    if (s < cc) return s;
    if (s == cc) return -2;
    if (s == cc + 1) return -3;
    if (2 * cc + 2 <= 8 && s < 2 * cc + 2) return s - cc - 2;
    return -1;
}

// ---- geometry of the bf16 operands ---------------------------------------------------------------
struct Bf16Geom {
    int npairs;      // ceil(C/2) perception chunks
    int K1;          // padded K of GEMM1 (multiple of 16)
    int N1;          // fc (multiple of 16)
    int tmem_cols;   // power of two >= N1 + 16
    uint32_t a1_bytes, b1_bytes, a2_bytes, b2_bytes;
};
// bytes of one tile's perception operand in the operand history
__host__ __device__ static inline uint32_t dynca_tc2_op_tile_bytes(const DyncaGeom& g) {
    const int npairs = (g.C + 1) / 2, K1 = ((npairs + 1) * 8 + 15) / 16 * 16;
    return (uint32_t)(K1 / 8) * 2048u;
}
static inline int dynca_bf16_geom(const DyncaGeom& g, Bf16Geom* b) {
    if (g.fc % 16 != 0 || g.fc < 16 || g.fc > 240) { nca_set_error("bf16 path needs fc %% 16 == 0 and 16 <= fc <= 240 (got %d)", g.fc); return NCA_ERR_UNSUPPORTED; }
    if (g.cc + 2 > 8) { nca_set_error("bf16 path supports at most 6 cond channels (got %d)", g.cc); return NCA_ERR_UNSUPPORTED; }
    b->npairs = (g.C + 1) / 2;
    b->K1 = ((b->npairs + 1) * 8 + 15) / 16 * 16;
    b->N1 = g.fc;
    int need = g.fc + 16, cols = 32;
    while (cols < need) cols *= 2;
    b->tmem_cols = cols;
    b->a1_bytes = (uint32_t)(b->K1 / 8) * 2048u;
    b->b1_bytes = (uint32_t)(b->K1 / 8) * (uint32_t)(g.fc / 8) * 128u;
    b->a2_bytes = (uint32_t)(g.fc / 8) * 2048u;
    b->b2_bytes = (uint32_t)(g.fc / 8) * 256u;
    return NCA_OK;
}
