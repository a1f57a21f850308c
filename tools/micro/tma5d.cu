// micro test: 5-D TMA tile load with negative coordinates (zero fill) from [S][B][C][H][W] fp32
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int bw, int bh, int C, int x, int y, int b, int s) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = (uint64_t*)smem;
    float* dst = (float*)(smem + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * C * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(smem_u32(dst)), "l"((uint64_t)&tm), "r"(x), "r"(y), "r"(0), "r"(b), "r"(s), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh * C; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
    int S = 3, B = 2, C = 12, H = 12, W = 16, bw = argc > 1 ? atoi(argv[1]) : 20, bh = argc > 2 ? atoi(argv[2]) : 10;
    int l2 = argc > 3 ? atoi(argv[3]) : 2; int cx = argc > 4 ? atoi(argv[4]) : -1; int cy = argc > 5 ? atoi(argv[5]) : -1; int cb = argc > 6 ? atoi(argv[6]) : 1; int cs = argc > 7 ? atoi(argv[7]) : 2;
    size_t n = (size_t)S * B * C * H * W;
    float* h = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn fn = (EncodeFn)p;
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)S};
    cuuint64_t strides[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4, (cuuint64_t)B * C * H * W * 4};
    cuuint32_t box[5] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)C, 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult rc = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d (box %d x %d, l2 %d)\n", (int)rc, bw, bh, l2);
    k<<<1, 128, 128 + bw * bh * C * 4>>>(tm, o, bw, bh, C, cx, cy, cb, cs); printf("coords %d %d %d %d\n", cx, cy, cb, cs);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float* r = (float*)malloc(bw * bh * C * 4);
        cudaMemcpy(r, o, bw * bh * C * 4, cudaMemcpyDeviceToHost);
        // expect r[c][yy][xx] = h[s=2][b=1][c][yy-1][xx-1] or 0
        int bad = 0;
        for (int c = 0; c < C; ++c) for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) {
            int gy = yy + cy, gx = xx + cx;
            float want = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[((((size_t)cs * B + cb) * C + c) * H + gy) * W + gx] : 0.f;
            if (r[(c * bh + yy) * bw + xx] != want) ++bad;
        }
        printf("mismatches %d\n", bad);
    }
    return 0;
}
