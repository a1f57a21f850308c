#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int bw, int bh, int x, int y) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = (uint64_t*)smem;
    float* dst = (float*)(smem + 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(dst)), "l"((uint64_t)&tm), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
    int H = 64, W = 64, bw = 16, bh = 8;
    int byver = argc > 1 ? atoi(argv[1]) : 0;
    size_t n = (size_t)H * W;
    float* h = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 1 << 20);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ee;
    if (byver) ee = cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q);
    else ee = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry: %s q=%d p=%p\n", cudaGetErrorString(ee), (int)q, p);
    EncodeFn fn = (EncodeFn)p;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc %d; map words:", (int)rc);
    for (int i = 0; i < 8; ++i) printf(" %016llx", (unsigned long long)((uint64_t*)&tm)[i]);
    printf("\n");
    k<<<1, 128, 128 + bw * bh * 4>>>(tm, o, bw, bh, 4, 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float r[128];
        cudaMemcpy(r, o, sizeof(r), cudaMemcpyDeviceToHost);
        printf("r[0]=%g (want %g) r[17]=%g (want %g)\n", r[0], h[4 * W + 4], r[17], h[5 * W + 5]);
    }
    return 0;
}
