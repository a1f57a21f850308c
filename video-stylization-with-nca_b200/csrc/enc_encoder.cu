// ImageEncoder of the encoder-conditioned NCA (EncoderConditioning/encoder.py:5-64), forward and weight-gradient pass, fused:
//   x [B,3,H,W] -> [sobel_x, sobel_y, laplacian of mean_rgb(x) | 5x5 gaussian blur of each colour channel]   (6 feature planes,
//   every convolution zero padded)  ->  conv3x3 (6 -> 16, bias) + ReLU  ->  conv3x3 (16 -> 16, no bias)  =  goal encoding,
// written straight into the zero-padded goal tensor [B,C,H,W] the rollout consumes (channels C-16 .. C-1; the leading channels are
// zeroed here too: nca.py:198-203), so there is no torch.cat / F.pad / slice around the rollout.  The backward reads d(goal) out
// of the BPTT's g_goal [B,C,H,W] at the same channel offset and produces the gradients of the three trained tensors
// (embed.0.weight, embed.0.bias, embed.2.weight) - the input image is data, it gets no gradient (conditioned_trainer.py:118-137).
//
// Once per rollout, ~6.5 kflop per pixel: CUDA-core fp32, one 16x16 output tile per CTA iteration, everything between the input
// tile and the output tile stays in shared memory (the reference runs 9 cuDNN / elementwise launches with 5 intermediates in HBM).
// The forward keeps the two tensors its backward needs (features [B,6,H,W], hidden layer [B,16,H,W]: 88 B per pixel).
#include "nca_common.cuh"

int nca_check_device();   // nca_api.cu

namespace {

constexpr int EC = 3;          // colour channels
constexpr int EF = EC + 3;     // feature planes
constexpr int EE = 16;         // embedding width
constexpr int ET = 16;         // tile edge
constexpr int ENT = 256;       // threads

struct EncoderConsts {
    float gauss[25];           // encoder.py:59-63 (normalised 5x5, sigma 1), computed on the host exactly as the reference does
};

// shared-memory layout (floats)
constexpr int SX_E = ET + 8, SG_E = ET + 6, SF_E = ET + 4, SH_E = ET + 2;          // 24, 22, 20, 18
constexpr int SX_N = EC * SX_E * SX_E, SG_N = SG_E * SG_E, SF_N = EF * SF_E * SF_E;
constexpr int SH_P = SH_E * SH_E + 1;                                                // 325: plane stride coprime with the banks
constexpr int SH_N = EE * SH_P;
constexpr int W1_N = EF * 9 * EE, W2_N = EE * 9 * EE;                                // [i][tap][o]

__device__ __forceinline__ bool inimg(int y, int x, int H, int W) { return y >= 0 && y < H && x >= 0 && x < W; }

// ------------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENT) encoder_fwd_kernel(int B, int H, int W, int C_out, int c0, const float* __restrict__ x,
                                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                                          const float* __restrict__ w2, float* __restrict__ feats,
                                                          float* __restrict__ h1g, float* __restrict__ out, EncoderConsts k,
                                                          int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float sm[];
    float* sW1 = sm;                  // [i][tap][o]
    float* sB1 = sW1 + W1_N;          // [o]
    float* sW2 = sB1 + EE;            // [i][tap][o]
    float* sX = sW2 + W2_N;           // [c][24][24]
    float* sG = sX + SX_N;            // [22][22]
    float* sF = sG + SG_N;            // [f][20][20]
    float* sH = sF + SF_N;            // [i][325]
    const int tid = threadIdx.x;
    for (int i = tid; i < W1_N; i += ENT) { const int o = i % EE, tap = (i / EE) % 9, ci = i / (EE * 9); sW1[i] = w1[(o * EF + ci) * 9 + tap]; }
    for (int i = tid; i < W2_N; i += ENT) { const int o = i % EE, tap = (i / EE) % 9, ci = i / (EE * 9); sW2[i] = w2[(o * EE + ci) * 9 + tap]; }
    if (tid < EE) sB1[tid] = b1[tid];
    const size_t plane = (size_t)H * W;
    const int n_tiles = B * tiles_x * tiles_y;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
        const int y0 = ty * ET, x0 = tx * ET;
        __syncthreads();
        // ---- input tile, 4-pixel ring, zero outside the image (= the zero padding of every convolution that reads x) ----
        for (int i = tid; i < SX_N; i += ENT) {
            const int q = i % SX_E, r = (i / SX_E) % SX_E, c = i / (SX_E * SX_E);
            const int yy = y0 - 4 + r, xx = x0 - 4 + q;
            sX[i] = inimg(yy, xx, H, W) ? __ldg(x + ((size_t)b * EC + c) * plane + (size_t)yy * W + xx) : 0.0f;
        }
        __syncthreads();
        // ---- gray = mean over the colour channels (encoder.py:40), 3-pixel ring ----
        for (int i = tid; i < SG_N; i += ENT) {
            const int q = i % SG_E, r = i / SG_E;
            const int o = (r + 1) * SX_E + q + 1;
            sG[i] = (sX[o] + sX[SX_E * SX_E + o] + sX[2 * SX_E * SX_E + o]) / 3.0f;
        }
        __syncthreads();
        // ---- features, 2-pixel ring; zero outside the image (the first embedding convolution pads ITS input with zeros) ----
        for (int i = tid; i < SF_N; i += ENT) {
            const int q = i % SF_E, r = (i / SF_E) % SF_E, f = i / (SF_E * SF_E);
            const int yy = y0 - 2 + r, xx = x0 - 2 + q;
            float v = 0.0f;
            if (inimg(yy, xx, H, W)) {
                if (f < 3) {
                    const float* gp = sG + r * SG_E + q;      // top-left of the 3x3 window around gray position (r+1, q+1)
                    const float a00 = gp[0], a01 = gp[1], a02 = gp[2], a10 = gp[SG_E], a11 = gp[SG_E + 1], a12 = gp[SG_E + 2],
                                a20 = gp[2 * SG_E], a21 = gp[2 * SG_E + 1], a22 = gp[2 * SG_E + 2];
                    if (f == 0) v = (a02 - a00) + 2.0f * (a12 - a10) + (a22 - a20);                       // sobel_x (encoder.py:12)
                    else if (f == 1) v = (a20 - a00) + 2.0f * (a21 - a01) + (a22 - a02);                  // sobel_y (:13)
                    else v = (a00 + a02 + a20 + a22) + 2.0f * (a01 + a10 + a12 + a21) - 12.0f * a11;      // laplacian (:25)
                } else {
                    const float* xp = sX + (f - 3) * SX_E * SX_E + r * SX_E + q;     // top-left of the 5x5 window around x position (r+2, q+2)
#pragma unroll
                    for (int a = 0; a < 5; ++a)
#pragma unroll
                        for (int c = 0; c < 5; ++c) v = fmaf(k.gauss[a * 5 + c], xp[a * SX_E + c], v);
                }
                if (feats != nullptr && r >= 2 && r < 2 + ET && q >= 2 && q < 2 + ET)
                    feats[((size_t)b * EF + f) * plane + (size_t)yy * W + xx] = v;
            }
            sF[i] = v;
        }
        __syncthreads();
        // ---- h1 = relu(conv3x3(features) + b1), 1-pixel ring, zero outside the image; item = (position, 4 outputs) ----
        for (int it = tid; it < SH_E * SH_E * 4; it += ENT) {
            const int oq = it & 3, pos = it >> 2;
            const int q = pos % SH_E, r = pos / SH_E;
            const int yy = y0 - 1 + r, xx = x0 - 1 + q;
            float acc[4] = {sB1[4 * oq], sB1[4 * oq + 1], sB1[4 * oq + 2], sB1[4 * oq + 3]};
#pragma unroll
            for (int ci = 0; ci < EF; ++ci)
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float v = sF[ci * SF_E * SF_E + (r + tap / 3) * SF_E + q + tap % 3];
                    const float4 w = *reinterpret_cast<const float4*>(sW1 + (ci * 9 + tap) * EE + 4 * oq);
                    acc[0] = fmaf(v, w.x, acc[0]); acc[1] = fmaf(v, w.y, acc[1]); acc[2] = fmaf(v, w.z, acc[2]); acc[3] = fmaf(v, w.w, acc[3]);
                }
            const bool in = inimg(yy, xx, H, W);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float h = in ? fmaxf(acc[j], 0.0f) : 0.0f;
                sH[(4 * oq + j) * SH_P + r * SH_E + q] = h;
                if (in && h1g != nullptr && r >= 1 && r <= ET && q >= 1 && q <= ET)
                    h1g[((size_t)b * EE + 4 * oq + j) * plane + (size_t)yy * W + xx] = h;
            }
        }
        __syncthreads();
        // ---- out = conv3x3(h1): thread = output position, 16 accumulators ----
        {
            const int q = tid % ET, r = tid / ET;
            const int yy = y0 + r, xx = x0 + q;
            float acc[EE];
#pragma unroll
            for (int o = 0; o < EE; ++o) acc[o] = 0.0f;
#pragma unroll 4
            for (int ci = 0; ci < EE; ++ci)
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float v = sH[ci * SH_P + (r + tap / 3) * SH_E + q + tap % 3];
                    const float4* w = reinterpret_cast<const float4*>(sW2 + (ci * 9 + tap) * EE);
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const float4 ww = w[o4];
                        acc[4 * o4] = fmaf(v, ww.x, acc[4 * o4]); acc[4 * o4 + 1] = fmaf(v, ww.y, acc[4 * o4 + 1]);
                        acc[4 * o4 + 2] = fmaf(v, ww.z, acc[4 * o4 + 2]); acc[4 * o4 + 3] = fmaf(v, ww.w, acc[4 * o4 + 3]);
                    }
                }
            if (yy < H && xx < W) {
                float* op = out + (size_t)b * C_out * plane + (size_t)yy * W + xx;
                for (int c = 0; c < c0; ++c) op[c * plane] = 0.0f;                    // the zero-padded leading channels (nca.py:199-203)
#pragma unroll
                for (int o = 0; o < EE; ++o) op[(c0 + o) * plane] = acc[o];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// backward: d(out) -> gW2, (through relu) gW1, gb1.  Persistent CTAs keep their partial sums in registers over all their tiles
// and flush once (one atomic per weight and CTA).
// ------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ENT) encoder_bwd_kernel(int B, int H, int W, int C_g, int c0, const float* __restrict__ feats,
                                                          const float* __restrict__ h1g, const float* __restrict__ w2,
                                                          const float* __restrict__ g, float* __restrict__ gw1,
                                                          float* __restrict__ gb1, float* __restrict__ gw2, int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float sm[];
    float* sW2 = sm;                        // [o][tap][i]  (transposed use: g_h1[i] = sum_{o,tap} g[o][shifted] * w2[o][i][tap])
    float* sGo = sW2 + W2_N;                // [o][325]: d(out), 1-pixel ring, zero outside the image
    float* sH = sGo + SH_N;                 // [i][325]: h1, 1-pixel ring
    float* sFt = sH + SH_N;                 // [f][325]: features, 1-pixel ring
    float* sGa = sFt + EF * SH_P;           // [o][257]: g_a1 = g_h1 * [h1 > 0] on the tile
    constexpr int SGA_P = ET * ET + 1;
    const int tid = threadIdx.x;
    for (int i = tid; i < W2_N; i += ENT) { const int ci = i % EE, tap = (i / EE) % 9, o = i / (EE * 9); sW2[i] = w2[(o * EE + ci) * 9 + tap]; }
    const size_t plane = (size_t)H * W;
    const int n_tiles = B * tiles_x * tiles_y;
    // gW2: thread (o = tid / 16, i = tid % 16) owns the 9 taps; gW1: threads 0..95 (o = tid / 6, i = tid % 6); gb1: threads 96..111
    float a2[9], a1[9], ab = 0.0f;
#pragma unroll
    for (int t = 0; t < 9; ++t) { a2[t] = 0.0f; a1[t] = 0.0f; }
    const int o2 = tid / EE, i2 = tid % EE, o1 = tid / EF, i1 = tid % EF;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / (tiles_x * tiles_y), ty = (tile / tiles_x) % tiles_y, tx = tile % tiles_x;
        const int y0 = ty * ET, x0 = tx * ET;
        __syncthreads();
        for (int i = tid; i < EE * SH_E * SH_E; i += ENT) {
            const int q = i % SH_E, r = (i / SH_E) % SH_E, c = i / (SH_E * SH_E);
            const int yy = y0 - 1 + r, xx = x0 - 1 + q;
            const bool in = inimg(yy, xx, H, W);
            const size_t pix = (size_t)yy * W + xx;
            sGo[c * SH_P + r * SH_E + q] = in ? __ldg(g + ((size_t)b * C_g + c0 + c) * plane + pix) : 0.0f;
            sH[c * SH_P + r * SH_E + q] = in ? __ldg(h1g + ((size_t)b * EE + c) * plane + pix) : 0.0f;
            if (c < EF) sFt[c * SH_P + r * SH_E + q] = in ? __ldg(feats + ((size_t)b * EF + c) * plane + pix) : 0.0f;
        }
        __syncthreads();
        // ---- g_a1 on the tile: thread = position; transposed 3x3: input (y, x) feeds outputs (y - ky + 1, x - kx + 1) ----
        {
            const int q = tid % ET, r = tid / ET;
            float acc[EE];
#pragma unroll
            for (int i = 0; i < EE; ++i) acc[i] = 0.0f;
#pragma unroll 4
            for (int o = 0; o < EE; ++o)
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float v = sGo[o * SH_P + (r + 2 - tap / 3) * SH_E + q + 2 - tap % 3];
                    const float4* w = reinterpret_cast<const float4*>(sW2 + (o * 9 + tap) * EE);
#pragma unroll
                    for (int i4 = 0; i4 < 4; ++i4) {
                        const float4 ww = w[i4];
                        acc[4 * i4] = fmaf(v, ww.x, acc[4 * i4]); acc[4 * i4 + 1] = fmaf(v, ww.y, acc[4 * i4 + 1]);
                        acc[4 * i4 + 2] = fmaf(v, ww.z, acc[4 * i4 + 2]); acc[4 * i4 + 3] = fmaf(v, ww.w, acc[4 * i4 + 3]);
                    }
                }
            const bool in = y0 + r < H && x0 + q < W;
#pragma unroll
            for (int i = 0; i < EE; ++i)
                sGa[i * SGA_P + tid] = (in && sH[i * SH_P + (r + 1) * SH_E + q + 1] > 0.0f) ? acc[i] : 0.0f;
        }
        __syncthreads();
        // ---- weight-gradient partial sums over the tile ----
        for (int p = 0; p < ET * ET; ++p) {
            const int q = p % ET, r = p / ET;
            const float go = sGo[o2 * SH_P + (r + 1) * SH_E + q + 1];
            const float* hp = sH + i2 * SH_P + r * SH_E + q;
#pragma unroll
            for (int t = 0; t < 9; ++t) a2[t] = fmaf(go, hp[(t / 3) * SH_E + t % 3], a2[t]);
            if (tid < EE * EF) {
                const float ga = sGa[o1 * SGA_P + p];
                const float* fp = sFt + i1 * SH_P + r * SH_E + q;
#pragma unroll
                for (int t = 0; t < 9; ++t) a1[t] = fmaf(ga, fp[(t / 3) * SH_E + t % 3], a1[t]);
            } else if (tid < EE * EF + EE) {
                ab += sGa[(tid - EE * EF) * SGA_P + p];
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(gw2 + (o2 * EE + i2) * 9 + t, a2[t]);
    if (tid < EE * EF) {
#pragma unroll
        for (int t = 0; t < 9; ++t) atomicAdd(gw1 + (o1 * EF + i1) * 9 + t, a1[t]);
    } else if (tid < EE * EF + EE) {
        atomicAdd(gb1 + tid - EE * EF, ab);
    }
}

EncoderConsts encoder_consts() {
    EncoderConsts k;
    const double pi = 3.14159265358979323846;
    float raw[25];
    float sum = 0.0f;
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j) {
            // torch.tensor([... python floats ...]) -> float32 values, then kernel / torch.sum(kernel) in float32 (encoder.py:60-61)
            const double v = (1.0 / (2.0 * pi)) * exp(-((double)((i - 2) * (i - 2) + (j - 2) * (j - 2))) / 2.0);
            raw[i * 5 + j] = (float)v;
        }
    for (int i = 0; i < 25; ++i) sum += raw[i];
    for (int i = 0; i < 25; ++i) k.gauss[i] = raw[i] / sum;
    return k;
}

constexpr size_t FWD_SMEM = (size_t)(W1_N + EE + W2_N + SX_N + SG_N + SF_N + SH_N) * sizeof(float);
constexpr size_t BWD_SMEM = (size_t)(W2_N + 2 * SH_N + EF * SH_P + EE * (ET * ET + 1)) * sizeof(float);

}  // namespace

extern "C" {

int nca_encoder_forward(int32_t B, int32_t channels, int32_t H, int32_t W, int32_t embedding_dim, const float* x, const float* w1,
                        const float* b1, const float* w2, float* feats, float* hidden, float* goal, int32_t goal_channels,
                        void* stream) {
    NCA_CHECK_ARG(B > 0 && H > 0 && W > 0 && x && w1 && b1 && w2 && goal, "bad encoder_forward arguments");
    if (channels != EC || embedding_dim != EE) {
        nca_set_error("the fused ImageEncoder supports channels == 3 and embedding_dim == 16 (got %d, %d)", channels, embedding_dim);
        return NCA_ERR_UNSUPPORTED;
    }
    NCA_CHECK_ARG(goal_channels >= EE, "goal tensor has %d channels, the embedding needs %d", goal_channels, EE);
    NCA_CHECK_ARG((feats == nullptr) == (hidden == nullptr), "feats and hidden are kept together (both or neither)");
    int rc = nca_check_device();
    if (rc) return rc;
    const int tx = (W + ET - 1) / ET, ty = (H + ET - 1) / ET, n = B * tx * ty;
    NCA_CUDA_OK(cudaFuncSetAttribute(encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    int grid = nca_sm_count() * 3;
    if (grid > n) grid = n;
    encoder_fwd_kernel<<<grid, ENT, FWD_SMEM, (cudaStream_t)stream>>>(B, H, W, goal_channels, goal_channels - EE, x, w1, b1, w2, feats,
                                                                      hidden, goal, encoder_consts(), tx, ty);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_encoder_backward(int32_t B, int32_t channels, int32_t H, int32_t W, int32_t embedding_dim, const float* feats,
                         const float* hidden, const float* w2, const float* g_goal, int32_t goal_channels, float* gw1, float* gb1,
                         float* gw2, void* stream) {
    NCA_CHECK_ARG(B > 0 && H > 0 && W > 0 && feats && hidden && w2 && g_goal && gw1 && gb1 && gw2, "bad encoder_backward arguments");
    if (channels != EC || embedding_dim != EE) {
        nca_set_error("the fused ImageEncoder supports channels == 3 and embedding_dim == 16 (got %d, %d)", channels, embedding_dim);
        return NCA_ERR_UNSUPPORTED;
    }
    NCA_CHECK_ARG(goal_channels >= EE, "goal gradient has %d channels, the embedding needs %d", goal_channels, EE);
    int rc = nca_check_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    NCA_CUDA_OK(cudaMemsetAsync(gw1, 0, (size_t)EE * EF * 9 * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gb1, 0, (size_t)EE * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw2, 0, (size_t)EE * EE * 9 * sizeof(float), s));
    const int tx = (W + ET - 1) / ET, ty = (H + ET - 1) / ET, n = B * tx * ty;
    NCA_CUDA_OK(cudaFuncSetAttribute(encoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    int grid = nca_sm_count() * 2;
    if (grid > n) grid = n;
    encoder_bwd_kernel<<<grid, ENT, BWD_SMEM, s>>>(B, H, W, goal_channels, goal_channels - EE, feats, hidden, w2, g_goal, gw1, gb1, gw2, tx, ty);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

}  // extern "C"
