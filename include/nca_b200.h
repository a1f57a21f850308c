/*
 * nca_b200.h — C ABI of the B200-native NCA step (libnca_b200.so).
 *
 * The reference (smehra34/Video-Stylization-with-NCA) has no FFI layer: its boundary for this
 * path is the Python nn.Module API of DyNCA / ConditionedNCA.  These entry points
 * are what a torch-side binding of that API calls; each one names the reference code it replaces.
 * Plain C types only: device pointers, sizes, a cudaStream_t passed as void*.
 *
 * Conventions
 *   - all tensors fp32, contiguous, NCHW, resident on the current CUDA device;
 *   - the callee never allocates, frees or retains device memory: state history, gradients and
 *     workspace are caller-owned (query sizes with the *_workspace_bytes functions);
 *   - every launch goes to `stream`; no internal synchronisation; graph-capturable;
 *   - return value 0 = ok, negative = error (message: nca_last_error(), thread local);
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns NCA_ERR_CUDA.
 */
#ifndef NCA_B200_H_
#define NCA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NCA_B200_ABI_VERSION 6

enum { NCA_OK = 0, NCA_ERR_ARG = -1, NCA_ERR_UNSUPPORTED = -2, NCA_ERR_CUDA = -3, NCA_ERR_WORKSPACE = -4 };

/* padding_mode of DyNCA.perceive_torch (ExtraChannels/models/dynca.py:81) */
enum { NCA_PAD_CONSTANT = 0, NCA_PAD_CIRCULAR = 1, NCA_PAD_REPLICATE = 2, NCA_PAD_REFLECT = 3 };
/* conditioning inputs appended to the perception vector (dynca.py:108-109) */
enum { NCA_COND_NONE = 0,   /* pos_emb=None / conditioning=None                                   */
       NCA_COND_CPE = 1,    /* CPE2D computed in-kernel from coordinates (dynca.py:180-207), cc=2 */
       NCA_COND_TENSOR = 2  /* caller-supplied [B,cc,H,W] (CD: EdgeExtractor output, cc=3)        */ };
/* arithmetic of the update MLP */
enum { NCA_PREC_FP32 = 0,   /* CUDA-core FFMA, parity 1e-5 per step                      */
       NCA_PREC_BF16 = 1,   /* tcgen05 BF16 operands, fp32 accumulate in TMEM, parity 1e-2 */
       NCA_PREC_F16X3 = 2   /* tcgen05, every operand split into FP16 hi + lo (~22 significant bits) and every product
                               issued as Ah.Bh + Al.Bh + Ah.Bl (fp32 accumulate): fp32-grade parity (state 1e-5,
                               gradients 1e-4) on the tensor cores.  Weights are pre-scaled by 2^8 and gradient operands
                               by a per-launch power of two (undone exactly), so |w| < 255 and |perception|, |hidden|
                               < 65504 are required (beyond: NaN).  DyNCA only (ConditionedNCA: treated as FP32) */ };
/* fire mask source */
enum { NCA_MASK_SUPPLIED = 0, /* float [T,B,1,H,W], 1 = fire (parity runs)                 */
       NCA_MASK_PHILOX = 1    /* in-kernel Philox4x32-10 keyed on (seed, t0+t, b, y, x)    */ };

/* Describes one DyNCA model + batch geometry.  Replaces the ctor arguments of
 * ExtraChannels/models/dynca.py:30-69 and ConditioneDyNCA/models/dynca.py:30-73. */
typedef struct NcaDyncaDesc {
    int32_t B, C, H, W;      /* state [B,C,H,W]                                                 */
    int32_t fc;              /* hidden width of the update MLP (fc_dim)                          */
    int32_t cond_kind;       /* NCA_COND_*                                                       */
    int32_t cc;              /* number of cond channels (0, 2 for CPE, any for TENSOR)           */
    int32_t pad_mode;        /* NCA_PAD_*                                                        */
    int32_t n_scales;        /* 1: perception_scales=[0]; 2: [0,1]  (H and W even)               */
    int32_t precision;       /* NCA_PREC_*                                                       */
    int32_t mask_mode;       /* NCA_MASK_*                                                       */
    float   update_rate;     /* fire probability (dynca.py:113,121)                              */
} NcaDyncaDesc;

/* Weights in the reference state_dict layout:
 *   w1 [fc, 4C+cc] (= w1.weight[:, :, 0, 0]),  b1 [fc],  w2 [C, fc],  b2 [C].                   */
typedef struct NcaDyncaWeights {
    const float* w1; const float* b1; const float* w2; const float* b2;
} NcaDyncaWeights;
typedef struct NcaDyncaWeightGrads {
    float* w1; float* b1; float* w2; float* b2;   /* overwritten, not accumulated */
} NcaDyncaWeightGrads;

const char* nca_last_error(void);
int nca_abi_version(void);
/* number of kernels launched by this library (all threads of the process) since the last reset */
long long nca_launch_count(void);
void nca_launch_count_reset(void);

/* ---- DyNCA ------------------------------------------------------------------------------- */

/* DyNCA.perceive_multiscale (dynca.py:98-111): z [B, 4C+cc, H, W]. */
int nca_dynca_perceive(const NcaDyncaDesc* d, const float* x, const float* cond, float* z, void* stream);

/* EdgeExtractor.forward (ConditioneDyNCA/models/dynca.py:204-213): img [B,1,H,W] -> out [B,3,H,W];
 * tanh_transform != 0 applies tanh. */
int nca_edge_extract(int B, int H, int W, const float* img, int tanh_transform, float* out, void* stream);

/* DyNCA.forward_nsteps (dynca.py:158-167): T steps of DyNCA.forward (dynca.py:113-128).
 *   states : keep_history != 0 -> [T+1,B,C,H,W], caller has written states[0]; step t writes states[t+1]
 *            keep_history == 0 -> [2,B,C,H,W] ping-pong, input in slot 0, result in slot (T & 1)
 *   cond   : [B,cc,H,W] when cond_kind == NCA_COND_TENSOR else NULL
 *   masks  : [T,B,1,H,W] when mask_mode == NCA_MASK_SUPPLIED else NULL
 *   seed,t0: Philox key / first step counter when mask_mode == NCA_MASK_PHILOX
 *   coarse_hist: optional (may be NULL), only used when n_scales == 2 and keep_history != 0: float
 *            [T+1,B,C,H/2,W/2]; the forward writes the 2x2-mean coarse state of every states[t] there so that
 *            nca_dynca_backward does not have to recompute it (pass the same buffer to it)
 *   op_hist: optional (may be NULL), only used when keep_history != 0 and nca_dynca_op_hist_bytes(d, T) != 0: that many
 *            bytes, 16-byte aligned.  The forward records the bf16 perception operand Z of every tile of every step there
 *            (one bulk copy per tile; with two scales Z = z_fine + up(z_coarse) is ONE operand) so that nca_dynca_backward loads
 *            it instead of recomputing the perception: memory (128-160 B per cell and step) traded for time; results are
 *            bit-identical with and without it
 *   workspace: nca_dynca_workspace_bytes(d, 0) bytes                                              */
int nca_dynca_forward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                      uint64_t seed, int32_t t0, int32_t T, int32_t keep_history, float* states, float* coarse_hist,
                      void* op_hist, void* workspace, size_t workspace_bytes, void* stream);

/* Size of the optional operand history of a T-step rollout; 0 when the description does not dispatch to the kernels that
 * use one (then pass op_hist = NULL). */
size_t nca_dynca_op_hist_bytes(const NcaDyncaDesc* d, int32_t T);

/* BPTT through the T steps recorded in `states` ([T+1,B,C,H,W] from nca_dynca_forward with keep_history).
 * Replaces autograd's replay of dynca.py:113-128 x T.
 *   g_final  : dL/d states[T]                       [B,C,H,W] or NULL (zeros)
 *   g_taps   : host array of n_taps device pointers, g_taps[i] = dL/d (tap_scale * states[tap_steps[i]][:, :tap_c]),
 *              each [B,tap_c,H,W]: the gradients of the rgb list of forward_nsteps(return_middle_feature=True)
 *              (dynca.py:130-131,163-165; fit_video_motion.py:230-235 uses steps 1, 65, 129)
 *   tap_steps: host array [n_taps], strictly increasing, each in 1..T
 *   gx0      : dL/d states[0]  [B,C,H,W] (written)
 *   gw       : weight gradients, reference layout (written)
 *   coarse_hist: NULL or the buffer nca_dynca_forward filled (n_scales == 2)
 *   op_hist  : NULL or the operand history nca_dynca_forward filled for the same d, T.  NULL: every BPTT step is preceded by
 *              a launch of the forward kernel that records the operand of states[t] into the workspace and stops
 *   workspace: nca_dynca_workspace_bytes(d, 1) bytes (includes one step of operands for the NULL case)
 * State gradients are accumulated with red.add (fp32 summation order is not deterministic).          */
int nca_dynca_backward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                       uint64_t seed, int32_t t0, int32_t T, const float* states, const float* coarse_hist,
                       const void* op_hist, const float* g_final, const float* const* g_taps, const int32_t* tap_steps, int32_t n_taps,
                       int32_t tap_c, float tap_scale, float* gx0, const NcaDyncaWeightGrads* gw,
                       void* workspace, size_t workspace_bytes, void* stream);

size_t nca_dynca_workspace_bytes(const NcaDyncaDesc* d, int32_t backward);

/* Which step kernel a description dispatches to (tests / reports): 0 = fp32 CUDA cores, 1 = tcgen05 with 4x32 tiles
 * and cp.async staging (any shape), 2 = tcgen05 with 8x16 tiles, TMA staging and the coarse scale on the tensor
 * cores (W % 4 == 0, W % 8 == 0 for two scales, fc % 32 == 0), 3 = the 4x32-tile tcgen05 kernels with split-precision
 * (hi + lo) operands (NCA_PREC_F16X3; falls back to 0 when the doubled operand images exceed shared memory).
 * Negative = invalid description. */
int nca_dynca_kernel_variant(const NcaDyncaDesc* d, int32_t backward);

/* The Philox fire mask the kernels generate, materialised as float [T,B,1,H,W] (tests / debugging).
 * enc != 0 uses the ConditionedNCA rule (u < rate, EncoderConditioning/nca.py:165-174). */
int nca_philox_mask(int32_t B, int32_t H, int32_t W, float rate, int32_t enc, uint64_t seed, int32_t t0,
                    int32_t T, float* out, void* stream);

/* The same masks for steps t0 + *t0_dev .. (t0_dev: device uint32[1] or NULL): the step counter lives on the device, so a
 * captured CUDA graph (mask draw + rollout with NCA_MASK_SUPPLIED + a counter increment) draws new masks on every replay. */
int nca_philox_mask_at(int32_t B, int32_t H, int32_t W, float rate, int32_t enc, uint64_t seed, int32_t t0, const uint32_t* t0_dev,
                       int32_t T, float* out, void* stream);

/* ---- ConditionedNCA (EncoderConditioning/nca.py) --------------------------------------- */

typedef struct NcaEncDesc {
    int32_t B, C, H, W;       /* state [B,C,H,W], C = num_channels (nca.py:79)               */
    int32_t hid;              /* UpdateNet hidden width (64, nca.py:40-46)                    */
    int32_t living_dim;       /* living_channel_dim (nca.py:66)                               */
    int32_t mask_mode;        /* NCA_MASK_*                                                   */
    float   alive_thr;        /* 0.1 (nca.py:163)                                             */
    float   fire_rate;        /* cell_fire_rate (nca.py:67)                                   */
    float   clamp;            /* 10.0 (nca.py:194)                                            */
    int32_t precision;        /* NCA_PREC_*: BF16 runs the update MLP of the forward AND of the BPTT on tcgen05
                                 (enc_tc.cu; W % 4 == 0, alive_thr >= 0; other shapes use the fp32 kernels) */
} NcaEncDesc;

/* wp [3C,1,3,3] perception_net.weight; wa [hid,3C], ba [hid]; wb [hid,hid], bb [hid]; wc [C,hid]. */
typedef struct NcaEncWeights {
    const float* wp; const float* wa; const float* ba; const float* wb; const float* bb; const float* wc;
} NcaEncWeights;
typedef struct NcaEncWeightGrads {
    float* wp; float* wa; float* ba; float* wb; float* bb; float* wc;
} NcaEncWeightGrads;

/* ConditionedNCA.grow's T-loop of ConditionedNCA.forward/update (nca.py:176-209), goal [B,C,H,W]
 * already encoded and zero-padded (nca.py:198-203).  states/masks/workspace as for DyNCA;
 * supplied masks are (u < rate) as float.
 *   life_hist: uint8 [T,B,H,W] (keep_history != 0) written by forward: life_hist[t] = pre & post life mask of
 *              step t (nca.py:181,191-193); backward reads it (the post-update alive mask depends on the
 *              updated state of the 3x3 neighbourhood and is not recoverable from states[t+1]). May be NULL
 *              when keep_history == 0. */
int nca_enc_forward(const NcaEncDesc* d, const NcaEncWeights* w, const float* goal, const float* masks,
                    uint64_t seed, int32_t t0, int32_t T, int32_t keep_history, float* states, uint8_t* life_hist,
                    void* workspace, size_t workspace_bytes, void* stream);
/* BPTT through nca_enc_forward's T steps. g_goal [B,C,H,W] = dL/d goal (consumed by nca_encoder_backward below, the
 * weight-gradient pass of the ImageEncoder, EncoderConditioning/encoder.py); gx0, g_goal and gw are written. */
int nca_enc_backward(const NcaEncDesc* d, const NcaEncWeights* w, const float* goal, const float* masks,
                     uint64_t seed, int32_t t0, int32_t T, const float* states, const uint8_t* life_hist,
                     const float* g_final, float* gx0, float* g_goal, const NcaEncWeightGrads* gw,
                     void* workspace, size_t workspace_bytes, void* stream);
size_t nca_enc_workspace_bytes(const NcaEncDesc* d, int32_t backward);

/* ImageEncoder of the encoder-conditioned NCA (EncoderConditioning/encoder.py:5-64; called once per rollout, nca.py:198), fused:
 *   x [B,3,H,W] -> [sobel_x, sobel_y, laplacian of mean_rgb(x) | 5x5 gaussian blur per colour channel] -> conv3x3(6->16)+b1, ReLU
 *   -> conv3x3(16->16)  = goal encoding, written into the LAST 16 channels of goal [B,goal_channels,H,W]; the leading channels are
 *   zeroed (the zero padding of nca.py:199-203), i.e. `goal` is exactly the tensor nca_enc_forward consumes.
 *   w1 [16,6,3,3] = encoder.embed.0.weight, b1 [16] = embed.0.bias, w2 [16,16,3,3] = embed.2.weight.
 *   feats [B,6,H,W] and hidden [B,16,H,W] (post-ReLU) are kept for the backward when given (both or neither).
 * Supported: channels == 3, embedding_dim == 16 (the reference's configuration, train.py:84); else NCA_ERR_UNSUPPORTED. */
int nca_encoder_forward(int32_t B, int32_t channels, int32_t H, int32_t W, int32_t embedding_dim, const float* x, const float* w1,
                        const float* b1, const float* w2, float* feats, float* hidden, float* goal, int32_t goal_channels,
                        void* stream);
/* Weight gradients of the ImageEncoder from d(goal) = the last 16 channels of g_goal [B,goal_channels,H,W] as written by
 * nca_enc_backward (replaces autograd's replay of encoder.py:37-57; the input image gets no gradient: it is data,
 * conditioned_trainer.py:118-137).  gw1 [16,6,3,3], gb1 [16], gw2 [16,16,3,3] are written. */
int nca_encoder_backward(int32_t B, int32_t channels, int32_t H, int32_t W, int32_t embedding_dim, const float* feats,
                         const float* hidden, const float* w2, const float* g_goal, int32_t goal_channels, float* gw1, float* gb1,
                         float* gw2, void* stream);

/* ---- the callers either side of the step (N1 pool + optimizer, N4 overflow loss, N2 frame stream) -------- */

/* Batch assembly from the sample pool — ExtraChannels/experiments.py:203-211 (ConditioneDyNCA/experiments.py:210-218):
 *   input_states = nca_pool[batch_idx]; input_states[:1] = seed; input_states = cat((input_states, aux_gs), 1)
 * and EncoderConditioning/conditioned_trainer.py:107-113,167 (batch[:2] = generate_seed(2)), in one pass.
 *   pool [N,Cp,H,W]; idx device int64 [B] (pool slots; an out-of-range slot reads as zeros);
 *   extra [B,Cx,H,W] or NULL with Cx == 0: appended as channels Cp..Cp+Cx-1 (the EC conditioning channel);
 *   the first inject_n samples, and every sample b with reseed_flags[b] != 0 (device uint8 [B] or NULL), are replaced by
 *   seed_state [Cp,H,W] (NULL = zeros, seed_mode 'zeros' dynca.py:142-143);
 *   out [B,Cp+Cx,H,W] (written). */
int nca_pool_gather(int32_t N, int32_t Cp, int32_t H, int32_t W, const float* pool, const int64_t* idx, int32_t B,
                    const float* extra, int32_t Cx, const float* seed_state, int32_t inject_n, const uint8_t* reseed_flags,
                    float* out, void* stream);

/* Dead-sample test of ConditionedNCATrainer.sample_batch (conditioned_trainer.py:107-112):
 *   `torch.sum(self.nca.alive(batch[i].unsqueeze(0))) == 0.0` for every sampled slot, without the B host round trips:
 *   flags[b] = 1 when pool[idx[b]][living_dim] has no value > alive_thr (the 3x3 max-pool of nca.py:152-163 is non-empty
 *   exactly when the plane itself is), else 0.  flags: device uint8 [B], fed to nca_pool_gather as reseed_flags. */
int nca_pool_dead_flags(int32_t N, int32_t Cp, int32_t H, int32_t W, const float* pool, const int64_t* idx, int32_t B,
                        int32_t living_dim, float alive_thr, uint8_t* flags, void* stream);

/* Pool write-back — experiments.py:259: nca_pool[batch_idx] = nca_states_after[:, :12]; conditioned_trainer.py:155-156.
 *   states [B,C,H,W], the first Cp channels go to pool[idx[b]]; idx must not repeat (np.random.choice(replace=False),
 *   random.sample); an out-of-range slot is skipped. */
int nca_pool_scatter(int32_t N, int32_t Cp, int32_t H, int32_t W, float* pool, const int64_t* idx, int32_t B,
                     const float* states, int32_t C, void* stream);

/* Per-parameter gradient normalisation fused with the Adam update, ONE launch for the whole model —
 * experiments.py:252-255 (`p.grad /= (p.grad.norm() + 1e-8)` for every parameter; `optimizer.step()`) and
 * conditioned_trainer.py:134-137 (eps 1e-10).  torch.optim.Adam semantics (amsgrad off, weight_decay 0): exp_avg lerp,
 * exp_avg_sq, bias corrections for the 1-based `step`, denom = sqrt(v)/sqrt(bc2) + eps, p -= lr/bc1 * m/denom.
 *   params / grads / exp_avg / exp_avg_sq: HOST arrays of n_tensors device pointers (n_tensors <= NCA_ADAM_MAX_TENSORS),
 *   numel: HOST array of element counts; normalise == 0 skips the normalisation; the normalised gradient is written back
 *   to grads (as the reference's in-place division does) or zeros when zero_grads != 0 (optimizer.zero_grad()).
 *   lr is the current learning rate (MultiStepLR stays on the host: it only changes this number). */
#define NCA_ADAM_MAX_TENSORS 16
int nca_normalized_adam_step(int32_t n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, int32_t step, float lr, float beta1, float beta2,
                             float eps, float norm_eps, int32_t normalise, int32_t zero_grads, void* stream);

/* Loss.get_overflow_loss (ExtraChannels/utils/loss/loss.py:33-36; EncoderConditioning/loss/loss.py):
 *   loss = mean |x - clamp(x, -1, 1)| over the n elements of the final state, and in the same pass its gradient
 *   grad_scale * sign(x) * [|x| > 1] / n written (accumulate == 0) or added (accumulate != 0) to grad_out (NULL = loss only):
 *   the BPTT's g_final gets the overflow term without a second pass over the state.  grad_scale_dev (device float[1] or NULL)
 *   multiplies grad_scale on the device: autograd's incoming dL/dloss is used without a host round trip or an extra pass.
 *   loss_out: device float[1]; workspace: nca_overflow_workspace_bytes() bytes.  Fixed reduction order (reproducible). */
size_t nca_overflow_workspace_bytes(void);
int nca_overflow_loss(const float* x, size_t n, float* loss_out, float* grad_out, float grad_scale, const float* grad_scale_dev,
                      int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Inference stream, per target frame — ExtraChannels/utils/misc/video_utils.py:72: h = cat((h, RGBToGrayscale(frame)), 1)
 * (preprocess_texture.py:178-179: mean over the three channels) and :76: h = nca_state[:, :-1].  The conditioning channel
 * lives in the state buffer: this writes mean_rgb(frame) into channel `ch` of state [B,C,H,W] in place, so there is no
 * per-frame cat / strip.  frame_rgb [B,3,H,W].  (C == 1, ch == 0 gives the plain grayscale image, CD's cond_img.) */
int nca_frame_to_cond_channel(int32_t B, int32_t C, int32_t H, int32_t W, const float* frame_rgb, float* state, int32_t ch,
                              void* stream);

/* Frame read-out — video_utils.py:78-82 + VideoWriter.add (:20-27): img = clip(scale * state[:, :3], -1, 1);
 * img = (img + 1) / 2; uint8(clip(img, 0, 1) * 255) (truncation), channels-last.  scale = 2 (DyNCA.to_rgb, dynca.py:130-131).
 *   state [B,C,H,W] (C >= 3) -> out_hwc uint8 [B,H,W,3]. */
int nca_state_to_rgb8(int32_t B, int32_t C, int32_t H, int32_t W, const float* state, float scale, uint8_t* out_hwc,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NCA_B200_H_ */
