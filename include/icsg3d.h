/*
 * libicsg3d — C ABI of the B200-native (sm_100a) hot path of by256/icsg3d.
 *
 * The reference (pure Keras 2.3.1 / TF 2.1 Python) has no FFI of its own: its hot path is the set of
 * TF ops its two Keras graphs lower to.  Each entry point below replaces one such op
 * family; the comment on each cites the reference lines whose layers dispatch to it.  The Python
 * host side (icsg3d_b200/) binds these with ctypes and mirrors the reference's model-builder API
 * (vae/lattice_vae.py::LatticeDFCVAE, unet/unet.py::AtomUnet).  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; icsg3d_last_error() gives a thread-local message;
 *   - all tensor arguments are caller-owned DEVICE pointers; nothing is allocated, nothing is synchronised;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - activations are channels-last NDHWC; "ld" arguments are the channel stride (elements) of one voxel,
 *     so that a channel slice of a wider (concatenated) tensor can be addressed in place;
 *   - bf16 storage + fp32 accumulation; statistics and losses in fp32/fp64.
 */
#ifndef ICSG3D_H_
#define ICSG3D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICSG3D_OK 0
#define ICSG3D_ERR_INVALID (-1)
#define ICSG3D_ERR_CUDA (-2)
#define ICSG3D_ERR_UNSUPPORTED (-3)

/* activation codes (Keras ReLU(), LeakyReLU(alpha=0.3): lattice_vae.py:175,226; unet.py:277) */
#define ICSG3D_ACT_NONE 0
#define ICSG3D_ACT_RELU 1
#define ICSG3D_ACT_LEAKY 2

/* post-ops fused into the BatchNorm apply pass */
#define ICSG3D_POST_NONE 0
#define ICSG3D_POST_POOL2 1 /* MaxPool3D(2): lattice_vae.py:176, unet.py:282,291,300 */
#define ICSG3D_POST_UP2 2   /* UpSampling3D(2): lattice_vae.py:217, unet.py:309,319,329 */

#define ICSG3D_DT_BF16 0
#define ICSG3D_DT_F32 1
#define ICSG3D_DT_F64 2

const char* icsg3d_last_error(void);
int icsg3d_version(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int icsg3d_sm_count(void);
/* Number of kernels this library has launched (or captured into a CUDA graph) in this process. */
int64_t icsg3d_launch_count(void);
/* Spin budget of the peer-memory all-reduce kernels (data parallel, see icsg3d_bn_reduce_allreduce_* and
 * icsg3d_adam_keras_allreduce_step): a rank that waits longer than `seconds` for a peer traps its context.  Default 600 s
 * (or the environment variable ICSG3D_PEER_TIMEOUT_S), so that ordinary host-side rank skew (batch loading, checkpoint
 * writes, a lazy graph capture) does not kill training; launches captured into a CUDA graph keep the value they were
 * captured with. */
int icsg3d_set_peer_timeout(double seconds);
/* Protocol of the BatchNorm statistic exchange: 1 (default) = flag-in-word — every 8-byte word pushed to a peer carries
 * its epoch tag, so there is no system fence and no separate flag round trip per exchange; 0 (or ICSG3D_PEER_LL=0) = data,
 * fence.sys, release flag, acquire poll.  Both live in the same symmetric buffer (icsg3d_bn_allreduce_buffer_bytes covers
 * both regions) and give bit-identical sums. */
int icsg3d_set_peer_ll(int on);
/* Programmatic dependent launch (every kernel waits on its predecessor with griddepcontrol.wait and is launched with the
 * programmatic-stream-serialization attribute, so consecutive kernels of a stream overlap launch/prologue with the
 * predecessor's tail).  Off by default (no measurable gain on the captured train step: 3.173 vs 3.179 ms); 1 (or
 * ICSG3D_PDL=1) turns it on.  Captured graphs keep the mode they were captured with. */
int icsg3d_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------
 * Conv3D 3x3x3, stride 1, "same" — Keras Conv3D(kernel_size=(3,3,3), padding="same")
 *   lattice_vae.py:173,178,213,219-224; unet.py:276-336.
 * Implicit GEMM on tcgen05/TMEM: M = B*D*H*W output voxels (128 per tile), N = output channels,
 * K = 27 taps x Cin; the A operand is fetched tap by tap with 5-D tiled TMA whose out-of-bounds zero
 * fill implements the "same" padding; accumulators live in TMEM (double buffered).
 *
 * x      : bf16 [B,D,H,W,ldx], first `cin` channels are read; cin % 16 == 0, ldx % 8 == 0
 * wpack  : bf16 [27][nout][cin] (K-major GEMM B operand), produced by icsg3d_pack_conv_w_*.
 *          For fprop it holds W[tap][co][ci]; for dgrad W[26-tap][ci][co] (mirrored taps), so the
 *          same kernel computes either pass.
 * bias   : fp32 [nout] or NULL
 * y      : bf16 or fp32 [B,D,H,W,ldy]; columns [0,n_store) of the result are written (n_store<=nout)
 * act    : ICSG3D_ACT_* applied after the bias (U-Net ordering Conv->ReLU->BN, unet.py:276-278)
 * D,H,W must be powers of two (2..128) and B*D*H*W a multiple of 128 or smaller than one tile.
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_conv3d_k3_igemm(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                           int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                           float leaky_alpha, void* stream);

/* Same conv with a caller-owned fp32 workspace: layers with too few output tiles to fill the GPU (4^3 / 2^3 grids with a
 * deep K = 27*Cin) run with the K range split over several CTAs and a fixed-order reduction (deterministic).
 * icsg3d_conv3d_k3_workspace_bytes() = bytes needed for this shape (0: the layer is not split; ws may then be NULL). */
/* Fused BatchNorm statistics in the halo kernel (layers with Cout >= 64 at 16^3 / 8^3): implemented and tested, OFF by
 * default because its fp32 shared-memory atomics make a train step depend on warp arrival order in the last bits (the
 * streaming kernel's fused statistics are always on).  icsg3d_conv3d_k3_stats_parts() reports 0 for halo layers unless
 * enabled here or with ICSG3D_HALO_STATS=1. */
int icsg3d_conv3d_set_halo_stats(int on);
/* Autotuning hook (tools/halo_autotune.py): restrict the halo planner to one (TD, TH, NT) configuration; (0,0,0) = off. */
int icsg3d_conv3d_halo_force(int td, int th, int nt);
/* Same for the plane-streaming kernel: only plans with this number of h-blocks per plane (0 = off). */
int icsg3d_conv3d_stream_force(int n_hblk);
int64_t icsg3d_conv3d_k3_workspace_bytes(int B, int D, int H, int W, int cin, int nout);
int icsg3d_conv3d_k3_igemm_ws(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                              int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                              float leaky_alpha, void* ws, int64_t ws_bytes, void* stream);
/* Conv3D 3x3x3 "same" over concatenate([skip, UpSampling3D(2)(low)]) (unet.py:309-332: c13 / c15 / c17) without the upsampled
 * tensor (csrc/conv3d_upfold.cu): each of the 8 output phases is a 2x2x2 convolution of the LOW-resolution
 * tensor with summed weights (8 taps instead of 27) accumulated with the 27-tap convolution of the skip channels.
 *   x_skip bf16 [B,D,H,W,ld_skip] (c_skip channels), x_low bf16 [B,D/2,H/2,W/2,ld_low] (c_low channels), both multiples of 64;
 *   wfold: icsg3d_pack_conv_w_upfold of the Keras kernel fp32 [27][cin][cout] (icsg3d_conv3d_upfold_wpack_elems bf16 elements);
 *   y bf16 [B,D,H,W,ldy] = post_scale * act(conv + bias) + post_shift (post_* optional, as icsg3d_conv3d_k3_igemm_post). */
int64_t icsg3d_conv3d_upfold_wpack_elems(int c_skip, int c_up, int nout);
int icsg3d_pack_conv_w_upfold(const float* w, int cin, int cout, int c_skip0, int c_skip, int c_up0, int c_up, void* wfold,
                              void* stream);
int icsg3d_conv3d_k3_upfold(const void* x_skip, int ld_skip, int c_skip, const void* x_low, int ld_low, int c_low,
                            const void* wfold, const float* bias, const float* post_scale, const float* post_shift, void* y,
                            int ldy, int n_store, int B, int D, int H, int W, int nout, int act, float leaky_alpha,
                            void* stream);
/* Backward of the folded convolution w.r.t. its LOW-resolution input (the training step's use of the fold): dlow bf16
 * [B,D/2,H/2,W/2,ld_low] (c_up channels) from dy bf16 [B,D,H,W,ld_dy] (cout channels, multiple of 64) with the transposed
 * folded weights of icsg3d_pack_conv_w_upfold_dgrad (icsg3d_conv3d_upfold_dgrad_wpack_elems bf16 elements).  Equals
 * UpSampling3D's backward (sum over the 8 children) of the full-resolution data gradient of those channels. */
int64_t icsg3d_conv3d_upfold_dgrad_wpack_elems(int cout, int c_up);
int icsg3d_pack_conv_w_upfold_dgrad(const float* w, int cin, int cout, int c_up0, int c_up, void* wpack, void* stream);
int icsg3d_conv3d_k3_upfold_dgrad_low(const void* dy, int ld_dy, int cout, const void* wpack, void* dlow, int ld_low, int c_up,
                                      int B, int D, int H, int W, void* stream);
/* Inference form of Conv3D + activation + BatchNormalization (unet.py:277-279 in learning phase 0): the per-channel affine
 * of the moving statistics (icsg3d_bn_inference_coeffs) is applied in the conv epilogue,
 * y = post_scale[c] * act(conv + bias[c]) + post_shift[c], so the BatchNorm pass over the activation disappears. */
int icsg3d_conv3d_k3_igemm_post(const void* x, int ldx, const void* wpack, const float* bias, const float* post_scale,
                                const float* post_shift, void* y, int ldy, int y_dtype, int n_store, int B, int D, int H,
                                int W, int cin, int nout, int act, float leaky_alpha, void* ws, int64_t ws_bytes,
                                void* stream);

/* Conv3D + BiasAdd(+activation) that also emits the BatchNorm statistics of its own (stored, bf16-rounded) output from the
 * epilogue — replaces a separate icsg3d_bn_stats read pass for the BatchNormalization() that follows the conv
 * (lattice_vae.py:174,214; unet.py:278).  stats: fp64 [parts][2][nout] (sum, sum of squares per channel), the layout of
 * icsg3d_bn_stats partials; parts = icsg3d_conv3d_k3_stats_parts(...) for the same shape (0 = not served by the fused
 * path on this device/shape: use icsg3d_conv3d_k3_igemm + icsg3d_bn_stats). */
int icsg3d_conv3d_k3_stats_parts(int B, int D, int H, int W, int cin, int nout);
int icsg3d_conv3d_k3_igemm_stats(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                                 int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                                 float leaky_alpha, double* stats, int stats_parts, void* stream);

/* Diagnostic (host only): kernel/tiling choice of the dispatcher for a layer shape.
 * out[0] = impl: 0 per-tap TMA kernel; 1 halo-reuse kernel {1, TD, TH, G, NT, a_bufs, b_stages, items, kc, smem};
 * 2 plane-streaming kd-folded kernel {2, R, TH, T, C, stages, issuers, grid, kc, smem}. */
int icsg3d_conv3d_k3_plan(int B, int D, int H, int W, int cin, int nout, int sms, int* out);

/* Diagnostic: restrict the dispatcher (A/B timing of the kernel generations).  impl = 0 automatic (default),
 * 1 per-tap TMA kernel only, 2 halo-reuse kernel (no plane-streaming kernel).  Same as env ICSG3D_CONV_IMPL=v1|halo. */
int icsg3d_conv3d_set_impl(int impl);

/* Diagnostic: role timeline of the plane-streaming kernel.  While buf != NULL, CTA 1 of every streaming launch writes
 * clock64 stamps into buf[4][steps][4] (device int64): role 0 producer {TMA issued}, 1 first MMA issuer {enter, ring slot
 * free, input plane landed, MMAs issued}, 2/3 the two epilogue warps of lane quarter 0 {enter, accumulators complete,
 * items drained, slot released}.  buf = NULL switches it off (default).  tools/stream_timeline.py. */
int icsg3d_conv3d_stream_debug(void* buf, int steps);

/* Conv3D 1x1x1 (the U-Net heads `soft`/`sig`, unet.py:339-352) through the same tcgen05 kernel with a single tap:
 * wpack bf16 [1][nout][cin]. */
int icsg3d_conv3d_k1_igemm(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                           int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                           float leaky_alpha, void* stream);
int64_t icsg3d_conv3d_k1_wgrad_workspace(int B, int D, int H, int W, int cin, int cout);
int icsg3d_conv3d_k1_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D, int H,
                           int W, int cin, int cout, void* workspace, int64_t workspace_bytes, void* stream);

/* Conv3DBackpropFilterV2: dW[tap][ci][co] = sum_voxels x[voxel+tap, ci] * dy[voxel, co]
 * (gradient of the layers above w.r.t. their kernels).  tcgen05 GEMM with both operands MN-major,
 * the voxel range split over CTAs; partials are reduced in a fixed order (deterministic).
 * x : bf16 [B,D,H,W,ldx] (cin % 16 == 0), dy : bf16 [B,D,H,W,ldy] (cout % 16 == 0)
 * dw : fp32 [27][cin][cout];  workspace: icsg3d_conv3d_k3_wgrad_workspace() bytes. */
int64_t icsg3d_conv3d_k3_wgrad_workspace(int B, int D, int H, int W, int cin, int cout);
int icsg3d_conv3d_k3_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D, int H,
                           int W, int cin, int cout, void* workspace, int64_t workspace_bytes, void* stream);

/* Plain CUDA-core direct convolutions, fp32 accumulate.  On-device cross-checks for the tcgen05
 * kernels (tests only; never on the training path). */
int icsg3d_ref_conv3d_k3(const void* x, int ldx, const void* wpack, const float* bias, float* y, int ldy,
                         int B, int D, int H, int W, int cin, int nout, int act, float leaky_alpha,
                         void* stream);
int icsg3d_ref_conv3d_k3_wgrad(const void* x, int ldx, const void* dy, int ldy, float* dw, int B, int D,
                               int H, int W, int cin, int cout, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight packing.  Master weights are fp32 in Keras layout (kd,kh,kw,Cin,Cout).
 *   fprop : wpack[tap][co][ci_pad]   = W[tap][ci][co]                (zero for padded rows/cols)
 *   dgrad : wpack[tap][ci][co_pad]   = W[26-tap][ci][co]
 * `fold` > 1 folds the encoder's tiled condition channels (lattice_vae.py:167-169, K.tile -> 4 replicas
 * of the 10-way one-hot): source channel cin_lead + r*fold_c + k (r < fold) is summed into packed
 * channel cin_lead + k.  With fold <= 1 the mapping is the identity.
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_pack_conv_w_fprop(const float* w, void* wpack, int cin, int cout, int cin_pad, int cout_pad,
                             int cin_lead, int fold, int fold_c, void* stream);
int icsg3d_pack_conv_w_dgrad(const float* w, void* wpack, int cin, int cout, int cin_pad, int cout_pad,
                             void* stream);
/* Same for the input-channel slice [ci0, ci0 + cin) of a kernel with cin_total input channels: bf16 [27][cin][cout]. */
int icsg3d_pack_conv_w_dgrad_slice(const float* w, void* wpack, int cin_total, int ci0, int cin, int cout, void* stream);

/* All weight packs of a model in one launch.  jobs: DEVICE array [njobs][10] of int64
 * {w ptr, wpack ptr, cin, cout, cin_pad, cout_pad, cin_lead, fold, fold_c, mode (0 = fprop layout, 1 = dgrad layout,
 * 2 / 3 = fprop / dgrad layout as bf16-pair split operands [w_hi | w_hi | w_lo] along K, 4 = the lean 32-channel split pack of
 * the encoder's first conv, see the fp32-class section)};
 * max_blocks = grid.x (each job strides over its own element count). */
int icsg3d_pack_conv_w_batch(const int64_t* jobs, int njobs, int max_blocks, void* stream);
/* dW (padded, from wgrad) -> gradient in Keras layout, undoing padding and the condition fold. */
int icsg3d_unpack_conv_dw(const float* dw_pad, float* dw, int cin, int cout, int cin_pad, int cout_pad,
                          int cin_lead, int fold, int fold_c, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNormalization in batch-statistics mode, fused with the activation and the MaxPool3D(2) /
 * UpSampling3D(2) that follow it — keras BatchNormalization() (lattice_vae.py:174,214,225;
 * unet.py:278-338; eps 1e-3, momentum 0.99 — SURVEY R3), LeakyReLU/ReLU, MaxPool3D, UpSampling3D.
 * Tensors are [B,D,H,W,ld] with C used channels, dtype bf16 (C % 8 == 0) or fp32 (C % 4 == 0).
 *
 * Forward:   bn_stats (per-block fp64 partials of sum x, sum x^2)  ->  bn_reduce_partials -> sums[2][C]
 *            [data parallel: all-reduce `sums` here]  ->  bn_finalize (mean, rstd, scale = gamma*rstd,
 *            shift = beta - mean*scale, moving-average update)  ->  bn_apply_fwd.
 * Backward:  bn_bwd_reduce (partials of sum g, sum g*xhat, g = act'(.) * un-post(dy)) -> bn_reduce_partials
 *            [all-reduce] -> bn_bwd_apply (dx = scale*(g - mean(g) - xhat*mean(g*xhat))), bn_param_grads.
 * All reductions are two-stage with a fixed summation order (deterministic).
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_bn_nparts(int64_t rows, int C, int dtype);
int icsg3d_bn_stats(const void* x, int ldx, int dtype, int64_t rows, int C, double* partials, int nparts,
                    void* stream);
int icsg3d_bn_reduce_partials(const double* partials, int nparts, int C, double* sums, void* stream);
int icsg3d_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps,
                       float* mean, float* rstd, float* scale, float* shift, float* moving_mean,
                       float* moving_var, float momentum, int C, void* stream);
/* Single-process fusions of the two steps above (same fixed summation order, one launch):
 * bn_reduce_finalize = bn_reduce_partials + bn_finalize;  bn_reduce_grads = bn_reduce_partials + bn_param_grads. */
int icsg3d_bn_reduce_finalize(const double* partials, int nparts, double count, const float* gamma, const float* beta,
                              float eps, double* sums, float* mean, float* rstd, float* scale, float* shift,
                              float* moving_mean, float* moving_var, float momentum, int C, void* stream);
int icsg3d_bn_reduce_grads(const double* partials, int nparts, int C, double* sums, float* dgamma, float* dbeta,
                           void* stream);
/* Data-parallel fusions (sync-BN statistics): partial reduction + all-reduce of the 2C sums over NVLink
 * PEER MEMORY + finalisation in one kernel, instead of bn_reduce_partials -> NCCL all-reduce -> bn_finalize.
 * `peers`: DEVICE array [world] with the base address of every rank's symmetric buffer as mapped into this process
 * (torch symmetric memory), each icsg3d_bn_allreduce_buffer_bytes(world, nslots, cmax) bytes, zeroed once;
 * `slot`: a distinct id per call site (layer x forward/backward), the same on every rank; `epoch`: DEVICE int64 >= 1 that
 * the caller increments once per step (read on the device, so the launch is CUDA-graph replayable); count_global = rows
 * summed over all ranks.  _finalize writes mean/rstd/scale/shift (+moving averages) from the GLOBAL statistics; _grads
 * writes the global sums (sum g, sum g*xhat) and the LOCAL dbeta/dgamma (the gradient all-reduce adds the ranks).
 * Deterministic: every rank adds the per-rank sums in rank order.  Every rank must launch the same sequence. */
int64_t icsg3d_bn_allreduce_buffer_bytes(int world, int nslots, int cmax);
int icsg3d_bn_reduce_allreduce_finalize(const double* partials, int nparts, double count_global, const float* gamma,
                                        const float* beta, float eps, double* sums, float* mean, float* rstd, float* scale,
                                        float* shift, float* moving_mean, float* moving_var, float momentum, int C,
                                        const uint64_t* peers, int world, int rank, int slot, int nslots, int cmax,
                                        const int64_t* epoch, void* stream);
int icsg3d_bn_reduce_allreduce_grads(const double* partials, int nparts, int C, double* sums_global, float* dgamma,
                                     float* dbeta, const uint64_t* peers, int world, int rank, int slot, int nslots, int cmax,
                                     const int64_t* epoch, void* stream);
/* The whole BatchNorm backward of one layer (bn_bwd_reduce + bn_reduce_grads / bn_reduce_allreduce_grads + bn_bwd_apply)
 * in ONE cooperative launch with two grid-wide barriers: two launches fewer and the second read of dy / x comes from L2
 * for the layers that fit.  partials: scratch of icsg3d_bn_bwd_fused_nparts() x 2C doubles; sums [2][C] receives the
 * (global) sums; dgamma/dbeta the LOCAL sums.  peers == NULL: single device; otherwise the peer-memory exchange described
 * above runs between the two barriers (count_global = rows of all ranks). */
int icsg3d_bn_bwd_fused_nparts(int C, int dtype);
int icsg3d_bn_bwd_fused(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype,
                        const float* mean, const float* rstd, const float* scale, const float* shift, int act, float alpha,
                        int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C, double* partials, double* sums,
                        double count_global, float* dgamma, float* dbeta, int pre_relu, const void* tap_other, int ld_other,
                        float tap_coef, void* dx, int lddx, const uint64_t* peers, int world, int rank, int slot, int nslots,
                        int cmax, const int64_t* epoch, void* stream);
/* ---- "fp32-class" operand mode (north_star: activations 1e-4, bit-exact argmax) ---------------------------------------
 * Conv operands are carried as bf16 pairs hi = bf16(v), lo = bf16(v - hi); the ordinary bf16 conv kernels (fp32 output)
 * compute x_hi*w_hi + x_lo*w_hi + x_hi*w_lo when the activation tensor stores [hi | lo | hi] (3x channels) and the
 * packed weights [w_hi | w_hi | w_lo] along Cin.  Producers of the split layouts (csrc/split3.cu, bn.cu): */
/* fmt: 0 = bf16 pairs (16 significant bits), 1 = IEEE fp16 pairs (22 bits: fp32 class; the 2-byte values live in the
 * same buffers and the conv is told through icsg3d_conv3d_k3_igemm_f16 / _k1_igemm_f16 below). */
int icsg3d_f32_to_split3(const float* src, int ld_src, int c, int64_t rows, void* dst, int ctot, int coff, int fmt,
                         void* stream);
int icsg3d_pack_vae_input_split3(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe, void* xp,
                                 int fmt, void* stream);
/* The train step's input pack when only the ENCODER runs on split operands: xe3 bf16 [B*vox][48] = [hi | lo | hi] of
 * (M, one-hot cond, 0) and xp16 bf16 [B*vox][16] = (M, 0..) for the bf16 perceptual model, one read of m. */
int icsg3d_pack_vae_input_mixed(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe3, void* xp16,
                                int fmt, void* stream);
/* Lean form of the same for the encoder (bf16 pairs): xe32 bf16 [B*vox][32] = [M_hi(4) | cond(10) | M_lo(0:2)]
 * [M_lo(2:4) | M_hi(4) | cond(10)] — the one-hot condition is exact in bf16 and carries no lo part; pairs with pack mode 4 of
 * icsg3d_pack_conv_w_batch ([27][cout_pad][32] = [w_hi(M) | w_hi(cond) | w_hi(M 0:2)] [w_hi(M 2:4) | w_lo(M) | w_lo(cond)]). */
int icsg3d_pack_vae_input_lean(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe32, void* xp16,
                               void* stream);
int icsg3d_pack_conv_w_fprop_x3(const float* w, void* wpack, int ntaps, int cin, int cout, int cin_pad, int cout_pad,
                                int cin_lead, int fold, int fold_c, int fmt, float wscale, void* stream);
/* The conv dispatcher on fp16 operands (same kernels and layouts; only the tcgen05 operand-format fields differ).
 * y = accumulator * out_scale + bias: weights packed with wscale = 2^k (lo parts stay normal fp16) use out_scale = 2^-k. */
/* Backward in the split form.  dgrad: the ordinary conv on dy stored [hi | lo | hi] with this pack of the kernel,
 * bf16 [ntaps][cin_pad][3*cout_pad] = [w_hi | w_hi | w_lo] along Cout, taps mirrored.  wgrad: the ordinary filter-gradient
 * kernel on x = [x_hi | x_lo] and dy = [dy_hi | dy_lo] (channel slices of the split tensors) gives P fp32
 * [ntaps][2*cin_pad][2*cout_pad]; wgrad_combine_x3 sums the three blocks hi*hi + lo*hi + hi*lo into dw [ntaps][cin_pad][cout_pad]. */
int icsg3d_pack_conv_w_dgrad_x3(const float* w, void* wpack, int ntaps, int cin, int cout, int cin_pad, int cout_pad, int fmt,
                                float wscale, void* stream);
int icsg3d_wgrad_combine_x3(const float* P, float* dw, int ntaps, int cin_pad, int cout_pad, void* stream);
/* fp32 forms of the backward glue for the split (fp32-class) training step: BatchNorm backward apply with an fp32 dx
 * (x, dy, dy2, tap_other all fp32), the loss-gradient seeds, the activation backward of enc_conv5, and the head losses
 * with an fp32 d(loss)/d(logits). */
int icsg3d_bn_bwd_apply_f32(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, const float* mean,
                            const float* rstd, const float* scale, const float* shift, int act, float alpha, int post,
                            const uint8_t* pool_idx, int B, int D, int H, int W, int C, const double* sums, double count,
                            int pre_relu, const void* tap_other, int ld_other, float tap_coef, float* dx, int lddx,
                            void* stream);
int icsg3d_xhat_grad_f32(const float* x, const float* xhat, float mse_coef, const float* dpm, int ld, int64_t rows, float* dy,
                         void* stream);
int icsg3d_tap_grad_relu_f32(const float* a, const float* other, float coef, int64_t n, float* dc, void* stream);
int icsg3d_act_bwd_f32(const float* dy, const float* y, int act, float alpha, int64_t n, float* dx, void* stream);
int icsg3d_heads_loss_f32grad(const float* logits, int ld, int c1, const uint8_t* species, const float* class_w, int64_t M,
                              float inv_count, uint8_t* argmax_out, float* sig_prob, float* probs, float* dlogits, int ldd,
                              double* partials, int nparts, void* stream);
int icsg3d_conv3d_k3_igemm_f16(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                               int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                               float leaky_alpha, float out_scale, void* stream);
int icsg3d_conv3d_k1_igemm_f16(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy,
                               int y_dtype, int n_store, int B, int D, int H, int W, int cin, int nout, int act,
                               float leaky_alpha, float out_scale, void* stream);
/* BatchNorm apply (+activation, +pool/upsample) of an fp32 conv output straight into the split tensor y bf16
 * [rows][3*ctot]; this layer's C channels start at `coff` inside every part (skip concatenations write side by side). */
int icsg3d_bn_apply_fwd_split3(const void* x, int ldx, int x_dtype, const float* scale, const float* shift, int act,
                               float alpha, int post, int B, int D, int H, int W, int C, void* y, int ldy,
                               uint8_t* pool_idx, int ctot, int coff, int fmt, void* stream);
/* learning phase 0 (predict / test_on_batch): scale/shift from the moving statistics */
int icsg3d_bn_inference_coeffs(const float* gamma, const float* beta, const float* moving_mean,
                               const float* moving_var, float eps, float* scale, float* shift, int C,
                               void* stream);
/* y = post(act(scale*x + shift)); post POOL2 also writes the uint8 index of the FIRST maximum of every
 * 2x2x2 window; y32 (fp32 copy, ld ldy32) is optional and only valid with POST_NONE. */
int icsg3d_bn_apply_fwd(const void* x, int ldx, int x_dtype, const float* scale, const float* shift, int act,
                        float alpha, int post, int B, int D, int H, int W, int C, void* y, int ldy, float* y32,
                        int ldy32, uint8_t* pool_idx, void* stream);
int icsg3d_bn_bwd_nparts(int B, int D, int H, int W, int C, int dtype, int post);
/* dy2 (optional, may be NULL): a second gradient w.r.t. the UN-pooled BN output with x's shape — the U-Net skip
 * connections feed both MaxPool3D and a later Concatenate (unet.py:282,332). */
int icsg3d_bn_bwd_reduce(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype,
                         const float* mean,
                         const float* rstd, const float* scale, const float* shift, int act, float alpha,
                         int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C, double* partials,
                         int nparts, void* stream);
/* pre_relu: U-Net ordering Conv->ReLU->BN (unet.py:276-278): x is the ReLU output and dx is masked by x>0.
 * tap_other/tap_coef: DFC feature-loss gradient tap_coef*(x - tap_other) added before the mask
 * (lattice_vae.py:257-270).  dx is bf16 with stride lddx. */
int icsg3d_bn_bwd_apply(const void* dy, int lddy, const void* dy2, int lddy2, const void* x, int ldx, int dtype,
                        const float* mean,
                        const float* rstd, const float* scale, const float* shift, int act, float alpha,
                        int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C, const double* sums,
                        double count, int pre_relu, const void* tap_other, int ld_other, float tap_coef,
                        void* dx, int lddx, void* stream);
/* bn_bwd_apply of a tapped layer that also returns the layer's DFC feature loss (lattice_vae.py:257-270):
 * tap_sq[b] = this block's sum (x - tap_other)^2 (fp64), tap_sq_nparts = icsg3d_bn_bwd_apply_nblocks(...) partials that the
 * caller sums (icsg3d_vae_loss_assemble): the separate icsg3d_sqdiff_partials pass over the two feature maps is not needed. */
int icsg3d_bn_bwd_apply_nblocks(int B, int D, int H, int W, int C, int dtype, int post);
int icsg3d_bn_bwd_apply_tapsq(const void* dy, int lddy, const void* x, int ldx, int dtype, const float* mean,
                              const float* rstd, const float* scale, const float* shift, int act, float alpha,
                              int post, const uint8_t* pool_idx, int B, int D, int H, int W, int C, const double* sums,
                              double count, int pre_relu, const void* tap_other, int ld_other, float tap_coef,
                              void* dx, int lddx, double* tap_sq, int tap_sq_nparts, void* stream);
int icsg3d_bn_param_grads(const double* sums, float* dgamma, float* dbeta, int C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * VAE glue: input packing, bottleneck Dense / sampling / KL, losses, loss-gradient seeds, Adam.
 * ---------------------------------------------------------------------------------------------- */
/* m fp32 [B*vox][4], cond fp32 [B][ncond] -> xe bf16 [B*vox][16] = (M, one-hot cond, 0) (encoder input:
 * Reshape+K.tile+Concatenate of lattice_vae.py:167-169 with the 4 replicas folded into the weights) and
 * xp bf16 [B*vox][16] = (M, 0...) (perceptual U-Net input).  Either output may be NULL. */
int icsg3d_pack_vae_input(const float* m, const float* cond, int ncond, int B, int64_t vox, void* xe, void* xp,
                          void* stream);
int icsg3d_f32_to_bf16_rows(const float* src, int c, int64_t rows, void* dst, int ld, void* stream);
int icsg3d_bf16_rows_to_f32(const void* src, int ld, int c, int64_t rows, float* dst, void* stream);
/* Dense (+ preceding Concatenate of two inputs): y = act([x1|x2] W + b); W fp32 (k1+k2, N) Keras layout
 * (lattice_vae.py:183-185, 207-208). */
int icsg3d_dense_fwd(const float* x1, int k1, const float* x2, int k2, const float* w, const float* bias,
                     int act, int B, int N, float* y, void* stream);
/* dy is masked in place by relu'(y) when act == RELU; dx1 (w.r.t. x1, may be NULL), dw (k1+k2,N), db (N). */
int icsg3d_dense_bwd(const float* x1, int k1, const float* x2, int k2, const float* w, const float* y, int act,
                     float* dy, int B, int N, float* dx1, int accumulate_dx, float* dw, float* db, void* stream);
/* sampling (lattice_vae.py:53-66) with an explicit eps, and the per-sample KL term (lattice_vae.py:235-239) */
int icsg3d_reparam_fwd(const float* mu, const float* lv, const float* eps, int B, int L, float* z, float* kl,
                       void* stream);
int icsg3d_reparam_bwd(const float* dz, const float* mu, const float* lv, const float* eps, float kl_coef, int B,
                       int L, float* dmu, float* dlv, void* stream);
int icsg3d_leaky_bwd_rows(const float* dy, const float* y, float alpha, int c, int64_t rows, void* dst, int ld,
                          void* stream);
/* sum (a-b)^2 as per-block fp64 partials (MSE and DFC feature losses, lattice_vae.py:232-233,266-269) */
int icsg3d_sqdiff_nparts(int64_t n);
int icsg3d_sqdiff_partials(const void* a, const void* b, int dtype, int64_t n, double* partials, int nparts,
                           void* stream);
/* out = [loss, pm, mse, kld] exactly as train_on_batch reports them (lattice_vae.py:124-125, 241-255) */
int icsg3d_vae_loss_assemble(const double* partials, const int* nparts, int stride, int nterms,
                             const double* scales, const float* kl, int B, double kl_scale, float alpha,
                             float beta, float* out, double* raw, void* stream);
int icsg3d_xhat_grad(const float* x, const float* xhat, float mse_coef, const void* dpm, int ld, int64_t rows,
                     float* dy, void* stream);
int icsg3d_tap_grad_relu(const void* a, const void* other, float coef, int64_t n, void* dc, void* stream);
int icsg3d_bias_grad(const void* dy, int ld, int64_t rows, int C, float* db, void* stream);
/* keras.optimizers.Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
 * state: double[2] on the device = {t, lr_t}; advanced on the device so the step is CUDA-graph replayable. */
int icsg3d_adam_keras_step(float* p, const float* g, float* m, float* v, double* state, double lr, double beta1,
                           double beta2, double eps, float grad_scale, int64_t n, void* stream);

/* Data parallel: all-reduce of the flat gradient over NVLink PEER MEMORY fused with the Adam update — one
 * kernel instead of NCCL all-reduce + icsg3d_adam_keras_step.  `peers`: DEVICE array [world] of every rank's symmetric
 * buffer (icsg3d_adam_allreduce_buffer_bytes(world, n) bytes each, zeroed once); `epoch`: DEVICE int64 >= 1, incremented by
 * the caller once per step.  On return (stream order) g holds the GLOBAL (rank-ordered, bit-identical on all ranks)
 * gradient sum and p/m/v are updated with it.  Every rank must launch it once per step. */
int64_t icsg3d_adam_allreduce_buffer_bytes(int world, int64_t n);
int icsg3d_adam_keras_allreduce_step(float* p, float* g, float* m, float* v, double* state, double lr, double beta1,
                                     double beta2, double eps, float grad_scale, int64_t n, const uint64_t* peers, int world,
                                     int rank, const int64_t* epoch, void* stream);

/* ------------------------------------------------------------------------------------------------
 * U-Net heads (unet.py:339-352) and their losses/metrics (unet.py:159-221, 249-259).
 * The two 1x1x1 convolutions (95-way softmax + 1 sigmoid) are ONE GEMM with nout = 96 columns
 * (icsg3d_conv3d_k1_igemm); pack_heads_w builds its operands from the Keras kernels; heads_loss is the fused
 * per-voxel pass over the fp32 logits [M][ld]: weighted CCE, logits-form BCE, argmax label, sigmoid
 * probability, optional softmax probabilities (fp32 [M][c1], what model.predict returns), metric counts, and
 * d(loss)/d(logits) (bf16 [M][ldd], already scaled by inv_count).
 * partials: fp64 [nparts][6] = {soft loss, sig loss, tp, predicted, tp_w, possible_w} sums.
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_pack_heads_w(const float* w_soft, const float* w_sig, const float* b_soft, const float* b_sig, int cin,
                        int c1, int nout, void* wf, void* wd, float* bias, void* stream);
int icsg3d_unpack_heads_grad(const float* dwcat, const double* colsum, int cin, int c1, int nout, float* dw_soft,
                             float* dw_sig, float* db_soft, float* db_sig, void* stream);
int icsg3d_heads_loss_nparts(int64_t M);
int icsg3d_heads_loss(const float* logits, int ld, int c1, const uint8_t* species, const float* class_w, int64_t M,
                      float inv_count, uint8_t* argmax_out, float* sig_prob, float* probs, void* dlogits, int ldd,
                      double* partials, int nparts, void* stream);
/* out = [loss, soft_loss, sig_loss, f1_m, wr_m]; raw (optional) = the six term sums (for data-parallel reduction) */
int icsg3d_heads_loss_finalize(const double* partials, int nparts, double count, float* out, double* raw,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device post-processing of the generate.py loop (generate.py:208-225) and the augmentation of create_matrices.py
 * (csrc/post.cu).
 *
 * to_lattice_params / to_voxel_params (utils.py:160-190): per-sample min/max over the three coordinate channels
 * [c0, c0+3) of p ([B][vox][ld], fp32 or fp64), in two launches:
 *   coord_minmax    -> partials [B][nsplit][6] (same dtype; nsplit = icsg3d_lattice_nsplit(B, vox)); with x16 != NULL
 *                      (fp32, ld == 4, c0 == 1: the decoder output) it ALSO writes the bf16 [B][vox][16] U-Net input, so
 *                      the decoder output is read once on its way into the segmentation network;
 *   lattice_finalize -> lp [B][3] = ((max-min)/(1+2 eps)/(1-1/d)) * (1-1/d as `ap -= ap/d`: the reference's quirk),
 *                       dv [B][3] = (lp + 2 lp eps)/d (optional), evaluated op by op in the array dtype with
 *                       round-to-nearest intrinsics => bit-identical to numpy on the same min/max.
 * heads_predict (generate.py:221-225): argmax species label (first index on ties) and the sigmoid >= threshold atom mask
 *   from the fp32 head logits [M][ld] (c1 soft columns + the sigmoid logit in column c1); every output is optional.
 * rotate90_batch (utils.py:193-222 random_rotation_3d, create_matrices.py:174-207): exact signed axis permutation of a
 *   batch of d^3 grids with `voxel_bytes` bytes per voxel; xforms int32 [B][6] = (perm0,perm1,perm2, flip0,flip1,flip2):
 *   out[b][o0][o1][o2] = in[b][s0][s1][s2], s_x = flip_x ? d-1-o[perm_x] : o[perm_x].  in != out.
 * metric_counts (unet.py:159-193 r_m / p_m / f1_m / wr_m): counts fp64 [5] = { sum round(clip(yt*yp,0,1)),
 *   sum round(clip(yt)), sum round(clip(yp)), the first two again without class 0 } over fp32 [n/C][C] tensors.
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_lattice_nsplit(int B, int64_t vox);
int icsg3d_coord_minmax(const void* p, int dtype, int ld, int c0, int B, int64_t vox, int nsplit, void* partials, void* x16,
                        void* stream);
int icsg3d_lattice_finalize(const void* partials, int dtype, int B, int nsplit, double eps_frac, int d, void* lp, void* dv,
                            void* stream);
int icsg3d_heads_predict(const float* logits, int ld, int c1, int64_t M, float threshold, uint8_t* argmax_out,
                         uint8_t* mask_out, float* sig_prob, void* stream);
/* generate.py:220-225 in ONE kernel (csrc/heads_fused.cu): the two 1x1x1 head convolutions (unet.py:338-341) on the
 * features x [M][ldx] (2-byte operands: bf16, or IEEE fp16 with op_f16 = 1 for the fp32-class split mode) with the packed
 * head weights of icsg3d_pack_heads_w ([nout][cin], column c1 = sigmoid head), then exactly icsg3d_heads_predict on the
 * logits = accumulator * out_scale + bias — which stay in tensor memory and are never written to HBM.
 * cin: multiple of 64, <= 384; nout: multiple of 16, <= 96; 1 <= c1 < nout. */
int icsg3d_heads_predict_fused(const void* x, int ldx, const void* wpack, const float* bias, int64_t M, int cin, int nout,
                               int c1, int op_f16, float out_scale, float threshold, uint8_t* argmax_out,
                               uint8_t* mask_out, float* sig_prob, void* stream);
/* Training form of the same kernel (unet.py:196-221, 252-259 in one pass over the features): head GEMM, weighted
 * categorical cross-entropy on the clipped soft-max + binary cross-entropy on the sigmoid head, the f1 / weighted-recall
 * counts (unet.py:159-193) and d(loss)/d(logits) as bf16 [M][ldd] — what icsg3d_conv3d_k1_igemm + icsg3d_heads_loss
 * compute through 384 B/voxel of fp32 logits in HBM.  partials: fp64 [icsg3d_heads_loss_fused_nparts(M)][6] for
 * icsg3d_heads_loss_finalize.  species uint8 [M] (binary target = species != 0), class_w fp32 [c1] or null. */
int icsg3d_heads_loss_fused_nparts(int64_t M);
int icsg3d_heads_loss_fused(const void* x, int ldx, const void* wpack, const float* bias, int64_t M, int cin, int nout, int c1,
                            const uint8_t* species, const float* class_w, float inv_count, double* partials,
                            uint8_t* argmax_out, float* sig_prob, void* dlogits, int ldd, void* stream);
int icsg3d_rotate90_batch(const void* in, void* out, int B, int d, int voxel_bytes, const int* xforms, void* stream);
int icsg3d_metric_counts(const float* y_true, const float* y_pred, int64_t n, int C, double* counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gaussian atomic-density voxeliser — utils.py:97-144 density_matrix + utils.py:88-94 coordinate_grid
 * (create_matrices.py:142-155).  sites: fp64 [ncells][max_sites][8] records
 * (x, y, z Cartesian, thr = sigma*label_frac, zs = Z/sigma^3, two_s2 = 2*sigma^2, Z, unused);
 * nsites int32 [ncells]; lattice fp64 [ncells][3] = (a,b,c).  Outputs (each optional):
 *   m32       fp32  [ncells][d][d][d][4]  network input (density, p_x, p_y, p_z)
 *   m64       fp64  [ncells][d][d][d]     density M exactly as the reference returns it
 *   species   uint8 [ncells][d][d][d]     species grid S (bit-exact; fp64 predicate, no FMA contraction)
 *   species64 fp64  [ncells][d][d][d]     S in the reference's dtype
 * ---------------------------------------------------------------------------------------------- */
int icsg3d_voxelize(const double* sites, const int* nsites, const double* lattice, int ncells, int max_sites,
                    int d, double eps_frac, float* m32, double* m64, uint8_t* species, double* species64,
                    void* stream);
/* On-device synthetic perovskite-like ABX3 cells (benchmark inputs); 5 sites per cell. */
int icsg3d_synth_perovskite_sites(uint64_t seed, int ncells, int max_sites, double label_frac, double* sites,
                                  int* nsites, double* lattice, void* stream);

/* Hardware probe (tools/tests only): UMMA K-major swizzled descriptors with row-shifted start addresses.
 * out: fp32 [2][nshift][128][n]; see csrc/probe.cu. */
int icsg3d_probe_shifted_desc(const void* a, const void* b, float* out, int rows, int kc, int n, int nshift,
                              void* stream);

/* Hardware probe: MN-major swizzled operands with overlapping MN blocks (leading byte offset = a_shift / b_shift rows):
 * out fp32 [128][nblk_b*cb]; see csrc/probe.cu "Probe 4". */
int icsg3d_probe_mn_fold(const void* x, const void* y, float* out, int rows, int ca, int cb, int a_shift, int nblk_b,
                         int b_shift, int ksteps, void* stream);

/* Hardware probe: cycles for `reps` back-to-back tcgen05.mma (SS, bf16, K=16) of shape (m,n); out[0] = issue
 * cycles, out[1] = cycles until the commit barrier fires. */
int icsg3d_probe_mma_rate(int64_t* out, int m, int n, int reps, int nacc, int swizzle_bytes, int a_step, int b_step,
                          void* stream);
int icsg3d_probe_mma_rate_mn(int64_t* out, int n, int reps, int nacc, int mn, void* stream);

/* Hardware probe: the halo kernel's MMA issue pattern without TMA/barriers/epilogue (see csrc/probe.cu). */
int icsg3d_probe_halo_pattern(int64_t* out, int G, int nt, int plane_rows, int WP, int row_bytes, int ksteps, int items,
                              int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICSG3D_H_ */
