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

const char* icsg3d_last_error(void);
int icsg3d_version(void);
/* Number of SMs of the current device (148 on B200); <0 on error. */
int icsg3d_sm_count(void);

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
/* dW (padded, from wgrad) -> gradient in Keras layout, undoing padding and the condition fold. */
int icsg3d_unpack_conv_dw(const float* dw_pad, float* dw, int cin, int cout, int cin_pad, int cout_pad,
                          int cin_lead, int fold, int fold_c, void* stream);

/* Hardware probe (tools/tests only): UMMA K-major swizzled descriptors with row-shifted start addresses.
 * out: fp32 [2][nshift][128][n]; see csrc/probe.cu. */
int icsg3d_probe_shifted_desc(const void* a, const void* b, float* out, int rows, int kc, int n, int nshift,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICSG3D_H_ */
