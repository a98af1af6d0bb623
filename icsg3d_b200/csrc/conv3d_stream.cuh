// Shared declarations of the plane-streaming, kd-folded Conv3D kernel (conv3d_stream.cu), used by the dispatcher
// in conv3d_igemm.cu.
#pragma once
#include "common.cuh"

namespace icsg3d {

struct ConvStreamParams {
  int B, D, H, W;
  int TH, HP, WP, n_hblk;
  int T, R, C;           // M tiles per plane slab, accumulator ring slots (output planes, power of two), channels per kd block (= nout)
  int tiles_n;           // N split: blockIdx.y owns output channels [y*C, (y+1)*C) with its own resident weight slice
  int r_log2, st_log2;   // log2(R), log2(stages): ring arithmetic is masks and shifts in the issue loop
  int kc, chunks, row_bytes;
  int stages;            // input-plane ring depth in shared memory (power of two)
  int issuers;           // MMA issuer warps in use (tiles are dealt round-robin)
  int total_steps, steps_per_cta;
  uint32_t a_chunk_bytes, a_stage_bytes, a_tx_bytes;
  uint32_t w_bytes, blk_bytes;
  uint32_t sbo, layout, idesc[3], tmem_cols;
  void* y;
  int ldy, y_dtype, n_store;
  const float* bias;
  int act;
  float alpha;
  float oscale;  // output = accumulator * oscale + bias (1 unless the weights were pre-scaled: fp16 split mode)
  double* stats;         // optional per-CTA BatchNorm partials [grid][2][C] (sum, sum of squares of the stored values)
  const float* post_scale;  // optional per-channel affine after the activation (inference BatchNorm); excludes stats
  const float* post_shift;
  long long* dbg;        // optional role timeline of CTA (0, 0): [4 roles][dbg_steps][4] clock64 stamps (tools/stream_timeline.py)
  int dbg_steps;
  int exp_flags;         // TIMING EXPERIMENTS ONLY (env ICSG3D_STREAM_EXP, results are wrong when set): 1 = no slot zeroing,
                         // 2 = no TMEM loads in the epilogue, 4 = no global stores
};

// false when the layer shape does not fit the streaming scheme (the dispatcher then falls back to the other kernels).
bool conv_stream_plan(int B, int D, int H, int W, int cin, int nout, int sms, ConvStreamParams* out);
int conv_stream_grid(const ConvStreamParams& p);
void conv_stream_force_nb(int nb);  // autotuning hook: only this number of h-blocks (0: off)
void conv_stream_set_debug(long long* buf, int steps);  // role timeline of CTA 1 into buf[4][steps][4] (nullptr: off)
int launch_conv_stream(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy, int y_dtype,
                       int n_store, int cin, int nout, int act, float alpha, double* stats, ConvStreamParams p,
                       cudaStream_t st, float oscale = 1.0f, const float* post_scale = nullptr,
                       const float* post_shift = nullptr);

}  // namespace icsg3d
