// Shared declarations of the halo-reuse Conv3D kernel (conv3d_halo.cu), used by the dispatcher in conv3d_igemm.cu.
#pragma once
#include "common.cuh"

namespace icsg3d {

struct ConvHaloParams {
  int B, D, H, W;
  int TD, TH, HP, WP;
  int plane_rows, block_rows, out_rows, G;
  int kc, chunks, row_bytes;
  int nt, tiles_n;
  int n_dblk, n_hblk, total_items;
  int a_bufs, b_stages;
  int b_resident;  // all 27*chunks weight tiles stay in shared memory for the whole kernel (loaded once)
  uint32_t a_chunk_bytes, a_buf_bytes, a_tx_bytes, b_unit_bytes;
  uint32_t sbo, layout, idesc, tmem_cols;
  void* y;
  int ldy, y_dtype, n_store;
  const float* bias;
  int act;
  float alpha;
  float oscale;  // output = accumulator * oscale + bias (1 unless the weights were pre-scaled: fp16 split mode)
  double* stats;  // optional [grid][2][nt*tiles_n]: per-CTA BatchNorm partials (sum, sum of squares) of the STORED output
  const float* post_scale;  // optional per-channel affine after the activation (inference BatchNorm); excludes stats
  const float* post_shift;
};

void conv_halo_force(int td, int th, int nt);  // autotuning hook: consider only this configuration (0,0,0: off)
// Pick (TD, TH, NT) from a simple cycle model; false when the layer shape does not fit the halo scheme.
bool conv_halo_plan(int B, int D, int H, int W, int cin, int nout, int sms, ConvHaloParams* out);
int launch_conv_halo(const void* x, int ldx, const void* wpack, const float* bias, void* y, int ldy, int y_dtype,
                     int n_store, int cin, int nout, int act, float alpha, ConvHaloParams p, int sms, cudaStream_t st,
                     float oscale = 1.0f, double* stats = nullptr, const float* post_scale = nullptr,
                     const float* post_shift = nullptr);
// CTAs the halo kernel launches for this plan (= rows of the fused-statistics partials)
inline int conv_halo_grid(const ConvHaloParams& p, int sms) { return p.total_items < sms ? p.total_items : sms; }

}  // namespace icsg3d
