// layernorm.cu -- hooked LayerNorm / LayerNormPre (reference models/layers/layer_norm.py:27-45, 75-93).
//
// HBM-bound: algorithmic traffic per row = cols*sizeof(in) read + cols*sizeof(out) write
// (+4 B scale, + cols*4 B when the fp32 hook_normalized copy or the tf32 residual is requested).
// One warp owns one row and keeps it in registers between the mean, the variance and the
// normalisation pass, so x is read from HBM exactly once; reductions are warp shuffles.
// The arithmetic order is the reference's: xc = x - mean;  scale = sqrt(mean(xc^2) + eps);
// y = (xc / scale) * w + b   -- all in fp32 whatever the storage dtype.
#include "common.cuh"

template <typename TI, typename TO, int CHUNKS>  // CHUNKS float4 per lane -> cols <= 128*CHUNKS
__global__ void __launch_bounds__(256) k_layernorm_reg(const TI* __restrict__ x, const TO* __restrict__ w, const TO* __restrict__ b,
                                                       float* __restrict__ scale_out, float* __restrict__ norm_f32,
                                                       TO* __restrict__ out, float* __restrict__ out_lo, int64_t rows, int cols,
                                                       float eps, const float* __restrict__ scale_in) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* xr = x + row * cols;
  const int nvec = cols >> 2;
  float v[CHUNKS][4];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c4 = i * 32 + lane;
    if (c4 < nvec) {
      ld4(xr + 4 * c4, v[i]);
      sum += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    } else {
      v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / (float)cols;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c4 = i * 32 + lane;
    if (c4 < nvec) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[i][j] -= mean; sq += v[i][j] * v[i][j]; }
    }
  }
  float scale = sqrtf(warp_sum(sq) / (float)cols + eps);
  if (scale_out && lane == 0) scale_out[row] = scale;
  if (scale_in) scale = scale_in[row];
#pragma unroll
  for (int i = 0; i < CHUNKS; ++i) {
    const int c4 = i * 32 + lane;
    if (c4 < nvec) {
      float y[4];
      if (w) {
        float ww[4], bb[4];
        ld4(w + 4 * c4, ww);
        ld4(b + 4 * c4, bb);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = (v[i][j] / scale) * ww[j] + bb[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = v[i][j] / scale;
      }
      const int64_t o = row * cols + 4 * c4;
      if (norm_f32) st4(norm_f32 + o, y);
      if (out) st4(out + o, y);
      if (out_lo) {
        float l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) l[j] = tf32_lo(y[j]);
        st4(out_lo + o, l);
      }
    }
  }
}

// any width / alignment: three passes over the row (L1/L2 serve passes 2 and 3)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_layernorm_generic(const TI* __restrict__ x, const TO* __restrict__ w, const TO* __restrict__ b,
                                                           float* __restrict__ scale_out, float* __restrict__ norm_f32,
                                                           TO* __restrict__ out, float* __restrict__ out_lo, int64_t rows, int cols,
                                                           float eps, const float* __restrict__ scale_in) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* xr = x + row * cols;
  float sum = 0.f;
  for (int c = lane; c < cols; c += 32) sum += ld_as_float(xr + c);
  const float mean = warp_sum(sum) / (float)cols;
  float sq = 0.f;
  for (int c = lane; c < cols; c += 32) { float d = ld_as_float(xr + c) - mean; sq += d * d; }
  float scale = sqrtf(warp_sum(sq) / (float)cols + eps);
  if (scale_out && lane == 0) scale_out[row] = scale;
  if (scale_in) scale = scale_in[row];
  for (int c = lane; c < cols; c += 32) {
    float y = (ld_as_float(xr + c) - mean) / scale;
    if (w) y = y * ld_as_float(w + c) + ld_as_float(b + c);
    const int64_t o = row * cols + c;
    if (norm_f32) norm_f32[o] = y;
    if (out) st_from_float(out + o, y);
    if (out_lo) out_lo[o] = tf32_lo(y);
  }
}

template <typename TI, typename TO>
static int launch_ln(const PbLayerNorm* p, cudaStream_t st) {
  const int warps = 8;
  const int grid = (int)ceil_div64(p->rows, warps);
  const TI* x = (const TI*)p->x;
  const TO* w = (const TO*)p->w;
  const TO* b = (const TO*)p->b;
  TO* out = (TO*)p->out;
  const bool aligned = ((((uintptr_t)p->x | (uintptr_t)p->w | (uintptr_t)p->b | (uintptr_t)p->out | (uintptr_t)p->norm_f32 |
                          (uintptr_t)p->out_lo) & 15) == 0) && (p->cols % 4 == 0);
#define PB_LN(CH) k_layernorm_reg<TI, TO, CH><<<grid, warps * 32, 0, st>>>(x, w, b, p->scale, p->norm_f32, out, p->out_lo, p->rows, p->cols, p->eps, p->scale_in)
  if (aligned && p->cols <= 128) PB_LN(1);
  else if (aligned && p->cols <= 256) PB_LN(2);
  else if (aligned && p->cols <= 512) PB_LN(4);
  else if (aligned && p->cols <= 768) PB_LN(6);
  else if (aligned && p->cols <= 1024) PB_LN(8);
  else if (aligned && p->cols <= 1536) PB_LN(12);
  else k_layernorm_generic<TI, TO><<<grid, warps * 32, 0, st>>>(x, w, b, p->scale, p->norm_f32, out, p->out_lo, p->rows, p->cols, p->eps, p->scale_in);
#undef PB_LN
  PB_LAUNCH_CHECK();
  return PB_OK;
}

extern "C" int pb_layernorm(const PbLayerNorm* p, pb_stream_t stream) {
  PB_CHECK_ARG(p && p->x && p->rows >= 0 && p->cols > 0, "pb_layernorm: bad arguments");
  PB_CHECK_ARG((p->w == nullptr) == (p->b == nullptr), "pb_layernorm: w and b must both be given or both be NULL");
  PB_CHECK_ARG(p->out || p->norm_f32 || p->scale, "pb_layernorm: no output requested");
  if (p->rows == 0) return PB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (p->dtype_in == PB_F32 && p->dtype_out == PB_F32) return launch_ln<float, float>(p, st);
  if (p->dtype_in == PB_BF16 && p->dtype_out == PB_BF16) return launch_ln<bf16, bf16>(p, st);
  if (p->dtype_in == PB_F32 && p->dtype_out == PB_BF16) return launch_ln<float, bf16>(p, st);
  if (p->dtype_in == PB_BF16 && p->dtype_out == PB_F32) return launch_ln<bf16, float>(p, st);
  PB_CHECK_ARG(false, "pb_layernorm: unknown dtype pair %d -> %d", p->dtype_in, p->dtype_out);
}
