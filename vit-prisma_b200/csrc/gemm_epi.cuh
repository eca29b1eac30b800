// gemm_epi.cuh -- the hooked GEMM epilogue shared by gemm_simt.cu and gemm_tc.cu.
//
//   v    = round_T(acc + bias[n])                   -> out0 / out_split[n / split_n]   ("pre" hook point)
//   out1 = residual + v   or   act(v)                                                   (second hook point)
// In bf16 mode v is rounded to bf16 *before* out1 is formed because the reference materialises the
// bf16 tensor first (e.g. mlp.hook_pre) and applies the next op to it.
#pragma once
#include "common.cuh"

struct EpiParams {
  int M, N;
  int act;
  const void* bias;
  const void* residual; int64_t ldr;
  void* out0; int64_t ld0;
  void* out1; int64_t ld1;
  float* out1_lo;
  int n_split, split_n;
  void* out_split[4];
  int vec_ok;
  int vec16_ok;   // every output / residual row segment of 32 columns is a whole number of aligned 16-byte vectors
};

template <typename T>
__device__ __forceinline__ void epilogue_store4(const EpiParams& ep, int row, int col, float (&acc)[4]) {
  // acc: 4 consecutive columns [col, col+3] of one row
  const T* bias = (const T*)ep.bias;
  if (ep.vec_ok && col + 3 < ep.N) {
    float v[4];
    if (bias) {
      float bb[4];
      ld4(bias + col, bb);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = round_to<T>(round_to<T>(acc[j]) + bb[j]);  // einsum result, then "+ b": two bf16 tensors in the reference
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = round_to<T>(acc[j]);
    }
    if (ep.n_split > 1) {
      const int blk = col / ep.split_n;
      st4((T*)ep.out_split[blk] + (int64_t)row * ep.ld0 + (col - blk * ep.split_n), v);
    } else if (ep.out0) {
      st4((T*)ep.out0 + (int64_t)row * ep.ld0 + col, v);
    }
    if (ep.out1) {
      float o[4];
      if (ep.residual) {
        float rr[4];
        ld4((const T*)ep.residual + (int64_t)row * ep.ldr + col, rr);
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = rr[j] + v[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = apply_act(v[j], ep.act);
      }
      st4((T*)ep.out1 + (int64_t)row * ep.ld1 + col, o);
      if (ep.out1_lo) {
        float l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) l[j] = tf32_lo(o[j]);
        st4(ep.out1_lo + (int64_t)row * ep.ld1 + col, l);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col + j;
      if (c >= ep.N) break;
      const float v = bias ? round_to<T>(round_to<T>(acc[j]) + ld_as_float(bias + c)) : round_to<T>(acc[j]);
      if (ep.n_split > 1) {
        const int blk = c / ep.split_n;
        st_from_float((T*)ep.out_split[blk] + (int64_t)row * ep.ld0 + (c - blk * ep.split_n), v);
      } else if (ep.out0) {
        st_from_float((T*)ep.out0 + (int64_t)row * ep.ld0 + c, v);
      }
      if (ep.out1) {
        const float o = ep.residual ? ld_as_float((const T*)ep.residual + (int64_t)row * ep.ldr + c) + v : apply_act(v, ep.act);
        st_from_float((T*)ep.out1 + (int64_t)row * ep.ld1 + c, o);
        if (ep.out1_lo) ep.out1_lo[(int64_t)row * ep.ld1 + c] = tf32_lo(o);
      }
    }
  }
}


static inline bool pb_aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// Host: PbGemm -> EpiParams; vec_ok says whether every epilogue access may be a 4-element vector.
static inline EpiParams pb_make_epi(const PbGemm* g) {
  EpiParams ep;
  ep.M = g->M; ep.N = g->N; ep.act = g->act;
  ep.bias = g->bias; ep.residual = g->residual; ep.ldr = g->ldr;
  ep.out0 = g->out0; ep.ld0 = g->ld0; ep.out1 = g->out1; ep.ld1 = g->ld1; ep.out1_lo = g->out1_lo;
  ep.n_split = g->n_split > 1 ? g->n_split : 1; ep.split_n = g->split_n;
  for (int i = 0; i < 4; ++i) ep.out_split[i] = g->out_split[i];
  bool v = (g->N % 4 == 0) && pb_aligned16(g->bias) && pb_aligned16(g->residual) && pb_aligned16(g->out0) && pb_aligned16(g->out1) && pb_aligned16(g->out1_lo) &&
           (g->ld0 % 4 == 0) && (g->ld1 % 4 == 0) && (g->ldr % 4 == 0);
  if (ep.n_split > 1) {
    v = v && (g->split_n % 4 == 0);
    for (int i = 0; i < ep.n_split; ++i) v = v && pb_aligned16(g->out_split[i]);
  }
  ep.vec_ok = v ? 1 : 0;
  const int e16 = g->dtype == PB_BF16 ? 8 : 4;   // elements per 16 bytes
  bool w = v && (g->ld0 % e16 == 0) && (g->ld1 % e16 == 0) && (g->ldr % e16 == 0);
  if (ep.n_split > 1) w = w && (g->split_n % 32 == 0);
  ep.vec16_ok = w ? 1 : 0;
  return ep;
}
