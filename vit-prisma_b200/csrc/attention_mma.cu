// attention_mma.cu -- fused hooked attention for d_head == 64 on warp-level tensor-core MMAs.
//
// Why not tcgen05 here: one head is a 50..257-token problem (S = Q K^T is 50x50 for ViT-B/32) and the op is bound by
// the HBM traffic of its hook points (scores + pattern = 2*B*H*T*T elements written) -- a 128-row UMMA tile with a TMEM
// round trip would be mostly padding.  Warp-level mma.sync keeps S and P in registers between QK^T, softmax and PV.
//
// One CTA = one (batch, head) x one slab of NW*16 query rows; K, V (and the Q slab) of the head sit in shared memory.
//   bf16 : mma.sync.m16n8k16 bf16 (fp32 accumulate).  V is stored transposed so B fragments are 32-bit loads.
//   fp32 : mma.sync.m16n8k8 tf32 in 3 passes (x = hi + lo, hi = what the tensor core reads of x, lo = x - hi:
//          lo*hi + hi*lo + hi*hi) -> fp32-grade products for the 1e-4 parity bar.
// Hook points are spilled through a per-warp shared-memory stage so that global stores are contiguous rows
// (the 16 x T score rows of a warp are one contiguous run of memory), never fragment-shaped partial sectors.
// Rounding points follow the reference graph: scores = round(round(q.k) / scale); pattern = round(softmax);
// z = round(pattern @ v) with the rounded pattern as the operand.
#include "common.cuh"

namespace {

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// 3xTF32: operands given as fp32 bit patterns
__device__ __forceinline__ void mma_tf32x3(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float hi = tf32_trunc(a[i]);
    ah[i] = __float_as_uint(hi);
    al[i] = __float_as_uint(a[i] - hi);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float hi = tf32_trunc(b[i]);
    bh[i] = __float_as_uint(hi);
    bl[i] = __float_as_uint(b[i] - hi);
  }
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int DH = 64;

template <typename T> struct Lay;
template <> struct Lay<float> {
  static constexpr int QK_LD = DH + 4;   // words; bank = 4g + t
  static constexpr int V_LD = DH + 4;    // row-major V [TPAD][68]; PV B rows 2t / 2t+1 -> bank 8t (+4) + g
  __host__ __device__ static constexpr size_t v_elems(int tpad) { return (size_t)tpad * V_LD; }
};
template <> struct Lay<bf16> {
  static constexpr int QK_LD = DH + 8;   // elements (36 words)
  __host__ __device__ static constexpr size_t v_elems(int tpad) { return (size_t)DH * (tpad + 8); }   // transposed V^T [64][TPAD+8]
};

template <typename T, int NT, int NW>  // NT = key tiles of 8 (TPAD = 8*NT, multiple of 16), NW warps of 16 query rows
__global__ void __launch_bounds__(NW * 32) k_attention_mma(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                          T* __restrict__ scores, T* __restrict__ pattern, T* __restrict__ z, int Tn, int H,
                                                          float attn_scale) {
  constexpr int TPAD = NT * 8;
  constexpr bool BF = sizeof(T) == 2;
  constexpr int QK_LD = Lay<T>::QK_LD;
  constexpr int ST_LD = TPAD + 4;                       // stage row stride (floats)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Ks = reinterpret_cast<T*>(smem_raw);                 // [TPAD][QK_LD]
  T* Vs = Ks + (size_t)TPAD * QK_LD;                      // fp32: [TPAD][68]; bf16: V^T [64][TPAD+8]
  T* Qs = Vs + Lay<T>::v_elems(TPAD);                     // [NW*16][QK_LD]
  float* stage_all = reinterpret_cast<float*>(Qs + (size_t)NW * 16 * QK_LD);   // [NW][16][ST_LD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  const int row0 = blockIdx.y * (NW * 16);
  const int64_t tok_stride = (int64_t)H * DH;
  const int64_t head_base = (int64_t)b * Tn * tok_stride + (int64_t)h * DH;

  // ---- stage K, V (whole head) and the Q slab; rows >= Tn are zero
  for (int idx = threadIdx.x; idx < TPAD * (DH / 4); idx += NW * 32) {
    const int j = idx / (DH / 4), e4 = (idx % (DH / 4)) * 4;
    float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < Tn) {
      ld4(k + head_base + (int64_t)j * tok_stride + e4, kv);
      ld4(v + head_base + (int64_t)j * tok_stride + e4, vv);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      st_from_float(Ks + (size_t)j * QK_LD + e4 + c, kv[c]);
      if (BF) st_from_float(Vs + (size_t)(e4 + c) * (TPAD + 8) + j, vv[c]);
      else st_from_float(Vs + (size_t)j * (DH + 4) + e4 + c, vv[c]);
    }
  }
  for (int idx = threadIdx.x; idx < NW * 16 * (DH / 4); idx += NW * 32) {
    const int r = idx / (DH / 4), e4 = (idx % (DH / 4)) * 4;
    float qv[4] = {0.f, 0.f, 0.f, 0.f};
    if (row0 + r < Tn) ld4(q + head_base + (int64_t)(row0 + r) * tok_stride + e4, qv);
#pragma unroll
    for (int c = 0; c < 4; ++c) st_from_float(Qs + (size_t)r * QK_LD + e4 + c, qv[c]);
  }
  __syncthreads();

  const int wrow0 = row0 + warp * 16;                     // first query row of this warp
  if (wrow0 >= Tn) return;
  const int nvalid = min(16, Tn - wrow0);
  float* stage = stage_all + (size_t)warp * 16 * ST_LD;
  const T* Qw = Qs + (size_t)warp * 16 * QK_LD;

  // ---- S = Q K^T
  float acc[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
  if (BF) {
    const uint32_t* Qw32 = reinterpret_cast<const uint32_t*>(Qw);
    const uint32_t* Ks32 = reinterpret_cast<const uint32_t*>(Ks);
    constexpr int LDW = QK_LD / 2;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      uint32_t a[4];
      a[0] = Qw32[g * LDW + kk * 8 + t];
      a[1] = Qw32[(g + 8) * LDW + kk * 8 + t];
      a[2] = Qw32[g * LDW + kk * 8 + 4 + t];
      a[3] = Qw32[(g + 8) * LDW + kk * 8 + 4 + t];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t bb[2];
        bb[0] = Ks32[(nt * 8 + g) * LDW + kk * 8 + t];
        bb[1] = Ks32[(nt * 8 + g) * LDW + kk * 8 + 4 + t];
        mma_bf16(acc[nt], a, bb);
      }
    }
  } else {
    const float* Qf = reinterpret_cast<const float*>(Qw);
    const float* Kf = reinterpret_cast<const float*>(Ks);
#pragma unroll
    for (int kk = 0; kk < DH / 8; ++kk) {
      float a[4];
      a[0] = Qf[g * QK_LD + kk * 8 + t];
      a[1] = Qf[(g + 8) * QK_LD + kk * 8 + t];
      a[2] = Qf[g * QK_LD + kk * 8 + 4 + t];
      a[3] = Qf[(g + 8) * QK_LD + kk * 8 + 4 + t];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float bb[2];
        bb[0] = Kf[(nt * 8 + g) * QK_LD + kk * 8 + t];
        bb[1] = Kf[(nt * 8 + g) * QK_LD + kk * 8 + 4 + t];
        mma_tf32x3(acc[nt], a, bb);
      }
    }
  }

  // ---- scores: scale + round; row max over valid keys. rows: lo = g, hi = g + 8; cols nt*8 + 2t + {0,1}
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = nt * 8 + 2 * t + (c & 1);
      const float s = round_to<T>(round_to<T>(acc[nt][c]) / attn_scale);
      acc[nt][c] = s;
      if (col < Tn) { if (c < 2) mx_lo = fmaxf(mx_lo, s); else mx_hi = fmaxf(mx_hi, s); }
    }
  }
  const int64_t sc_base = (((int64_t)b * H + h) * Tn + wrow0) * (int64_t)Tn;   // warp's rows are one contiguous run
  if (scores) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      stage[g * ST_LD + nt * 8 + 2 * t] = acc[nt][0];
      stage[g * ST_LD + nt * 8 + 2 * t + 1] = acc[nt][1];
      stage[(g + 8) * ST_LD + nt * 8 + 2 * t] = acc[nt][2];
      stage[(g + 8) * ST_LD + nt * 8 + 2 * t + 1] = acc[nt][3];
    }
    __syncwarp();
    for (int i = lane; i < nvalid * Tn; i += 32) {
      const int r = i / Tn, c = i - r * Tn;
      st_from_float(scores + sc_base + i, stage[r * ST_LD + c]);
    }
    __syncwarp();
  }
  // ---- softmax (NaN -> 0), rounded to T
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = nt * 8 + 2 * t + (c & 1);
      const float e = col < Tn ? expf(acc[nt][c] - (c < 2 ? mx_lo : mx_hi)) : 0.f;
      acc[nt][c] = e;
      if (c < 2) sum_lo += e; else sum_hi += e;
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int col = nt * 8 + 2 * t + (c & 1);
      float p = acc[nt][c] / (c < 2 ? sum_lo : sum_hi);
      if (isnan(p)) p = 0.f;
      acc[nt][c] = col < Tn ? round_to<T>(p) : 0.f;
    }
  }
  if (pattern) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      stage[g * ST_LD + nt * 8 + 2 * t] = acc[nt][0];
      stage[g * ST_LD + nt * 8 + 2 * t + 1] = acc[nt][1];
      stage[(g + 8) * ST_LD + nt * 8 + 2 * t] = acc[nt][2];
      stage[(g + 8) * ST_LD + nt * 8 + 2 * t + 1] = acc[nt][3];
    }
    __syncwarp();
    for (int i = lane; i < nvalid * Tn; i += 32) {
      const int r = i / Tn, c = i - r * Tn;
      st_from_float(pattern + sc_base + i, stage[r * ST_LD + c]);
    }
    __syncwarp();
  }

  // ---- Z = P V  (P straight from the accumulator registers)
  float o[DH / 8][4];
#pragma unroll
  for (int nn = 0; nn < DH / 8; ++nn) o[nn][0] = o[nn][1] = o[nn][2] = o[nn][3] = 0.f;
  if (BF) {
    const uint32_t* Vt32 = reinterpret_cast<const uint32_t*>(Vs);
    constexpr int LDW = (TPAD + 8) / 2;
#pragma unroll
    for (int kk = 0; kk < TPAD / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(acc[2 * kk][0], acc[2 * kk][1]);
      a[1] = pack_bf16(acc[2 * kk][2], acc[2 * kk][3]);
      a[2] = pack_bf16(acc[2 * kk + 1][0], acc[2 * kk + 1][1]);
      a[3] = pack_bf16(acc[2 * kk + 1][2], acc[2 * kk + 1][3]);
#pragma unroll
      for (int nn = 0; nn < DH / 8; ++nn) {
        uint32_t bb[2];
        bb[0] = Vt32[(nn * 8 + g) * LDW + kk * 8 + t];
        bb[1] = Vt32[(nn * 8 + g) * LDW + kk * 8 + 4 + t];
        mma_bf16(o[nn], a, bb);
      }
    }
  } else {
    const float* Vf = reinterpret_cast<const float*>(Vs);
    constexpr int VLD = DH + 4;
#pragma unroll
    for (int kk = 0; kk < TPAD / 8; ++kk) {
      // k-slot t <-> key kk*8 + 2t, k-slot t+4 <-> key kk*8 + 2t + 1 (same permutation on A and B: the sum over k is unchanged)
      float a[4] = {acc[kk][0], acc[kk][2], acc[kk][1], acc[kk][3]};
#pragma unroll
      for (int nn = 0; nn < DH / 8; ++nn) {
        float bb[2];
        bb[0] = Vf[(kk * 8 + 2 * t) * VLD + nn * 8 + g];
        bb[1] = Vf[(kk * 8 + 2 * t + 1) * VLD + nn * 8 + g];
        mma_tf32x3(o[nn], a, bb);
      }
    }
  }
  // ---- z rows through the stage: [16][DH] -> each token row is 64 contiguous elements in [B,T,H,dh]
#pragma unroll
  for (int nn = 0; nn < DH / 8; ++nn) {
    stage[g * ST_LD + nn * 8 + 2 * t] = o[nn][0];
    stage[g * ST_LD + nn * 8 + 2 * t + 1] = o[nn][1];
    stage[(g + 8) * ST_LD + nn * 8 + 2 * t] = o[nn][2];
    stage[(g + 8) * ST_LD + nn * 8 + 2 * t + 1] = o[nn][3];
  }
  __syncwarp();
  for (int r = 0; r < nvalid; ++r) {
    T* zr = z + head_base + (int64_t)(wrow0 + r) * tok_stride;
    st_from_float(zr + lane, stage[r * ST_LD + lane]);
    st_from_float(zr + 32 + lane, stage[r * ST_LD + 32 + lane]);
  }
}

template <typename T, int NT, int NW>
int launch_mma(const PbAttention* p, cudaStream_t st) {
  constexpr int TPAD = NT * 8;
  const size_t es = sizeof(T);
  const size_t smem = ((size_t)TPAD * Lay<T>::QK_LD + Lay<T>::v_elems(TPAD) + (size_t)NW * 16 * Lay<T>::QK_LD) * es + (size_t)NW * 16 * (TPAD + 4) * 4;
  if (smem > 227 * 1024) return PB_EUNSUPPORTED;
  auto kern = k_attention_mma<T, NT, NW>;
  static bool attr_done = false;
  if (!attr_done && smem > 48 * 1024) {
    PB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  dim3 grid(p->B * p->H, (p->T + NW * 16 - 1) / (NW * 16));
  kern<<<grid, NW * 32, smem, st>>>((const T*)p->q, (const T*)p->k, (const T*)p->v, (T*)p->scores, (T*)p->pattern, (T*)p->z, p->T, p->H,
                                    p->attn_scale);
  PB_LAUNCH_CHECK();
  return PB_OK;
}

template <typename T>
int dispatch_mma(const PbAttention* p, cudaStream_t st) {
  if (p->T <= 64) return launch_mma<T, 8, 4>(p, st);
  if (p->T <= 128) return launch_mma<T, 16, 4>(p, st);
  if (p->T <= 208) return launch_mma<T, 26, 2>(p, st);
  if (p->T <= 272) return launch_mma<T, 34, 2>(p, st);
  return PB_EUNSUPPORTED;
}

}  // namespace

// PB_OK when the tensor-core kernel took the call, PB_EUNSUPPORTED when the shape is not covered (caller falls back to
// the FFMA kernel in attention.cu), anything else is an error.
int pb_attention_mma(const PbAttention* p, cudaStream_t st) {
  if (p->dh != DH) return PB_EUNSUPPORTED;
  if (((uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v) & 15) return PB_EUNSUPPORTED;
  return p->dtype == PB_F32 ? dispatch_mma<float>(p, st) : dispatch_mma<bf16>(p, st);
}
