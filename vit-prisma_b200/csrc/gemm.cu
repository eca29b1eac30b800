// gemm.cu -- pb_gemm: argument validation and dispatch between the two GEMM engines.
#include "common.cuh"
#include <stdlib.h>

int pb_gemm_simt(const PbGemm* g, cudaStream_t st);
int pb_gemm_tc(const PbGemm* g, cudaStream_t st);
bool pb_gemm_tc_eligible(const PbGemm* g);

extern "C" int pb_gemm(const PbGemm* g, pb_stream_t stream) {
  PB_CHECK_ARG(g, "pb_gemm: null descriptor");
  PB_CHECK_ARG(g->M >= 0 && g->N > 0 && g->K > 0, "pb_gemm: bad shape M=%d N=%d K=%d", g->M, g->N, g->K);
  PB_CHECK_ARG(g->dtype == PB_F32 || g->dtype == PB_BF16, "pb_gemm: unknown dtype %d", g->dtype);
  PB_CHECK_ARG(g->A && g->B, "pb_gemm: A and B are required");
  PB_CHECK_ARG(g->lda >= g->K && g->ldb >= g->K, "pb_gemm: leading dimensions smaller than K");
  PB_CHECK_ARG(!(g->residual && g->act != PB_ACT_NONE), "pb_gemm: residual and activation epilogues are exclusive");
  if (g->n_split > 1) {
    PB_CHECK_ARG(g->n_split <= 4 && g->split_n > 0 && (int64_t)g->n_split * g->split_n == g->N, "pb_gemm: n_split*split_n must equal N");
    for (int i = 0; i < g->n_split; ++i) PB_CHECK_ARG(g->out_split[i], "pb_gemm: out_split[%d] is NULL", i);
    PB_CHECK_ARG(g->ld0 >= g->split_n, "pb_gemm: ld0 smaller than split_n");
  } else {
    PB_CHECK_ARG(g->out0 || g->out1, "pb_gemm: no output requested");
    PB_CHECK_ARG(!g->out0 || g->ld0 >= g->N, "pb_gemm: ld0 smaller than N");
  }
  PB_CHECK_ARG(!g->out1 || g->ld1 >= g->N, "pb_gemm: ld1 smaller than N");
  PB_CHECK_ARG(!g->residual || g->ldr >= g->N, "pb_gemm: ldr smaller than N");
  PB_CHECK_ARG(!g->residual || g->out1, "pb_gemm: residual given without out1");
  PB_CHECK_ARG(!g->out1_lo || (g->out1 && g->dtype == PB_F32), "pb_gemm: out1_lo needs out1 and dtype F32");
  if (g->M == 0) return PB_OK;

  int impl = g->impl;
  if (impl == PB_GEMM_AUTO) {
    static int forced = -1;  // PB_GEMM_IMPL=simt|tc : debugging override of AUTO only
    if (forced < 0) {
      const char* e = getenv("PB_GEMM_IMPL");
      forced = !e ? 0 : (!strcmp(e, "simt") ? PB_GEMM_SIMT : (!strcmp(e, "tc") ? PB_GEMM_TC : 0));
    }
    if (forced == PB_GEMM_SIMT) impl = PB_GEMM_SIMT;
    else impl = (pb_gemm_tc_eligible(g) && g->M >= 64 && g->N >= 64) ? PB_GEMM_TC : PB_GEMM_SIMT;
  }
  if (impl == PB_GEMM_TC) return pb_gemm_tc(g, (cudaStream_t)stream);
  if (impl == PB_GEMM_SIMT) return pb_gemm_simt(g, (cudaStream_t)stream);
  PB_CHECK_ARG(false, "pb_gemm: unknown impl %d", impl);
}
