// tcgen05 tensor-core path of the dense layers (gemm_tc.cu), dispatched from tn_dense_*.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tn {

// dense path selector (tn_set_dense_mode): 0 auto (tensor cores, 3xTF32, when the shape allows),
// 1 CUDA-core SGEMM only, 2 tensor cores with a single TF32 pass, 3 tensor cores 3xTF32
extern int g_dense_mode;

bool dense_tc_ok(int n_in, int n_out, const void *p0, const void *p1, const void *p2);
int dense_tc_fwd(const float *x, const float *W, const float *bias, float *out, int B, int n_in,
                 int n_out, int act, float act_nn, int mask_on, uint32_t thr, uint64_t seed,
                 const int32_t *ctl, const float *mask_inj, float scale, int split, int max_sms,
                 cudaStream_t st);
int dense_tc_bwd_data(const float *gr, const float *W, float *dx, int B, int n_in, int n_out,
                      const float *prev_out, int act, float act_nn, int mask_on, uint32_t thr,
                      uint64_t seed, const int32_t *ctl, const float *mask_inj, int split,
                      int max_sms, cudaStream_t st);
int dense_tc_bwd_weights(const float *x, const float *gr, float *dW, int B, int n_in, int n_out,
                         int split, int max_sms, cudaStream_t st);
// `split`: 0 = one TF32 pass, 1 = 3xTF32 on the cluster split-K kernel, 2 = 3xTF32 on the
// first-generation kernel (per-k-block promotion); `max_sms` > 0 bounds the CTAs (= SMs) of the
// cluster split-K kernel so that two products can run side by side
void dense_tc_set_debug(long long *buf);
int tc_make_map_2d(CUtensorMap *map, const float *ptr, int rows, int cols, int ld, int box_cols,
                   int box_rows, int atom32, const char *who);

}  // namespace tn
