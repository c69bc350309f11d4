// ce_wide.cuh -- internal interface of the split-fp16 tcgen05 scoring + cross-entropy kernels for d = 128 / 256 (ce_wide.cu),
// dispatched from score_ce.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

bool rbm_ce_wide_supported(int V1, int d, const void* h, const void* w);
size_t rbm_ce_wide_ws_floats(int64_t cap, int V1, int d);
int rbm_ce_wide_dw_splits(int64_t cap, int V1);
int rbm_ce_wide_fwd(const float* h, const int32_t* rows, const int64_t* tgt, const int32_t* count, const float* w, const float* bias,
                    float* lse, float* partial, int64_t cap, int V1, int d, float* extra_ws, int* nblk_out, cudaStream_t st);
int rbm_ce_wide_bwd(const float* h, const int32_t* rows, const int64_t* tgt, const int32_t* count, const float* w, const float* bias,
                    const float* lse, const float* dloss, float* dh_full, float* part_w, float* part_b, int S, int64_t cap, int V1, int d,
                    float* extra_ws, cudaStream_t st);
