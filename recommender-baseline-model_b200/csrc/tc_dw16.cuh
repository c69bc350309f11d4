// tc_dw16.cuh -- internal interface of the split-fp16 tcgen05 weight-gradient kernel for wide Linear layers (tc_dw16.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

bool rbm_dw16_supported(int64_t M, int N, int K, int64_t lda, int64_t ldb, const void* a, const void* b);
int rbm_dw16_splits(int64_t M, int N, int K);
size_t rbm_dw16_extra_bytes();
int rbm_dw16_launch(const float* dpre, int64_t lda, const float* x, int64_t ldb, float* part, float* part_b, void* aux, int64_t M, int N, int K,
                    cudaStream_t st);
