// tc_gemm16.cuh -- internal interface of the split-fp16 tiled GEMM for wide Linear layers (tc_gemm16.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "tc_gemm.cuh"

bool rbm_gemm16_supported(int64_t M, int N, int K, int64_t lda, const void* a, const void* b);
size_t rbm_gemm16_ws_bytes(int N, int K);
int rbm_gemm16_launch(const float* a, int64_t lda, const float* b, int64_t M, int N, int K, const RbmTcEpilogue& ep, void* ws, cudaStream_t st);
