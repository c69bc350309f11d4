// tc_gemm.cuh -- internal interface of the tcgen05 GEMM (tc_gemm.cu) used by the Linear entry points (linear.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct RbmTcEpilogue {
  float* y;
  int64_t ldy;
  float* pre;
  const float* bias;
  const float* residual;
  int64_t ldres;
  const int64_t* row_tok;
  int act;
  uint32_t thrA, thrB;
  float invA, invB;
  uint64_t siteA, siteB, seed;
};

// D[M,N] = A[M,K] . B[N,K]^T: shapes/alignments the TMA + UMMA path accepts (K % 32 == 0, N % 16 == 0, ...)
bool rbm_tc_linear_supported(int64_t M, int N, int K, int64_t lda, const void* a, const void* b);
int rbm_tc_linear_launch(const float* a, int64_t lda, const float* b, int64_t M, int N, int K, const RbmTcEpilogue& ep,
                         cudaStream_t st);

// dW[N,K] = dpre[M,N]^T . x[M,K] on tcgen05 (both operands MN-major): split over tokens into rbm_tc_dw_splits(M) partials
int rbm_tc_dw_splits(int64_t M);
bool rbm_tc_dw_supported(int64_t M, int N, int K, int64_t lda, int64_t ldb, const void* a, const void* b);
bool rbm_tc_dw_bias_fused(int N, int K);  // the bias gradient (column sums of dpre) fits next to dW in TMEM
int rbm_tc_dw_launch(const float* dpre, int64_t lda, const float* x, int64_t ldb, float* part, float* part_b, int64_t M, int N, int K,
                     cudaStream_t st);
