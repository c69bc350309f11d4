// attention_seq.cuh -- internal interface of the split-fp16 tcgen05 attention kernels for d_k = 64, 64 < L <= 256
// (attention_seq.cu), dispatched from attention.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

bool rbm_attn_seq_supported(int L, int dk, int mask_mode);
int rbm_attn_seq_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const int64_t* tok, float* out,
                     int64_t ldo, float* stats, int B, int L, int h, int mask_mode, float scale, float p, uint64_t seed, uint64_t site,
                     cudaStream_t st);
int rbm_attn_seq_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, const int64_t* tok, const float* out,
                     int64_t ldo, const float* dout, int64_t lddo, const float* stats, float* dq, int64_t lddq, float* dk_, int64_t lddk,
                     float* dv, int64_t lddv, int B, int L, int h, int mask_mode, float scale, float p, uint64_t seed, uint64_t site,
                     cudaStream_t st);
