// sas_score.cu -- SASRec train scoring (gather + row dot), pos/neg BCE-with-logits, candidate scoring.
// Replaces SAS.forward NN/models/sas_model/sas.py:93-100, SASTrainer.calculate_loss NN/trainers/sas.py:38-49,
// SAS.predict NN/models/sas_model/sas.py:110-114 and the candidate gather NN/trainers/bert.py:47-49.
// HBM-bound gathers: a group of GW lanes owns one row and moves it with 16-byte loads.
#include "common.cuh"

namespace {

__device__ __forceinline__ float group_sum_rt(float v, int gw) {
  for (int o = gw >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
int pick_gw(int d4) { return d4 <= 4 ? 4 : d4 <= 8 ? 8 : d4 <= 16 ? 16 : 32; }

__global__ void __launch_bounds__(256) sas_score_fwd_kernel(const float* __restrict__ f, const float* __restrict__ table,
                                                            const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
                                                            float* __restrict__ pl, float* __restrict__ nl, int64_t rows, int d4, int gw) {
  int lane = threadIdx.x % gw;
  int64_t r = (int64_t)blockIdx.x * (blockDim.x / gw) + threadIdx.x / gw;
  bool live = r < rows;
  if (!live) r = rows - 1;
  const float* fr = f + r * d4 * 4;
  const float* pr = table + pos[r] * (int64_t)d4 * 4;
  const float* nr = table + neg[r] * (int64_t)d4 * 4;
  float sp = 0.f, sn = 0.f;
  for (int c4 = lane; c4 < d4; c4 += gw) {
    float4 a = ld4(fr + c4 * 4), p = ld4(pr + c4 * 4), n = ld4(nr + c4 * 4);
    sp += (a.x * p.x + a.y * p.y) + (a.z * p.z + a.w * p.w);
    sn += (a.x * n.x + a.y * n.y) + (a.z * n.z + a.w * n.w);
  }
  sp = group_sum_rt(sp, gw);
  sn = group_sum_rt(sn, gw);
  if (live && lane == 0) {
    pl[r] = sp;
    nl[r] = sn;
  }
}

__global__ void __launch_bounds__(256) sas_score_bwd_kernel(const float* __restrict__ table, const int64_t* __restrict__ pos,
                                                            const int64_t* __restrict__ neg, const float* __restrict__ dpl,
                                                            const float* __restrict__ dnl, float* __restrict__ df, int64_t total4, int d4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int64_t r = i / d4;
  int c4 = (int)(i - r * d4);
  float a = dpl[r], b = dnl[r];
  float4 p = ld4(table + pos[r] * (int64_t)d4 * 4 + c4 * 4), n = ld4(table + neg[r] * (int64_t)d4 * 4 + c4 * 4);
  st4(df + i * 4, make_float4(a * p.x + b * n.x, a * p.y + b * n.y, a * p.z + b * n.z, a * p.w + b * n.w));
}

__device__ __forceinline__ float softplus_f(float z) { return fmaxf(z, 0.f) + log1pf(expf(-fabsf(z))); }
__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }

// partial[b] = {sum softplus(-pl), sum softplus(nl), count} over the block's rows with pos != 0
__global__ void __launch_bounds__(256) bce_partial_kernel(const float* __restrict__ pl, const float* __restrict__ nl,
                                                          const int64_t* __restrict__ pos, int64_t rows, float* __restrict__ partial) {
  __shared__ float s0[256], s1[256], s2[256];
  int64_t base = (int64_t)blockIdx.x * 1024;
  float a = 0.f, b = 0.f, c = 0.f;
  for (int t = 0; t < 4; ++t) {
    int64_t i = base + t * 256 + threadIdx.x;
    if (i < rows && pos[i] != 0) {
      a += softplus_f(-pl[i]);
      b += softplus_f(nl[i]);
      c += 1.f;
    }
  }
  s0[threadIdx.x] = a; s1[threadIdx.x] = b; s2[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s0[threadIdx.x] += s0[threadIdx.x + o];
      s1[threadIdx.x] += s1[threadIdx.x + o];
      s2[threadIdx.x] += s2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[blockIdx.x * 3] = s0[0];
    partial[blockIdx.x * 3 + 1] = s1[0];
    partial[blockIdx.x * 3 + 2] = s2[0];
  }
}

__global__ void bce_finalize_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ loss, int32_t* __restrict__ count) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float a = 0.f, b = 0.f, c = 0.f;
    for (int i = 0; i < nblk; ++i) {
      a += partial[i * 3];
      b += partial[i * 3 + 1];
      c += partial[i * 3 + 2];
    }
    *loss = a / c + b / c;
    *count = (int32_t)c;
  }
}

__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ pl, const float* __restrict__ nl,
                                                      const int64_t* __restrict__ pos, const int32_t* __restrict__ count,
                                                      const float* __restrict__ dloss, float* __restrict__ dpl, float* __restrict__ dnl, int64_t rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float g = *dloss / (float)(*count);
  bool on = pos[i] != 0;
  dpl[i] = on ? g * (sigmoid_f(pl[i]) - 1.f) : 0.f;
  dnl[i] = on ? g * sigmoid_f(nl[i]) : 0.f;
}

__global__ void __launch_bounds__(256) candidate_scores_kernel(const float* __restrict__ f, int64_t ldf, const float* __restrict__ table,
                                                               const float* __restrict__ bias, const int64_t* __restrict__ cand,
                                                               float* __restrict__ out, int64_t total, int C, int d4, int gw) {
  int lane = threadIdx.x % gw;
  int64_t e = (int64_t)blockIdx.x * (blockDim.x / gw) + threadIdx.x / gw;
  bool live = e < total;
  if (!live) e = total - 1;
  int64_t u = e / C;
  int64_t item = cand[e];
  const float* fr = f + u * ldf;
  const float* tr = table + item * (int64_t)d4 * 4;
  float s = 0.f;
  for (int c4 = lane; c4 < d4; c4 += gw) {
    float4 a = ld4(fr + c4 * 4), t = ld4(tr + c4 * 4);
    s += (a.x * t.x + a.y * t.y) + (a.z * t.z + a.w * t.w);
  }
  s = group_sum_rt(s, gw);
  if (live && lane == 0) out[e] = s + (bias ? bias[item] : 0.f);
}

}  // namespace

extern "C" int rbm_sas_score_fwd(const float* f, const float* table, const int64_t* pos, const int64_t* neg, float* pos_logit,
                                 float* neg_logit, int64_t rows, int d, rbm_stream_t stream) {
  RBM_REQUIRE(f && table && pos && neg && pos_logit && neg_logit, "rbm_sas_score_fwd: null pointer");
  RBM_REQUIRE(rows > 0 && d >= 4 && d % 4 == 0, "rbm_sas_score_fwd: need rows>0, d%%4==0");
  RBM_REQUIRE(rbm_aligned16(f) && rbm_aligned16(table), "rbm_sas_score_fwd: pointers must be 16B aligned");
  int gw = pick_gw(d / 4);
  sas_score_fwd_kernel<<<(unsigned)rbm_cdiv(rows, 256 / gw), 256, 0, (cudaStream_t)stream>>>(f, table, pos, neg, pos_logit, neg_logit, rows, d / 4, gw);
  RBM_LAUNCH_CHECK("rbm_sas_score_fwd");
  return 0;
}

extern "C" int rbm_sas_score_bwd(const float* table, const int64_t* pos, const int64_t* neg, const float* dpl, const float* dnl,
                                 float* df, int64_t rows, int d, rbm_stream_t stream) {
  RBM_REQUIRE(table && pos && neg && dpl && dnl && df, "rbm_sas_score_bwd: null pointer");
  RBM_REQUIRE(rows > 0 && d >= 4 && d % 4 == 0, "rbm_sas_score_bwd: need rows>0, d%%4==0");
  RBM_REQUIRE(rbm_aligned16(df) && rbm_aligned16(table), "rbm_sas_score_bwd: pointers must be 16B aligned");
  int64_t total4 = rows * (d / 4);
  sas_score_bwd_kernel<<<(unsigned)rbm_cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(table, pos, neg, dpl, dnl, df, total4, d / 4);
  RBM_LAUNCH_CHECK("rbm_sas_score_bwd");
  return 0;
}

extern "C" size_t rbm_bce_ws_bytes(int64_t rows) { return (size_t)rbm_cdiv(rows, 1024) * 3 * sizeof(float); }

extern "C" int rbm_bce_pair_fwd(const float* pl, const float* nl, const int64_t* pos, float* loss, int32_t* count, int64_t rows,
                                void* ws, size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(pl && nl && pos && loss && count && ws, "rbm_bce_pair_fwd: null pointer");
  RBM_REQUIRE(rows > 0 && ws_bytes >= rbm_bce_ws_bytes(rows), "rbm_bce_pair_fwd: rows must be > 0 and workspace large enough");
  int nblk = (int)rbm_cdiv(rows, 1024);
  bce_partial_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(pl, nl, pos, rows, (float*)ws);
  bce_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)ws, nblk, loss, count);
  RBM_LAUNCH_CHECK("rbm_bce_pair_fwd");
  return 0;
}

extern "C" int rbm_bce_pair_bwd(const float* pl, const float* nl, const int64_t* pos, const int32_t* count, const float* dloss,
                                float* dpl, float* dnl, int64_t rows, rbm_stream_t stream) {
  RBM_REQUIRE(pl && nl && pos && count && dloss && dpl && dnl, "rbm_bce_pair_bwd: null pointer");
  RBM_REQUIRE(rows > 0, "rbm_bce_pair_bwd: rows must be > 0");
  bce_bwd_kernel<<<(unsigned)rbm_cdiv(rows, 256), 256, 0, (cudaStream_t)stream>>>(pl, nl, pos, count, dloss, dpl, dnl, rows);
  RBM_LAUNCH_CHECK("rbm_bce_pair_bwd");
  return 0;
}

extern "C" int rbm_candidate_scores(const float* f, int64_t ldf, const float* table, const float* bias, const int64_t* cand,
                                    float* out, int64_t U, int C, int d, rbm_stream_t stream) {
  RBM_REQUIRE(f && table && cand && out, "rbm_candidate_scores: null pointer");
  RBM_REQUIRE(U > 0 && C > 0 && d >= 4 && d % 4 == 0 && ldf % 4 == 0, "rbm_candidate_scores: bad sizes");
  RBM_REQUIRE(rbm_aligned16(f) && rbm_aligned16(table), "rbm_candidate_scores: pointers must be 16B aligned");
  int gw = pick_gw(d / 4);
  int64_t total = U * C;
  candidate_scores_kernel<<<(unsigned)rbm_cdiv(total, 256 / gw), 256, 0, (cudaStream_t)stream>>>(f, ldf, table, bias, cand, out, total, C, d / 4, gw);
  RBM_LAUNCH_CHECK("rbm_candidate_scores");
  return 0;
}
