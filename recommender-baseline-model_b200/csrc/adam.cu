// adam.cu -- dense Adam over every parameter tensor in ONE launch + flat-bucket pack/unpack for the DP all-reduce.
// Replaces optim.Adam(...).step() (NN/trainers/base.py:123,225-233): dense semantics -- every element of every
// tensor (whole item tables included) is updated every step; 28 B/param of HBM traffic, nothing else.
// Arithmetic follows torch's formula:  g += wd*p;  m = m + (g-m)*(1-b1);  v = v*b2 + (1-b2)*g*g;
//                                      p += -(lr/(1-b1^t)) * ( m / ( sqrt(v)/sqrt(1-b2^t) + eps ) ).
#include "common.cuh"

namespace {

struct AdamHyper {
  float one_minus_b1, b2, one_minus_b2, eps, wd, neg_step_size, bc2_sqrt;
  double lr, beta1, beta2;  // for the in-kernel bias correction under a device-side step counter (CUDA-graph replay)
  int step;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamHyper& h) {
  if (h.wd != 0.f) g = g + h.wd * p;
  m = m + (g - m) * h.one_minus_b1;
  v = v * h.b2 + h.one_minus_b2 * g * g;
  float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p + h.neg_step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const rbm_adam_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_map,
                                                         AdamHyper h) {
  if (rbm_step_ptr_dev != nullptr) {  // replayed graph: the step advances on the device (same double arithmetic as the host path)
    const double t = (double)h.step + (double)*rbm_step_ptr_dev;
    h.neg_step_size = (float)(-(h.lr / (1.0 - pow(h.beta1, t))));
    h.bc2_sqrt = (float)sqrt(1.0 - pow(h.beta2, t));
  }
  const int ti = chunk_map[blockIdx.x * 2], ci = chunk_map[blockIdx.x * 2 + 1];
  const rbm_adam_tensor t = tensors[ti];
  const int64_t b = (int64_t)ci * RBM_ADAM_CHUNK;
  const int64_t e = b + RBM_ADAM_CHUNK < t.n ? b + RBM_ADAM_CHUNK : t.n;
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15) == 0;
  if (vec) {
    const int64_t e4 = b + ((e - b) & ~(int64_t)3);
    for (int64_t i = b + threadIdx.x * 4; i < e4; i += 256 * 4) {
      float4 p = ld4(t.p + i), g = ld4(t.g + i), m = ld4(t.m + i), v = ld4(t.v + i);
      adam_elem(p.x, g.x, m.x, v.x, h);
      adam_elem(p.y, g.y, m.y, v.y, h);
      adam_elem(p.z, g.z, m.z, v.z, h);
      adam_elem(p.w, g.w, m.w, v.w, h);
      st4(t.p + i, p);
      st4(t.m + i, m);
      st4(t.v + i, v);
    }
    for (int64_t i = e4 + threadIdx.x; i < e; i += 256) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_elem(p, t.g[i], m, v, h);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
    }
  } else {
    for (int64_t i = b + threadIdx.x; i < e; i += 256) {
      float p = t.p[i], m = t.m[i], v = t.v[i];
      adam_elem(p, t.g[i], m, v, h);
      t.p[i] = p; t.m[i] = m; t.v[i] = v;
    }
  }
}

__global__ void __launch_bounds__(256) bucket_pack_kernel(const rbm_bucket_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_map,
                                                          float* __restrict__ bucket, float scale, int unpack) {
  const int ti = chunk_map[blockIdx.x * 2], ci = chunk_map[blockIdx.x * 2 + 1];
  const rbm_bucket_tensor t = tensors[ti];
  const int64_t b = (int64_t)ci * RBM_ADAM_CHUNK;
  const int64_t e = b + RBM_ADAM_CHUNK < t.n ? b + RBM_ADAM_CHUNK : t.n;
  float* bk = bucket + t.offset;
  for (int64_t i = b + threadIdx.x; i < e; i += 256) {
    if (unpack) t.ptr[i] = bk[i] * scale;
    else bk[i] = t.ptr[i] * scale;
  }
}

}  // namespace

extern "C" int rbm_adam_multi(const rbm_adam_tensor* tensors, const int32_t* chunk_map, int total_chunks, double lr, double beta1,
                              double beta2, double eps, double weight_decay, int step, rbm_stream_t stream) {
  RBM_REQUIRE(tensors && chunk_map, "rbm_adam_multi: null pointer");
  RBM_REQUIRE(total_chunks >= 0 && step >= 1, "rbm_adam_multi: need step >= 1");
  if (total_chunks == 0) return 0;
  // hyper-parameters arrive as doubles (python floats) and are rounded to fp32 once, like torch's scalar arguments
  double bc1 = 1.0 - pow(beta1, (double)step);
  double bc2 = 1.0 - pow(beta2, (double)step);
  AdamHyper h;
  h.one_minus_b1 = (float)(1.0 - beta1);
  h.b2 = (float)beta2;
  h.one_minus_b2 = (float)(1.0 - beta2);
  h.eps = (float)eps;
  h.wd = (float)weight_decay;
  h.neg_step_size = (float)(-(lr / bc1));
  h.bc2_sqrt = (float)sqrt(bc2);
  h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.step = step;
  adam_multi_kernel<<<total_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_map, h);
  RBM_LAUNCH_CHECK("rbm_adam_multi");
  return 0;
}

extern "C" int rbm_bucket_pack(const rbm_bucket_tensor* tensors, const int32_t* chunk_map, int total_chunks, float* bucket, float scale,
                               int unpack, rbm_stream_t stream) {
  RBM_REQUIRE(tensors && chunk_map && bucket, "rbm_bucket_pack: null pointer");
  if (total_chunks <= 0) return 0;
  bucket_pack_kernel<<<total_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunk_map, bucket, scale, unpack);
  RBM_LAUNCH_CHECK("rbm_bucket_pack");
  return 0;
}

RBM_DEFINE_STEP_PTR_SETTER(rbm_step_ptr_set_adam)

// ---- public: device-side step counter (see common.cuh)
int rbm_step_ptr_set_embed(const unsigned long long*);
int rbm_step_ptr_set_linear(const unsigned long long*);
int rbm_step_ptr_set_tc_gemm(const unsigned long long*);
int rbm_step_ptr_set_attention(const unsigned long long*);
int rbm_step_ptr_set_attention_tc(const unsigned long long*);
int rbm_step_ptr_set_attention_pair(const unsigned long long*);
int rbm_step_ptr_set_attention_seq(const unsigned long long*);
int rbm_step_ptr_set_gemm16(const unsigned long long*);
int rbm_step_ptr_set_attention_live(const unsigned long long*);

extern "C" int rbm_set_step_counter(const uint64_t* counter) {
  const unsigned long long* p = reinterpret_cast<const unsigned long long*>(counter);
  int rc = rbm_step_ptr_set_embed(p);
  if (!rc) rc = rbm_step_ptr_set_linear(p);
  if (!rc) rc = rbm_step_ptr_set_tc_gemm(p);
  if (!rc) rc = rbm_step_ptr_set_attention(p);
  if (!rc) rc = rbm_step_ptr_set_attention_tc(p);
  if (!rc) rc = rbm_step_ptr_set_attention_pair(p);
  if (!rc) rc = rbm_step_ptr_set_attention_seq(p);
  if (!rc) rc = rbm_step_ptr_set_gemm16(p);
  if (!rc) rc = rbm_step_ptr_set_attention_live(p);
  if (!rc) rc = rbm_step_ptr_set_adam(p);
  if (rc) rbm_set_error("rbm_set_step_counter: cudaMemcpyToSymbol: %s", cudaGetErrorString((cudaError_t)rc));
  return rc;
}
