// attention_last.cu -- evaluation only: attention of the LAST query position of every sequence (K20).  Both models rank items from
// the hidden state of the final position alone (NN/trainers/bert.py:47 `scores[:, -1, :]`, NN/models/sas_model/sas.py:111
// `log_feats[:, -1, :]`), so in the last block only that position's query, output projection and feed-forward are needed; keys and
// values still come from every position.  One CTA per (sequence, head): a thread per key for the scores, a thread per output column
// for P.v.  Same masking rules as the training kernels (attention/single.py:18-22: padded keys filled with -1e9; the causal mask of
// sas.py:69-76 leaves the last query every key); no dropout (evaluation).
#include "common.cuh"
#include "mma_tiles.cuh"  // RBM_PADFILL

namespace {

constexpr int LQ_THREADS = 128;

__global__ void __launch_bounds__(LQ_THREADS) attn_last_query_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k,
                                                                      int64_t ldk, const float* __restrict__ v, int64_t ldv,
                                                                      const int64_t* __restrict__ tok, float* __restrict__ out, int64_t ldo,
                                                                      int L, int h, int dk, int mask_mode, float scale) {
  __shared__ float sq[128], sp[256], red[LQ_THREADS / 32];
  const int b = blockIdx.x / h, hh = blockIdx.x - b * h, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int c = t; c < dk; c += LQ_THREADS) sq[c] = q[(int64_t)b * ldq + hh * dk + c];
  __syncthreads();
  // scores of the last query against every key (fp32, sequential over d_k like the SIMT training kernel)
  float mx = -INFINITY;
  for (int j = t; j < L; j += LQ_THREADS) {
    const float* kr = k + ((int64_t)b * L + j) * ldk + hh * dk;
    float s = 0.f;
    for (int c = 0; c < dk; c += 4) {
      const float4 kv = ld4(kr + c);
      s = fmaf(sq[c], kv.x, s); s = fmaf(sq[c + 1], kv.y, s); s = fmaf(sq[c + 2], kv.z, s); s = fmaf(sq[c + 3], kv.w, s);
    }
    s *= scale;
    if (mask_mode == RBM_MASK_KEYPAD && tok[(int64_t)b * L + j] == 0) s = RBM_PADFILL;
    sp[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int j = t; j < L; j += LQ_THREADS) {
    const float p = expf(sp[j] - mx);
    sp[j] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
  for (int c = t; c < dk; c += LQ_THREADS) {
    const float* vc = v + (int64_t)b * L * ldv + hh * dk + c;
    float acc = 0.f;
    for (int j = 0; j < L; ++j) acc = fmaf(sp[j], vc[(int64_t)j * ldv], acc);
    out[(int64_t)b * ldo + hh * dk + c] = acc * inv;
  }
}

}  // namespace

extern "C" int rbm_attn_last_query(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                                   const int64_t* tok, float* out, int64_t ldo, int B, int L, int h, int dk, int mask_mode,
                                   float scale, rbm_stream_t stream) {
  RBM_REQUIRE(q && k && v && out, "rbm_attn_last_query: null pointer");
  RBM_REQUIRE(B >= 0 && L >= 1 && L <= 256 && h >= 1 && dk >= 4 && dk <= 128 && dk % 4 == 0,
              "rbm_attn_last_query: need 1 <= L <= 256, d_k %% 4 == 0, d_k <= 128 (L=%d d_k=%d)", L, dk);
  RBM_REQUIRE(mask_mode == RBM_MASK_NONE || mask_mode == RBM_MASK_CAUSAL || mask_mode == RBM_MASK_KEYPAD, "rbm_attn_last_query: bad mask mode %d",
              mask_mode);
  RBM_REQUIRE(mask_mode != RBM_MASK_KEYPAD || tok, "rbm_attn_last_query: the key-padding mask needs the token ids");
  RBM_REQUIRE(ldk % 4 == 0 && rbm_aligned16(k), "rbm_attn_last_query: key rows must be 16B aligned");
  if (B == 0) return 0;
  attn_last_query_kernel<<<(unsigned)(B * h), LQ_THREADS, 0, (cudaStream_t)stream>>>(q, ldq, k, ldk, v, ldv, tok, out, ldo, L, h, dk, mask_mode,
                                                                                      scale);
  RBM_LAUNCH_CHECK("rbm_attn_last_query");
  return 0;
}
