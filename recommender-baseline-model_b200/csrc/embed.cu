// embed.cu -- embedding stage: row gather (+scale) + positional add + dropout + pad-row zeroing, fwd/bwd.
// Replaces NN/models/sas_model/sas.py:60-67 and NN/models/bert_modules/embedding/bert.py:29-31.
// HBM-bound: one 16-byte load per thread from the table row, one from the positional row, one 16-byte store.
#include "common.cuh"

__global__ void __launch_bounds__(256) embed_fwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ table,
                                                        const float* __restrict__ pos, float* __restrict__ out,
                                                        int64_t total4, int L, int d4, int64_t vocab, int64_t v_begin,
                                                        int64_t v_end, float scale, int zero_pad, uint32_t thr,
                                                        float inv_keep, uint64_t seed, uint64_t site) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int64_t r = i / d4;
  int c4 = (int)(i - r * d4);
  int64_t t = tok[r];
  float4 o;
  if (zero_pad && t == 0) {
    o = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (t < 0 || t >= vocab) {
    float n = __int_as_float(0x7fc00000);  // out-of-range index: poison the row so it cannot go unnoticed
    o = make_float4(n, n, n, n);
  } else if (t < v_begin || t >= v_end) {
    o = make_float4(0.f, 0.f, 0.f, 0.f);  // row held by another shard: its owner writes the value, the exchange adds zeros
  } else {
    float4 e = ld4(table + (t - v_begin) * (int64_t)d4 * 4 + c4 * 4);
    float4 pe = ld4(pos + (int64_t)(r % L) * d4 * 4 + c4 * 4);
    o = make_float4(e.x * scale + pe.x, e.y * scale + pe.y, e.z * scale + pe.z, e.w * scale + pe.w);
    if (thr) {
      float4 m = rbm_drop4(seed, rbm_site(site), (uint64_t)i, thr, inv_keep);
      o.x *= m.x; o.y *= m.y; o.z *= m.z; o.w *= m.w;
    }
  }
  st4(out + i * 4, o);
}

// block (16,16): x = float4 column inside a 16-column group, y = batch lane; one block per (position l, group)
__global__ void __launch_bounds__(256) embed_bwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ dout,
                                                        float* __restrict__ g, float* __restrict__ dpos, int B, int L,
                                                        int d4, int zero_pad, uint32_t thr, float inv_keep,
                                                        uint64_t seed, uint64_t site, uint64_t elem_offset4) {
  __shared__ float4 red[16][16];
  int l = blockIdx.x;
  int c4 = blockIdx.y * 16 + threadIdx.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 < d4) {
    for (int b = threadIdx.y; b < B; b += 16) {
      int64_t r = (int64_t)b * L + l;
      int64_t i = r * d4 + c4;
      float4 v = ld4(dout + i * 4);
      if (zero_pad && tok[r] == 0) {
        v = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (thr) {
        float4 m = rbm_drop4(seed, rbm_site(site), (uint64_t)i + elem_offset4, thr, inv_keep);
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
      }
      st4(g + i * 4, v);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c4 < d4) {
    float4 s = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 16; ++y) {
      float4 t = red[y][threadIdx.x];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    st4(dpos + ((int64_t)l * d4 + c4) * 4, s);
  }
}

__global__ void dropout_mask_kernel(uint8_t* out, int64_t n, uint32_t thr, uint64_t seed, uint64_t site) {
  int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  uint4 r = rbm_philox(seed, site, (uint64_t)i4);
  uint32_t v[4] = {r.x, r.y, r.z, r.w};
  for (int c = 0; c < 4; ++c)
    if (i4 * 4 + c < n) out[i4 * 4 + c] = v[c] >= thr ? 1 : 0;
}

__global__ void dropout_mask_attn_kernel(uint8_t* out, int64_t rows, int L, uint32_t thr16, uint64_t seed, uint64_t site) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * L) return;
  int64_t R = e / L;  // R = bh * L + i
  int j = (int)(e - R * L);
  out[e] = rbm_attn_keep(seed, site, (uint64_t)(R / L), (int)(R % L), j, thr16) ? 1 : 0;
}

extern "C" int rbm_embed_fwd_shard(const int64_t* tok, const float* table_shard, const float* pos, float* out, int64_t rows, int L,
                                   int d, int64_t vocab, int64_t v_begin, int64_t v_end, float scale, int zero_pad, float p,
                                   uint64_t seed, uint64_t site, rbm_stream_t stream);

extern "C" int rbm_embed_fwd(const int64_t* tok, const float* table, const float* pos, float* out, int64_t rows, int L,
                             int d, int64_t vocab, float scale, int zero_pad, float p, uint64_t seed, uint64_t site,
                             rbm_stream_t stream) {
  return rbm_embed_fwd_shard(tok, table, pos, out, rows, L, d, vocab, 0, vocab, scale, zero_pad, p, seed, site, stream);
}

extern "C" int rbm_embed_fwd_shard(const int64_t* tok, const float* table, const float* pos, float* out, int64_t rows, int L,
                                   int d, int64_t vocab, int64_t v_begin, int64_t v_end, float scale, int zero_pad, float p,
                                   uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(tok && table && pos && out, "rbm_embed_fwd: null pointer");
  RBM_REQUIRE(v_begin >= 0 && v_begin <= v_end && v_end <= vocab, "rbm_embed_fwd: bad shard range [%lld, %lld) of %lld rows",
              (long long)v_begin, (long long)v_end, (long long)vocab);
  RBM_REQUIRE(d > 0 && d % 4 == 0, "rbm_embed_fwd: d=%d must be a positive multiple of 4", d);
  RBM_REQUIRE(L > 0 && rows >= 0 && rows % L == 0, "rbm_embed_fwd: rows=%lld not a multiple of L=%d", (long long)rows, L);
  RBM_REQUIRE(p >= 0.f && p < 1.f, "rbm_embed_fwd: dropout p=%f out of [0,1)", p);
  RBM_REQUIRE(rbm_aligned16(table) && rbm_aligned16(pos) && rbm_aligned16(out), "rbm_embed_fwd: pointers must be 16B aligned");
  if (rows == 0) return 0;
  int64_t total4 = rows * (d / 4);
  uint32_t thr = rbm_drop_threshold(p);
  embed_fwd_kernel<<<(unsigned)rbm_cdiv(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      tok, table, pos, out, total4, L, d / 4, vocab, v_begin, v_end, scale, zero_pad, thr, 1.f / (1.f - p), seed, site);
  RBM_LAUNCH_CHECK("rbm_embed_fwd");
  return 0;
}

extern "C" int rbm_embed_bwd_offset(const int64_t* tok, const float* dout, float* g, float* dpos, int64_t rows, int L, int d,
                                    int zero_pad, float p, uint64_t seed, uint64_t site, uint64_t row_offset, rbm_stream_t stream);

extern "C" int rbm_embed_bwd(const int64_t* tok, const float* dout, float* g, float* dpos, int64_t rows, int L, int d,
                             int zero_pad, float p, uint64_t seed, uint64_t site, rbm_stream_t stream) {
  return rbm_embed_bwd_offset(tok, dout, g, dpos, rows, L, d, zero_pad, p, seed, site, 0, stream);
}

extern "C" int rbm_embed_bwd_offset(const int64_t* tok, const float* dout, float* g, float* dpos, int64_t rows, int L, int d,
                                    int zero_pad, float p, uint64_t seed, uint64_t site, uint64_t row_offset, rbm_stream_t stream) {
  RBM_REQUIRE(tok && dout && g && dpos, "rbm_embed_bwd: null pointer");
  RBM_REQUIRE(d > 0 && d % 4 == 0, "rbm_embed_bwd: d=%d must be a positive multiple of 4", d);
  RBM_REQUIRE(L > 0 && rows > 0 && rows % L == 0, "rbm_embed_bwd: rows=%lld not a positive multiple of L=%d", (long long)rows, L);
  RBM_REQUIRE(p >= 0.f && p < 1.f, "rbm_embed_bwd: dropout p=%f out of [0,1)", p);
  RBM_REQUIRE(rbm_aligned16(dout) && rbm_aligned16(g) && rbm_aligned16(dpos), "rbm_embed_bwd: pointers must be 16B aligned");
  int d4 = d / 4;
  dim3 grid(L, (unsigned)rbm_cdiv(d4, 16)), block(16, 16);
  embed_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(tok, dout, g, dpos, (int)(rows / L), L, d4, zero_pad,
                                                             rbm_drop_threshold(p), 1.f / (1.f - p), seed, site, row_offset * (uint64_t)d4);
  RBM_LAUNCH_CHECK("rbm_embed_bwd");
  return 0;
}

extern "C" int rbm_dropout_mask(uint8_t* out, int64_t n, float p, uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(out && n >= 0 && p >= 0.f && p < 1.f, "rbm_dropout_mask: bad argument");
  if (n == 0) return 0;
  dropout_mask_kernel<<<(unsigned)rbm_cdiv(rbm_cdiv(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(out, n, rbm_drop_threshold(p), seed, site);
  RBM_LAUNCH_CHECK("rbm_dropout_mask");
  return 0;
}

extern "C" int rbm_dropout_mask_attn(uint8_t* out, int64_t rows, int L, float p, uint64_t seed, uint64_t site,
                                     rbm_stream_t stream) {
  RBM_REQUIRE(out && rows >= 0 && L > 0 && L <= 256 && p >= 0.f && p < 1.f, "rbm_dropout_mask_attn: bad argument");
  if (rows == 0) return 0;
  dropout_mask_attn_kernel<<<(unsigned)rbm_cdiv(rows * L, 256), 256, 0, (cudaStream_t)stream>>>(out, rows, L, rbm_drop_threshold16(p), seed, site);
  RBM_LAUNCH_CHECK("rbm_dropout_mask_attn");
  return 0;
}

RBM_DEFINE_STEP_PTR_SETTER(rbm_step_ptr_set_embed)
