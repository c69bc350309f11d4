// score_ce.cu -- BERT4Rec output scoring fused with masked cross-entropy; logits are never written to HBM.
// Replaces  logits = self.out(h)  (NN/models/bert.py:16, materialises [B,L,V+1]) followed by
// CrossEntropyLoss(ignore_index=0) (NN/trainers/bert.py:11,36-40) and their autograd.
//   fwd : per compacted (labels != 0) row, stream 64-wide vocab tiles, online (max, sum-exp), capture target logit.
//   bwd : logits are recomputed tile-wise;  dH = G.W  (CTA per row tile),  dW = G^T.H, db = colsum(G) (CTA per
//         vocab tile x row split, fixed-order split reduction)  with  G = (softmax - onehot) * dloss / count.
// Round-1 implementation: register-tiled SIMT fp32 (exact-parity baseline for the tcgen05 version).
#include "common.cuh"

namespace {

constexpr int T = 64;        // tile edge (rows and vocab columns)
constexpr int LDT = T + 4;   // padded smem row length (16B aligned)
constexpr int KC = 16;       // k-chunk for streamed operands

// ---------------------------------------------------------------------------------------------- compaction
__global__ void __launch_bounds__(256) compact_count_kernel(const int64_t* __restrict__ labels, int64_t n, int32_t* __restrict__ blk_cnt) {
  __shared__ int wc[8];
  int64_t base = (int64_t)blockIdx.x * 1024;
  int c = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int64_t i = base + t * 256 + threadIdx.x;
    c += (i < n && labels[i] != 0) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += wc[w];
    blk_cnt[blockIdx.x] = s;
  }
}

__global__ void compact_scan_kernel(int32_t* blk_cnt, int nblk, int32_t* count_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < nblk; ++b) {
      int c = blk_cnt[b];
      blk_cnt[b] = run;
      run += c;
    }
    *count_out = run;
  }
}

__global__ void __launch_bounds__(256) compact_write_kernel(const int64_t* __restrict__ labels, int64_t n, const int32_t* __restrict__ blk_off,
                                                            int32_t* __restrict__ rows_out, int64_t* __restrict__ tgt_out) {
  __shared__ int wc[8];
  int64_t base = (int64_t)blockIdx.x * 1024;
  int run = blk_off[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = 0; t < 4; ++t) {
    int64_t i = base + t * 256 + threadIdx.x;
    int64_t lab = i < n ? labels[i] : 0;
    bool f = lab != 0;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) wc[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      int c = wc[w];
      if (w < warp) woff += c;
      tot += c;
    }
    if (f) {
      int pos = run + woff + __popc(bal & ((1u << lane) - 1u));
      rows_out[pos] = (int32_t)i;
      tgt_out[pos] = lab;
    }
    run += tot;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------ shared helpers
// Hst[k][r] <- h[rows[r0+r], k]  (transposed; rows >= count zero-filled)
__device__ __forceinline__ void load_h_transposed(float* Hst, const float* __restrict__ h, const int32_t* __restrict__ rows,
                                                  int r0, int count, int d) {
  int d4 = d >> 2;
  for (int idx = threadIdx.x; idx < T * d4; idx += blockDim.x) {
    int r = idx / d4, c4 = idx - r * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < count) v = ld4(h + (int64_t)rows[r0 + r] * d + c4 * 4);
    Hst[(c4 * 4 + 0) * LDT + r] = v.x;
    Hst[(c4 * 4 + 1) * LDT + r] = v.y;
    Hst[(c4 * 4 + 2) * LDT + r] = v.z;
    Hst[(c4 * 4 + 3) * LDT + r] = v.w;
  }
}
// Xrow[r][k] (row stride ldr) <- src rows; gather via rows[] when given, zero-fill invalid
__device__ __forceinline__ void load_rows_rowmajor(float* X, int ldr, const float* __restrict__ src, const int32_t* __restrict__ rows,
                                                   int64_t r0, int64_t limit, int d) {
  int d4 = d >> 2;
  for (int idx = threadIdx.x; idx < T * d4; idx += blockDim.x) {
    int r = idx / d4, c4 = idx - r * d4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < limit) {
      int64_t gr = rows ? (int64_t)rows[r0 + r] : (r0 + r);
      v = ld4(src + gr * d + c4 * 4);
    }
    st4(X + r * ldr + c4 * 4, v);
  }
}
// Wst chunk [KC][LDT] <- w[v0 + c, k0 + kk]  (transposed), 256 threads, one float4 each
__device__ __forceinline__ void load_w_chunk(float* Wc, const float* __restrict__ w, int v0, int V1, int k0, int d) {
  int c = threadIdx.x >> 2, kq = threadIdx.x & 3;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (v0 + c < V1 && k0 + kq * 4 < d) v = ld4(w + (int64_t)(v0 + c) * d + k0 + kq * 4);
  Wc[(kq * 4 + 0) * LDT + c] = v.x;
  Wc[(kq * 4 + 1) * LDT + c] = v.y;
  Wc[(kq * 4 + 2) * LDT + c] = v.z;
  Wc[(kq * 4 + 3) * LDT + c] = v.w;
}

#define FMA44(acc, a, b)                                                                       \
  do {                                                                                         \
    float _a[4] = {a.x, a.y, a.z, a.w}, _b[4] = {b.x, b.y, b.z, b.w};                          \
    _Pragma("unroll") for (int _i = 0; _i < 4; ++_i) _Pragma("unroll") for (int _j = 0; _j < 4; ++_j) acc[_i][_j] = \
        fmaf(_a[_i], _b[_j], acc[_i][_j]);                                                     \
  } while (0)

__device__ __forceinline__ float half_sum(float v) {  // across the 16 tx lanes that share a row
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// logits tile acc[4][4] (rows ty*4+i, cols tx*4+j) = Hst^T . W[v0:v0+64]^T, W streamed in KC-chunks through Wc
__device__ __forceinline__ void logits_tile_streamW(float (&acc)[4][4], const float* Hst, float* Wc, const float* __restrict__ w,
                                                    int v0, int V1, int d, int tx, int ty) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < d; k0 += KC) {
    __syncthreads();
    load_w_chunk(Wc, w, v0, V1, k0, d);
    __syncthreads();
    int kmax = d - k0 < KC ? d - k0 : KC;
    for (int kk = 0; kk < kmax; ++kk) {
      float4 a = ld4(Hst + (k0 + kk) * LDT + ty * 4);
      float4 b = ld4(Wc + kk * LDT + tx * 4);
      FMA44(acc, a, b);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ h, const int32_t* __restrict__ rows,
                                                     const int64_t* __restrict__ tgt, const int32_t* __restrict__ count_p,
                                                     const float* __restrict__ w, const float* __restrict__ bias,
                                                     float* __restrict__ lse_out, float* __restrict__ partial, int V1, int d) {
  extern __shared__ __align__(16) float sm[];
  float* Hst = sm;             // [d][LDT]
  float* Wc = Hst + d * LDT;   // [KC][LDT]
  float* contrib = Wc + KC * LDT;  // [T]
  const int count = *count_p;
  const int r0 = blockIdx.x * T;
  if (r0 >= count) {
    if (threadIdx.x == 0) partial[blockIdx.x] = 0.f;
    return;
  }
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  load_h_transposed(Hst, h, rows, r0, count, d);
  int64_t rt[4];
  float m[4], l[4], tl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    rt[i] = r < count ? tgt[r] : -1;
    m[i] = -INFINITY;
    l[i] = 0.f;
    tl[i] = 0.f;
  }
  for (int v0 = 0; v0 < V1; v0 += T) {
    float acc[4][4];
    logits_tile_streamW(acc, Hst, Wc, w, v0, V1, d, tx, ty);
    int cb = v0 + tx * 4;
    float bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = (bias && cb + j < V1) ? bias[cb + j] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float lg[4], tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        lg[j] = cb + j < V1 ? acc[i][j] + bv[j] : -INFINITY;
        tmax = fmaxf(tmax, lg[j]);
        if ((int64_t)(cb + j) == rt[i]) tl[i] += lg[j];
      }
      tmax = half_max(tmax);
      float mn = fmaxf(m[i], tmax);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) ps += expf(lg[j] - mn);
      ps = half_sum(ps);
      l[i] = l[i] * expf(m[i] - mn) + ps;
      m[i] = mn;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t = half_sum(tl[i]);
    int r = r0 + ty * 4 + i;
    float lse = m[i] + logf(l[i]);
    if (tx == 0) {
      if (r < count) lse_out[r] = lse;
      contrib[ty * 4 + i] = r < count ? lse - t : 0.f;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int r = 0; r < T; ++r) s += contrib[r];
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) ce_loss_finalize_kernel(const float* __restrict__ partial, int nblk,
                                                               const int32_t* __restrict__ count_p, float* __restrict__ loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int b = threadIdx.x; b < nblk; b += 256) s += partial[b];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += red[i];
    *loss = t / (float)(*count_p);
  }
}

// ---------------------------------------------------------------------------------------- backward: dH
template <int NT2>
__global__ void __launch_bounds__(256) ce_bwd_dh_kernel(const float* __restrict__ h, const int32_t* __restrict__ rows,
                                                        const int64_t* __restrict__ tgt, const int32_t* __restrict__ count_p,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ lse, const float* __restrict__ dloss,
                                                        float* __restrict__ dh_full, int V1, int d) {
  extern __shared__ __align__(16) float sm[];
  const int ldw = d + 4;
  float* Hst = sm;               // [d][LDT]
  float* Wc = Hst + d * LDT;     // [KC][LDT]
  float* Gst = Wc + KC * LDT;    // [T cols][LDT]  G transposed: Gst[c][r]
  float* Wrow = Gst + T * LDT;   // [T cols][ldw]  row-major W tile
  const int count = *count_p;
  const int r0 = blockIdx.x * T;
  if (r0 >= count) return;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float gscale = *dloss / (float)count;
  load_h_transposed(Hst, h, rows, r0, count, d);
  int64_t rt[4];
  float rl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    rt[i] = r < count ? tgt[r] : -1;
    rl[i] = r < count ? lse[r] : 0.f;
  }
  float acc2[NT2][4][4];
#pragma unroll
  for (int t = 0; t < NT2; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[t][i][j] = 0.f;

  for (int v0 = 0; v0 < V1; v0 += T) {
    float acc[4][4];
    logits_tile_streamW(acc, Hst, Wc, w, v0, V1, d, tx, ty);  // starts with __syncthreads(): previous GEMM2 finished
    load_rows_rowmajor(Wrow, ldw, w, nullptr, v0, V1, d);
    int cb = v0 + tx * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float bj = (bias && cb + j < V1) ? bias[cb + j] : 0.f;
      float g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float p = (cb + j < V1 && rt[i] >= 0) ? expf(acc[i][j] + bj - rl[i]) : 0.f;
        if ((int64_t)(cb + j) == rt[i]) p -= 1.f;
        g[i] = p * gscale;
      }
      st4(Gst + (tx * 4 + j) * LDT + ty * 4, make_float4(g[0], g[1], g[2], g[3]));
    }
    __syncthreads();
    // GEMM2: dH[r][kk] += sum_c G[r][c] * W[c][kk];  thread owns rows ty*4.., cols tx*4 + 64*t ..
    for (int c = 0; c < T; ++c) {
      float4 a = ld4(Gst + c * LDT + ty * 4);
#pragma unroll
      for (int t = 0; t < NT2; ++t) {
        int kk = tx * 4 + 64 * t;
        if (kk < d) {
          float4 b = ld4(Wrow + c * ldw + kk);
          FMA44(acc2[t], a, b);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = r0 + ty * 4 + i;
    if (r < count) {
#pragma unroll
      for (int t = 0; t < NT2; ++t) {
        int kk = tx * 4 + 64 * t;
        if (kk < d) st4(dh_full + (int64_t)rows[r] * d + kk, make_float4(acc2[t][i][0], acc2[t][i][1], acc2[t][i][2], acc2[t][i][3]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------ backward: dW and db
// grid (vocab tiles, S): CTA owns vocab rows [v0, v0+64) and row tiles  {s, s+S, s+2S, ...}.
template <int NT2>
__global__ void __launch_bounds__(256) ce_bwd_dw_kernel(const float* __restrict__ h, const int32_t* __restrict__ rows,
                                                        const int64_t* __restrict__ tgt, const int32_t* __restrict__ count_p,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ lse, const float* __restrict__ dloss,
                                                        float* __restrict__ part_w, float* __restrict__ part_b, int V1, int d) {
  extern __shared__ __align__(16) float sm[];
  const int ldh = d + 4;
  float* Wst = sm;               // [d][LDT]   fixed vocab tile, transposed: Wst[k][c]
  float* Hrow = Wst + d * LDT;   // [T rows][ldh] row-major gathered H tile
  float* Gs = Hrow + T * ldh;    // [T rows][LDT]  G[r][c]
  const int count = *count_p;
  const int v0 = blockIdx.x * T;
  const int S = gridDim.y, s = blockIdx.y;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float gscale = *dloss / (float)count;
  // W tile transposed (zero beyond V1)
  {
    int d4 = d >> 2;
    for (int idx = threadIdx.x; idx < T * d4; idx += blockDim.x) {
      int c = idx / d4, k4 = idx - c * d4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + c < V1) v = ld4(w + (int64_t)(v0 + c) * d + k4 * 4);
      Wst[(k4 * 4 + 0) * LDT + c] = v.x;
      Wst[(k4 * 4 + 1) * LDT + c] = v.y;
      Wst[(k4 * 4 + 2) * LDT + c] = v.z;
      Wst[(k4 * 4 + 3) * LDT + c] = v.w;
    }
  }
  const int cb = v0 + tx * 4;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (bias && cb + j < V1) ? bias[cb + j] : 0.f;
  float acc2[NT2][4][4];  // dW rows (vocab) ty*4+i, cols tx*4 + 64*t + j
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < NT2; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[t][i][j] = 0.f;

  const int ntiles = (count + T - 1) / T;
  for (int rt_ = s; rt_ < ntiles; rt_ += S) {
    const int r0 = rt_ * T;
    __syncthreads();  // previous iteration's GEMM2 is done with Hrow / Gs
    load_rows_rowmajor(Hrow, ldh, h, rows, r0, count, d);
    __syncthreads();
    // GEMM1: logits[r][c] for rows ty*4+i, cols tx*4+j
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < d; ++k) {
      float4 b = ld4(Wst + k * LDT + tx * 4);
      float4 a = make_float4(Hrow[(ty * 4 + 0) * ldh + k], Hrow[(ty * 4 + 1) * ldh + k], Hrow[(ty * 4 + 2) * ldh + k],
                             Hrow[(ty * 4 + 3) * ldh + k]);
      FMA44(acc, a, b);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int r = r0 + ty * 4 + i;
      int64_t tg = r < count ? tgt[r] : -1;
      float rl = r < count ? lse[r] : 0.f;
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = (cb + j < V1 && tg >= 0) ? expf(acc[i][j] + bv[j] - rl) : 0.f;
        if ((int64_t)(cb + j) == tg) p -= 1.f;
        g[j] = p * gscale;
      }
      st4(Gs + (ty * 4 + i) * LDT + tx * 4, make_float4(g[0], g[1], g[2], g[3]));
    }
    __syncthreads();
    // GEMM2: dW[c][kk] += sum_r G[r][c] * H[r][kk];  thread owns vocab rows ty*4.., cols tx*4 + 64*t ..
    for (int r = 0; r < T; ++r) {
      float4 a = ld4(Gs + r * LDT + ty * 4);
      bsum[0] += a.x; bsum[1] += a.y; bsum[2] += a.z; bsum[3] += a.w;
#pragma unroll
      for (int t = 0; t < NT2; ++t) {
        int kk = tx * 4 + 64 * t;
        if (kk < d) {
          float4 b = ld4(Hrow + r * ldh + kk);
          FMA44(acc2[t], a, b);
        }
      }
    }
  }
  float* pw = part_w + (int64_t)s * V1 * d;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = v0 + ty * 4 + i;
    if (c < V1) {
#pragma unroll
      for (int t = 0; t < NT2; ++t) {
        int kk = tx * 4 + 64 * t;
        if (kk < d) st4(pw + (int64_t)c * d + kk, make_float4(acc2[t][i][0], acc2[t][i][1], acc2[t][i][2], acc2[t][i][3]));
      }
      if (tx == 0) part_b[(int64_t)s * V1 + c] = bsum[i];
    }
  }
}

__global__ void ce_reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int S) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < S; ++k) s += part[(int64_t)k * n + i];
  out[i] = s;
}

int dw_splits(int64_t cap, int V1) {
  int64_t vt = rbm_cdiv(V1, T), rtiles = rbm_cdiv(cap, T);
  int64_t s = rbm_cdiv((int64_t)RBM_NUM_SMS * 4, vt);
  if (s > rtiles) s = rtiles;
  return (int)(s < 1 ? 1 : s);
}
int nt2_of(int d) { return d <= 64 ? 1 : d <= 128 ? 2 : 4; }

}  // namespace

extern "C" size_t rbm_compact_ws_bytes(int64_t n) { return (size_t)rbm_cdiv(n, 1024) * sizeof(int32_t) + 16; }

extern "C" int rbm_compact_labels(const int64_t* labels, int64_t n, int32_t* rows_out, int64_t* tgt_out, int32_t* count_out,
                                  void* ws, size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(labels && rows_out && tgt_out && count_out && ws, "rbm_compact_labels: null pointer");
  RBM_REQUIRE(n > 0 && n < (int64_t)1 << 31, "rbm_compact_labels: n=%lld out of range", (long long)n);
  RBM_REQUIRE(ws_bytes >= rbm_compact_ws_bytes(n), "rbm_compact_labels: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = (int)rbm_cdiv(n, 1024);
  int32_t* blk = (int32_t*)ws;
  compact_count_kernel<<<nblk, 256, 0, st>>>(labels, n, blk);
  compact_scan_kernel<<<1, 32, 0, st>>>(blk, nblk, count_out);
  compact_write_kernel<<<nblk, 256, 0, st>>>(labels, n, blk, rows_out, tgt_out);
  RBM_LAUNCH_CHECK("rbm_compact_labels");
  return 0;
}

extern "C" size_t rbm_ce_ws_bytes(int64_t cap, int V1, int d) {
  size_t nblk = (size_t)rbm_cdiv(cap, T);
  size_t S = (size_t)dw_splits(cap, V1);
  return (nblk + S * ((size_t)V1 * d + V1)) * sizeof(float) + 64;
}

static int ce_check(const char* name, int64_t cap, int V1, int d) {
  RBM_REQUIRE(cap > 0 && cap < (int64_t)1 << 31 && V1 > 0, "%s: bad sizes cap=%lld V1=%d", name, (long long)cap, V1);
  RBM_REQUIRE(d >= 4 && d % 4 == 0 && d <= 256, "%s: unsupported hidden size d=%d (need d%%4==0, d<=256)", name, d);
  return 0;
}

extern "C" int rbm_ce_fwd(const float* h, const int32_t* rows, const int64_t* tgt, const int32_t* count, const float* w,
                          const float* bias, float* lse, float* loss, int64_t cap, int V1, int d, void* ws, size_t ws_bytes,
                          rbm_stream_t stream) {
  RBM_REQUIRE(h && rows && tgt && count && w && lse && loss && ws, "rbm_ce_fwd: null pointer");
  if (ce_check("rbm_ce_fwd", cap, V1, d)) return -1;
  RBM_REQUIRE(ws_bytes >= rbm_ce_ws_bytes(cap, V1, d), "rbm_ce_fwd: workspace too small");
  RBM_REQUIRE(rbm_aligned16(h) && rbm_aligned16(w) && rbm_aligned16(ws), "rbm_ce_fwd: pointers must be 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = (int)rbm_cdiv(cap, T);
  size_t smem = sizeof(float) * ((size_t)d * LDT + KC * LDT + T);
  cudaFuncSetAttribute(ce_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ce_fwd_kernel<<<nblk, 256, smem, st>>>(h, rows, tgt, count, w, bias, lse, (float*)ws, V1, d);
  RBM_LAUNCH_CHECK("rbm_ce_fwd");
  ce_loss_finalize_kernel<<<1, 256, 0, st>>>((const float*)ws, nblk, count, loss);
  RBM_LAUNCH_CHECK("rbm_ce_fwd(finalize)");
  return 0;
}

extern "C" int rbm_ce_bwd(const float* h, const int32_t* rows, const int64_t* tgt, const int32_t* count, const float* w,
                          const float* bias, const float* lse, const float* dloss, float* dh_full, float* dw, float* db,
                          int64_t cap, int V1, int d, void* ws, size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(h && rows && tgt && count && w && lse && dloss && dh_full && dw && db && ws, "rbm_ce_bwd: null pointer");
  if (ce_check("rbm_ce_bwd", cap, V1, d)) return -1;
  RBM_REQUIRE(ws_bytes >= rbm_ce_ws_bytes(cap, V1, d), "rbm_ce_bwd: workspace too small");
  RBM_REQUIRE(rbm_aligned16(h) && rbm_aligned16(w) && rbm_aligned16(ws) && rbm_aligned16(dh_full) && rbm_aligned16(dw),
              "rbm_ce_bwd: pointers must be 16B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = (int)rbm_cdiv(cap, T);
  int S = dw_splits(cap, V1);
  int nt2 = nt2_of(d);
  size_t smem_dh = sizeof(float) * ((size_t)d * LDT + KC * LDT + T * LDT + (size_t)T * (d + 4));
  size_t smem_dw = sizeof(float) * ((size_t)d * LDT + (size_t)T * (d + 4) + T * LDT);
  float* part_w = (float*)ws + nblk;
  part_w = (float*)(((uintptr_t)part_w + 15) & ~(uintptr_t)15);
  float* part_b = part_w + (size_t)S * V1 * d;
  dim3 gdw((unsigned)rbm_cdiv(V1, T), S);
#define CE_BWD(NT2)                                                                                                   \
  do {                                                                                                                \
    cudaFuncSetAttribute(ce_bwd_dh_kernel<NT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dh);           \
    ce_bwd_dh_kernel<NT2><<<nblk, 256, smem_dh, st>>>(h, rows, tgt, count, w, bias, lse, dloss, dh_full, V1, d);      \
    cudaFuncSetAttribute(ce_bwd_dw_kernel<NT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw);           \
    ce_bwd_dw_kernel<NT2><<<gdw, 256, smem_dw, st>>>(h, rows, tgt, count, w, bias, lse, dloss, part_w, part_b, V1, d); \
  } while (0)
  if (nt2 == 1) CE_BWD(1);
  else if (nt2 == 2) CE_BWD(2);
  else CE_BWD(4);
#undef CE_BWD
  RBM_LAUNCH_CHECK("rbm_ce_bwd");
  int64_t n = (int64_t)V1 * d;
  ce_reduce_splits_kernel<<<(unsigned)rbm_cdiv(n, 256), 256, 0, st>>>(part_w, dw, n, S);
  ce_reduce_splits_kernel<<<(unsigned)rbm_cdiv(V1, 256), 256, 0, st>>>(part_b, db, V1, S);
  RBM_LAUNCH_CHECK("rbm_ce_bwd(reduce)");
  return 0;
}
