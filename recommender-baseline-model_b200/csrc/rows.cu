// rows.cu -- live-row compaction around the token-wise layers of SASRec.  The reference multiplies the activations by the timeline
// mask (`seqs *= ~timeline_mask`, NN/models/sas_model/sas.py:67,86) after the embedding and after every block: the rows of padding
// positions are exactly zero at every block input, their keys / values equal the projection bias, and whatever the block computes
// for them is multiplied by zero again -- value and gradient.  With Amazon-Beauty-like histories (mean length 9 of L = 50) 86 % of
// the rows are such rows.  LayerNorm, the Linear layers and the feed-forward therefore run on the live rows only:
//   rbm_rows_gather   [n, d] -> [cap, d]   compact row r = source row rows[r] (r < *count; optionally times coef[rows[r]]), zero rows after that
//   rbm_rows_scatter  [cap, d] -> [n, d]   row rows[r] = compact row r; rows of padding positions = fill (a [d] vector: the k/v bias,
//                                          beta of the last LayerNorm = what the reference computes there) or zero
//   rbm_rows_dead_colsum                   sum of the rows of padding positions (gradient of `fill`), fixed summation order
// rows / count come from rbm_compact_labels applied to the token ids; cap is a host-side capacity >= *count (fixed under a CUDA graph).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) rows_gather_kernel(const float* __restrict__ src, int64_t ld, const int32_t* __restrict__ rows,
                                                          const int32_t* __restrict__ count, int64_t cap, int d4, const float* __restrict__ coef,
                                                          float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap * d4) return;
  const int64_t r = i / d4;
  const int c = (int)(i - r * d4) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < *count) {
    v = ld4(src + (int64_t)rows[r] * ld + c);
    if (coef) {  // the product rbm_scatter_add_sorted would form from (coef, src) itself: same rounding
      const float cf = coef[rows[r]];
      v.x = __fmul_rn(cf, v.x); v.y = __fmul_rn(cf, v.y); v.z = __fmul_rn(cf, v.z); v.w = __fmul_rn(cf, v.w);
    }
  }
  st4(dst + r * (int64_t)d4 * 4 + c, v);
}

// one launch, two row ranges: [0, n) writes the fill into the rows of padding positions, [n, n + cap) copies the compact rows home
__global__ void __launch_bounds__(256) rows_scatter_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows,
                                                           const int32_t* __restrict__ count, int64_t cap, int d4, const float* __restrict__ fill,
                                                           const int64_t* __restrict__ tok, int64_t n, float* __restrict__ dst, int64_t ldd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (n + cap) * d4) return;
  const int64_t r = i / d4;
  const int c = (int)(i - r * d4) * 4;
  if (r < n) {
    if (tok[r] == 0) st4(dst + r * ldd + c, fill ? ld4(fill + c) : make_float4(0.f, 0.f, 0.f, 0.f));
  } else {
    const int64_t rc = r - n;
    if (rc < *count) st4(dst + (int64_t)rows[rc] * ldd + c, ld4(src + rc * (int64_t)d4 * 4 + c));
  }
}

// compact rows <-> a per-sequence padded layout [B, Lq, d]: slot o of sequence b = compact row seq_start[b] + o (zero rows past the
// sequence's last live row).  The caller guarantees that no sequence has more than Lq live rows.
__global__ void __launch_bounds__(256) rows_to_seq_kernel(const float* __restrict__ src, const int32_t* __restrict__ seq_start, int64_t nslots,
                                                          int Lq, int d4, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots * d4) return;
  const int64_t slot = i / d4;
  const int c = (int)(i - slot * d4) * 4;
  const int b = (int)(slot / Lq), o = (int)(slot - (int64_t)b * Lq);
  const int r = seq_start[b] + o;
  st4(dst + slot * (int64_t)d4 * 4 + c, r < seq_start[b + 1] ? ld4(src + (int64_t)r * d4 * 4 + c) : make_float4(0.f, 0.f, 0.f, 0.f));
}
__global__ void __launch_bounds__(256) seq_to_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ rows,
                                                          const int32_t* __restrict__ seq_start, const int32_t* __restrict__ count, int64_t cap,
                                                          int L, int Lq, int d4, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap * d4) return;
  const int64_t r = i / d4;
  const int c = (int)(i - r * d4) * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < *count) {
    const int b = rows[r] / L, o = (int)r - seq_start[b];
    if (o < Lq) v = ld4(src + ((int64_t)b * Lq + o) * d4 * 4 + c);
  }
  st4(dst + r * (int64_t)d4 * 4 + c, v);
}

constexpr int DC_BLOCKS = 8 * RBM_NUM_SMS;
// stage 1: block b sums the padding rows of its contiguous row range.  Thread t owns the float4 column group t % d4 of the rows
// r0 + t / d4, + 256 / d4, ... (ascending); the 256 / d4 row lanes are then added in lane order: a fixed summation tree.
// rows that count: tok[r] == 0 (the padding rows of a [B*L] layout), or, with tok == nullptr, the first *count rows (compact layout)
__global__ void __launch_bounds__(256) dead_colsum_partial_kernel(const float* __restrict__ src, int64_t ld, const int64_t* __restrict__ tok,
                                                                  const int32_t* __restrict__ count, int64_t n, int d, float* __restrict__ part) {
  __shared__ float4 red[256];
  const int d4 = d >> 2, nro = 256 / d4;  // d <= 1024: at least one row lane; threads past nro * d4 idle
  const int cg = threadIdx.x % d4, ro = threadIdx.x / d4;
  if (!tok && n > *count) n = *count;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = r0 + per < n ? r0 + per : n;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ro < nro)
    for (int64_t r = r0 + ro; r < r1; r += nro)
      if (!tok || tok[r] == 0) {
        const float4 v = ld4(src + r * ld + cg * 4);
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
  red[threadIdx.x] = s;
  __syncthreads();
  if (ro == 0) {
    for (int k = 1; k < nro; ++k) {
      const float4 v = red[k * d4 + cg];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    st4(part + (int64_t)blockIdx.x * d + cg * 4, s);
  }
}
// one block per column: strided partial sums per thread, then a fixed-shape tree in shared memory (deterministic)
__global__ void __launch_bounds__(256) dead_colsum_final_kernel(const float* __restrict__ part, int nb, int d, float* __restrict__ out) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int b = threadIdx.x; b < nb; b += 256) s += part[(int64_t)b * d + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = red[0];
}

}  // namespace

extern "C" int rbm_rows_gather(const float* src, int64_t ld, const int32_t* rows, const int32_t* count, int64_t cap, int d,
                               const float* coef, float* dst, rbm_stream_t stream) {
  RBM_REQUIRE(src && rows && count && dst, "rbm_rows_gather: null pointer");
  RBM_REQUIRE(cap >= 0 && d > 0 && d % 4 == 0 && ld % 4 == 0 && ld >= d, "rbm_rows_gather: need d %% 4 == 0 and ld >= d (d=%d)", d);
  RBM_REQUIRE(rbm_aligned16(src) && rbm_aligned16(dst), "rbm_rows_gather: pointers must be 16B aligned");
  if (cap == 0) return 0;
  const int64_t total = cap * (d / 4);
  rows_gather_kernel<<<(unsigned)rbm_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, ld, rows, count, cap, d / 4, coef, dst);
  RBM_LAUNCH_CHECK("rbm_rows_gather");
  return 0;
}

extern "C" int rbm_rows_scatter(const float* src, const int32_t* rows, const int32_t* count, int64_t cap, int d, const float* fill,
                                const int64_t* tok, int64_t n, float* dst, int64_t ldd, rbm_stream_t stream) {
  RBM_REQUIRE(src && rows && count && tok && dst, "rbm_rows_scatter: null pointer");
  RBM_REQUIRE(cap >= 0 && n >= 0 && d > 0 && d % 4 == 0 && ldd % 4 == 0 && ldd >= d, "rbm_rows_scatter: need d %% 4 == 0 and ldd >= d (d=%d)", d);
  RBM_REQUIRE(rbm_aligned16(src) && rbm_aligned16(dst) && rbm_aligned16(fill), "rbm_rows_scatter: pointers must be 16B aligned");
  if (n + cap == 0) return 0;
  const int64_t total = (n + cap) * (d / 4);
  rows_scatter_kernel<<<(unsigned)rbm_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, rows, count, cap, d / 4, fill, tok, n, dst, ldd);
  RBM_LAUNCH_CHECK("rbm_rows_scatter");
  return 0;
}

extern "C" size_t rbm_rows_dead_colsum_ws_bytes(int d) { return (size_t)DC_BLOCKS * d * sizeof(float); }

extern "C" int rbm_rows_dead_colsum(const float* src, int64_t ld, const int64_t* tok, int64_t n, int d, float* out, void* ws,
                                    size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(src && tok && out && ws, "rbm_rows_dead_colsum: null pointer");
  RBM_REQUIRE(n >= 0 && d >= 4 && d <= 1024 && d % 4 == 0 && ld >= d && ld % 4 == 0 && rbm_aligned16(src),
              "rbm_rows_dead_colsum: need d %% 4 == 0, d <= 1024 (d=%d)", d);
  RBM_REQUIRE(ws_bytes >= rbm_rows_dead_colsum_ws_bytes(d), "rbm_rows_dead_colsum: workspace too small");
  float* part = (float*)ws;
  dead_colsum_partial_kernel<<<DC_BLOCKS, 256, 0, (cudaStream_t)stream>>>(src, ld, tok, nullptr, n, d, part);
  RBM_LAUNCH_CHECK("rbm_rows_dead_colsum");
  dead_colsum_final_kernel<<<(unsigned)d, 256, 0, (cudaStream_t)stream>>>(part, DC_BLOCKS, d, out);
  RBM_LAUNCH_CHECK("rbm_rows_dead_colsum(final)");
  return 0;
}

// out[c] = sum of the first *count rows of src [cap, d] (compact layout), same fixed summation tree
extern "C" int rbm_rows_live_colsum(const float* src, int64_t ld, const int32_t* count, int64_t cap, int d, float* out, void* ws,
                                    size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(src && count && out && ws, "rbm_rows_live_colsum: null pointer");
  RBM_REQUIRE(cap >= 0 && d >= 4 && d <= 1024 && d % 4 == 0 && ld >= d && ld % 4 == 0 && rbm_aligned16(src),
              "rbm_rows_live_colsum: need d %% 4 == 0, d <= 1024 (d=%d)", d);
  RBM_REQUIRE(ws_bytes >= rbm_rows_dead_colsum_ws_bytes(d), "rbm_rows_live_colsum: workspace too small");
  float* part = (float*)ws;
  dead_colsum_partial_kernel<<<DC_BLOCKS, 256, 0, (cudaStream_t)stream>>>(src, ld, nullptr, count, cap, d, part);
  RBM_LAUNCH_CHECK("rbm_rows_live_colsum");
  dead_colsum_final_kernel<<<(unsigned)d, 256, 0, (cudaStream_t)stream>>>(part, DC_BLOCKS, d, out);
  RBM_LAUNCH_CHECK("rbm_rows_live_colsum(final)");
  return 0;
}

extern "C" int rbm_rows_to_seq(const float* src, const int32_t* seq_start, int B, int Lq, int d, float* dst, rbm_stream_t stream) {
  RBM_REQUIRE(src && seq_start && dst && B >= 0 && Lq >= 1 && d > 0 && d % 4 == 0, "rbm_rows_to_seq: bad arguments");
  RBM_REQUIRE(rbm_aligned16(src) && rbm_aligned16(dst), "rbm_rows_to_seq: pointers must be 16B aligned");
  const int64_t total = (int64_t)B * Lq * (d / 4);
  if (total == 0) return 0;
  rows_to_seq_kernel<<<(unsigned)rbm_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, seq_start, (int64_t)B * Lq, Lq, d / 4, dst);
  RBM_LAUNCH_CHECK("rbm_rows_to_seq");
  return 0;
}

extern "C" int rbm_seq_to_rows(const float* src, const int32_t* rows, const int32_t* seq_start, const int32_t* count, int64_t cap, int L,
                               int Lq, int d, float* dst, rbm_stream_t stream) {
  RBM_REQUIRE(src && rows && seq_start && count && dst && cap >= 0 && L >= 1 && Lq >= 1 && d > 0 && d % 4 == 0, "rbm_seq_to_rows: bad arguments");
  RBM_REQUIRE(rbm_aligned16(src) && rbm_aligned16(dst), "rbm_seq_to_rows: pointers must be 16B aligned");
  const int64_t total = cap * (d / 4);
  if (total == 0) return 0;
  seq_to_rows_kernel<<<(unsigned)rbm_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(src, rows, seq_start, count, cap, L, Lq, d / 4, dst);
  RBM_LAUNCH_CHECK("rbm_seq_to_rows");
  return 0;
}
