// scatter.cu -- embedding-gradient scatter-add as stable sort + sequential segment reduce (bit-deterministic).
// Replaces embedding_dense_backward (autograd of nn.Embedding: NN/models/sas_model/sas.py:30,
// NN/models/bert_modules/embedding/token.py:6; triggered by loss.backward() NN/trainers/base.py:121), which on
// stock CUDA uses float atomics (run-to-run different) and on CPU a serial index-ordered loop.  Here:
//   1. stable LSD radix sort (8-bit digits) of (row id, position) pairs -- positions stay ascending per row id;
//   2. segment reduce in fixed 64-contribution pieces (see below): rows with <= 64 contributions are summed in
//      ascending position == the CPU reference's order (bit-identical); longer rows follow a fixed piece tree.
//      Either way the result is run-to-run bit-stable (no atomics).
#include "common.cuh"

namespace {

constexpr int RADIX = 256;
constexpr int SORT_WARPS = 8;
constexpr int PER_WARP = 256;                 // elements per warp per tile (8 rounds of 32)
constexpr int TILE = SORT_WARPS * PER_WARP;   // 2048

__device__ __forceinline__ uint32_t load_key(const int64_t* idx, const uint32_t* keys_in, int64_t i) {
  return idx ? (uint32_t)idx[i] : keys_in[i];
}

// per-warp digit histogram of the warp's 256 elements -> wh[warp][256]
__device__ __forceinline__ void warp_histogram(const int64_t* idx, const uint32_t* keys_in, int64_t n, int shift, int64_t wbase,
                                               int lane, uint32_t* wh_w) {
  for (int b = lane; b < RADIX; b += 32) wh_w[b] = 0;
  __syncwarp();
  for (int r = 0; r < PER_WARP / 32; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool in = i < n;
    uint32_t dig = in ? (load_key(idx, keys_in, i) >> shift) & (RADIX - 1) : RADIX;  // RADIX = "none"
    unsigned m = __match_any_sync(0xffffffffu, dig);
    if (in && (__ffs(m) - 1) == lane) wh_w[dig] += __popc(m);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(32 * SORT_WARPS) radix_hist_kernel(const int64_t* __restrict__ idx, const uint32_t* __restrict__ keys_in,
                                                                      int64_t n, int shift, uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t wh[SORT_WARPS][RADIX];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  warp_histogram(idx, keys_in, n, shift, (int64_t)blockIdx.x * TILE + warp * PER_WARP, lane, wh[warp]);
  __syncthreads();
  int b = threadIdx.x;  // 256 threads <-> 256 bins
  uint32_t s = 0;
#pragma unroll
  for (int w = 0; w < SORT_WARPS; ++w) s += wh[w][b];
  hist[(int64_t)b * nblk + blockIdx.x] = s;
}

// exclusive scan of E = 256*nblk counters (bin-major) in one block of 1024 threads
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* __restrict__ hist, int64_t E) {
  __shared__ uint32_t sums[1024];
  int t = threadIdx.x;
  int64_t chunk = (E + 1023) / 1024;
  int64_t b = t * chunk, e = b + chunk < E ? b + chunk : E;
  uint32_t s = 0;
  for (int64_t i = b; i < e; ++i) s += hist[i];
  sums[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    uint32_t v = t >= o ? sums[t - o] : 0;
    __syncthreads();
    sums[t] += v;
    __syncthreads();
  }
  uint32_t run = sums[t] - s;  // exclusive prefix of this thread's chunk
  for (int64_t i = b; i < e; ++i) {
    uint32_t c = hist[i];
    hist[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(32 * SORT_WARPS) radix_scatter_kernel(const int64_t* __restrict__ idx, const uint32_t* __restrict__ keys_in,
                                                                         const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                                         uint32_t* __restrict__ vals_out, int64_t n, int shift,
                                                                         const uint32_t* __restrict__ hist, int nblk) {
  __shared__ uint32_t wh[SORT_WARPS][RADIX];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wbase = (int64_t)blockIdx.x * TILE + warp * PER_WARP;
  warp_histogram(idx, keys_in, n, shift, wbase, lane, wh[warp]);
  __syncthreads();
  {
    int b = threadIdx.x;
    uint32_t run = hist[(int64_t)b * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      uint32_t c = wh[w][b];
      wh[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
  uint32_t* off = wh[warp];
  for (int r = 0; r < PER_WARP / 32; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool in = i < n;
    uint32_t key = in ? load_key(idx, keys_in, i) : 0;
    uint32_t val = in ? (vals_in ? vals_in[i] : (uint32_t)i) : 0;
    uint32_t dig = in ? (key >> shift) & (RADIX - 1) : RADIX;
    unsigned m = __match_any_sync(0xffffffffu, dig);
    if (in) {
      uint32_t pos = off[dig] + __popc(m & ((1u << lane) - 1u));
      keys_out[pos] = key;
      vals_out[pos] = val;
    }
    __syncwarp();
    if (in && (__ffs(m) - 1) == lane) off[dig] += __popc(m);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Segment reduce.  A segment = the sorted slots of one destination row.  It is cut into PIECES of SEG_CHUNK
// consecutive contributions counted from the segment start; a piece is summed sequentially (ascending position),
// the piece sums of a segment are then added in piece order.  Segments of <= SEG_CHUNK contributions are therefore
// bit-identical to the CPU reference's index-ordered loop; longer ones (popular items, [MASK]) follow this fixed
// tree -- deterministic, and restated by oracle.common.embedding_grad_scatter_chunked.
// ---------------------------------------------------------------------------------------------------------
constexpr int SEG_CHUNK = 64;

__device__ __forceinline__ int64_t lower_bound_u32(const uint32_t* __restrict__ keys, int64_t n, uint32_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float4 load_contrib(const float* __restrict__ src, const float* __restrict__ coef, float alpha,
                                               uint32_t row, int d4, int c4) {
  float4 v = ld4(src + ((int64_t)row * d4 + c4) * 4);
  if (coef) {
    float cf = coef[row];
    v.x = __fmul_rn(cf, v.x); v.y = __fmul_rn(cf, v.y); v.z = __fmul_rn(cf, v.z); v.w = __fmul_rn(cf, v.w);
  }
  if (alpha != 1.f) {
    v.x = __fmul_rn(v.x, alpha); v.y = __fmul_rn(v.y, alpha); v.z = __fmul_rn(v.z, alpha); v.w = __fmul_rn(v.w, alpha);
  }
  return v;
}

// one group of gw lanes per sorted slot; only piece starts work.  Single-piece segments go straight to grad;
// pieces of longer segments are parked in pieces[(slot / SEG_CHUNK) * 2 + is_first_piece] (provably collision-free:
// two parked pieces inside one SEG_CHUNK window belong to different segments, one a non-first, one a first piece).
__global__ void __launch_bounds__(256) segment_pieces_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                             const float* __restrict__ src, const float* __restrict__ coef, float alpha,
                                                             float* __restrict__ grad, float* __restrict__ pieces, int64_t n, int d4,
                                                             int64_t padding_idx, int gw) {
  int lane = threadIdx.x % gw;
  int64_t s = (int64_t)blockIdx.x * (blockDim.x / gw) + threadIdx.x / gw;
  if (s >= n) return;
  uint32_t key = keys[s];
  if ((int64_t)key == padding_idx) return;
  bool head = s == 0 || keys[s - 1] != key;
  int64_t off = 0;
  if (!head) {
    if (s >= SEG_CHUNK && keys[s - SEG_CHUNK] == key) {  // possibly a later piece start: need the exact offset
      off = s - lower_bound_u32(keys, n, key);
      if (off % SEG_CHUNK != 0) return;
    } else {
      return;  // within the first SEG_CHUNK slots of its segment and not the head
    }
  }
  int64_t e = s + SEG_CHUNK < n ? s + SEG_CHUNK : n;
  bool more = e < n && keys[e] == key;        // the segment continues past this piece
  bool single = off == 0 && !more;
  // contributions of this piece: slots [s, s + cnt); the segment end inside the window by bisection (keys are sorted), so
  // that the ordered sum below can issue its loads in independent batches
  int64_t cnt = e - s;
  if (!more && keys[e - 1] != key) {
    int64_t lo = 0, hi = e - s - 1;  // keys[s + lo] == key, keys[s + hi] != key
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (keys[s + mid] == key) lo = mid; else hi = mid;
    }
    cnt = hi;
  }
  for (int c4 = lane; c4 < d4; c4 += gw) {
    float4 acc = load_contrib(src, coef, alpha, vals[s], d4, c4);
    for (int64_t k0 = 1; k0 < cnt; k0 += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        v[u] = k0 + u < cnt ? load_contrib(src, coef, alpha, vals[s + k0 + u], d4, c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < cnt) {  // index order, one rounding per contribution (the CPU reference's order)
          acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y); acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
        }
    }
    if (single) {
      float* g = grad + ((int64_t)key * d4 + c4) * 4;
      float4 o = ld4(g);
      st4(g, make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w));
    } else {
      st4(pieces + (((s / SEG_CHUNK) * 2 + (off == 0 ? 1 : 0)) * (int64_t)d4 + c4) * 4, acc);
    }
  }
}

// heads of multi-piece segments add their pieces in order
__global__ void __launch_bounds__(256) segment_combine_kernel(const uint32_t* __restrict__ keys, float* __restrict__ grad,
                                                              const float* __restrict__ pieces, int64_t n, int d4, int64_t padding_idx, int gw) {
  int lane = threadIdx.x % gw;
  int64_t s = (int64_t)blockIdx.x * (blockDim.x / gw) + threadIdx.x / gw;
  if (s >= n) return;
  uint32_t key = keys[s];
  if ((int64_t)key == padding_idx) return;
  if (s > 0 && keys[s - 1] == key) return;                 // not a head
  if (!(s + SEG_CHUNK < n && keys[s + SEG_CHUNK] == key)) return;  // single piece: already in grad
  // Number of follow-up pieces (t = s + k*SEG_CHUNK still inside the segment): gallop, then bisect -- a popular row (the
  // BERT mask token holds ~15 % of a batch) has hundreds of pieces, and the ordered sum below wants its loads issued in
  // independent batches instead of one dependent load per iteration.
  int64_t lo = 1, hi = 2;  // invariant: piece lo exists
  while (s + hi * SEG_CHUNK < n && keys[s + hi * SEG_CHUNK] == key) { lo = hi; hi *= 2; }
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (s + mid * SEG_CHUNK < n && keys[s + mid * SEG_CHUNK] == key) lo = mid; else hi = mid;
  }
  const int64_t np = lo;  // pieces 1..np follow the head piece
  const int64_t p0 = s / SEG_CHUNK;
  for (int c4 = lane; c4 < d4; c4 += gw) {
    float4 acc = ld4(pieces + ((p0 * 2 + 1) * (int64_t)d4 + c4) * 4);
    for (int64_t k0 = 1; k0 <= np; k0 += 16) {
      float4 v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u)
        v[u] = k0 + u <= np ? ld4(pieces + (((p0 + k0 + u) * 2 + 0) * (int64_t)d4 + c4) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (k0 + u <= np) {  // same order as the piece-tree contract: ascending, one rounding per piece
          acc.x = __fadd_rn(acc.x, v[u].x); acc.y = __fadd_rn(acc.y, v[u].y); acc.z = __fadd_rn(acc.z, v[u].z); acc.w = __fadd_rn(acc.w, v[u].w);
        }
    }
    float* g = grad + ((int64_t)key * d4 + c4) * 4;
    float4 o = ld4(g);
    st4(g, make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w));
  }
}

int sort_blocks(int64_t n) { return (int)rbm_cdiv(n, TILE); }

}  // namespace

static size_t scatter_sort_bytes(int64_t n) {
  size_t nblk = (size_t)sort_blocks(n < 1 ? 1 : n);
  size_t b = (size_t)4 * (size_t)(n < 1 ? 1 : n) * sizeof(uint32_t) + nblk * RADIX * sizeof(uint32_t);
  return (b + 255) & ~(size_t)255;
}
// d is not part of the query: the piece buffer is sized for d <= 1024
extern "C" size_t rbm_scatter_ws_bytes(int64_t n, int64_t vocab) {
  (void)vocab;
  size_t pieces = ((size_t)(n < 1 ? 1 : n) / SEG_CHUNK + 1) * 2 * 1024 * sizeof(float);
  return scatter_sort_bytes(n) + pieces + 256;
}

extern "C" int rbm_scatter_add_sorted(const int64_t* idx, const float* src, const float* coef, float alpha, float* grad,
                                      int64_t n, int d, int64_t vocab, int64_t padding_idx, void* ws, size_t ws_bytes,
                                      rbm_stream_t stream) {
  RBM_REQUIRE(idx && src && grad && ws, "rbm_scatter_add_sorted: null pointer");
  RBM_REQUIRE(n >= 0 && n < ((int64_t)1 << 32) && vocab > 0 && vocab < ((int64_t)1 << 31), "rbm_scatter_add_sorted: n/vocab out of range");
  RBM_REQUIRE(d >= 4 && d % 4 == 0 && d <= 1024, "rbm_scatter_add_sorted: d=%d must be a multiple of 4, <= 1024", d);
  RBM_REQUIRE(ws_bytes >= rbm_scatter_ws_bytes(n, vocab), "rbm_scatter_add_sorted: workspace too small");
  RBM_REQUIRE(rbm_aligned16(src) && rbm_aligned16(grad) && rbm_aligned16(ws), "rbm_scatter_add_sorted: pointers must be 16B aligned");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int nblk = sort_blocks(n);
  uint32_t* kA = (uint32_t*)ws;
  uint32_t* kB = kA + n;
  uint32_t* vA = kB + n;
  uint32_t* vB = vA + n;
  uint32_t* hist = vB + n;
  int bits = 1;
  while (((int64_t)1 << bits) < vocab) ++bits;
  int passes = (bits + 7) / 8;
  const uint32_t *kin = nullptr, *vin = nullptr;
  uint32_t *kout = kA, *vout = vA;
  for (int p = 0; p < passes; ++p) {
    const int64_t* idx_in = p == 0 ? idx : nullptr;
    radix_hist_kernel<<<nblk, 32 * SORT_WARPS, 0, st>>>(idx_in, kin, n, p * 8, hist, nblk);
    radix_scan_kernel<<<1, 1024, 0, st>>>(hist, (int64_t)RADIX * nblk);
    radix_scatter_kernel<<<nblk, 32 * SORT_WARPS, 0, st>>>(idx_in, kin, vin, kout, vout, n, p * 8, hist, nblk);
    RBM_LAUNCH_CHECK("rbm_scatter_add_sorted(sort)");
    kin = kout;
    vin = vout;
    kout = (kout == kA) ? kB : kA;
    vout = (vout == vA) ? vB : vA;
  }
  int d4 = d / 4;
  int gw = d4 <= 4 ? 4 : d4 <= 8 ? 8 : d4 <= 16 ? 16 : 32;
  float* pieces = (float*)((char*)ws + scatter_sort_bytes(n));
  segment_pieces_kernel<<<(unsigned)rbm_cdiv(n, 256 / gw), 256, 0, st>>>(kin, vin, src, coef, alpha, grad, pieces, n, d4, padding_idx, gw);
  segment_combine_kernel<<<(unsigned)rbm_cdiv(n, 256 / gw), 256, 0, st>>>(kin, grad, pieces, n, d4, padding_idx, gw);
  RBM_LAUNCH_CHECK("rbm_scatter_add_sorted(reduce)");
  return 0;
}
