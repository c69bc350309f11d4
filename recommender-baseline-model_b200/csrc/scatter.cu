// scatter.cu -- embedding-gradient scatter-add as stable sort + sequential segment reduce (bit-deterministic).
// Replaces embedding_dense_backward (autograd of nn.Embedding: NN/models/sas_model/sas.py:30,
// NN/models/bert_modules/embedding/token.py:6; triggered by loss.backward() NN/trainers/base.py:121), which on
// stock CUDA uses float atomics (run-to-run different) and on CPU a serial index-ordered loop.  Here:
//   1. stable LSD radix sort (8-bit digits) of (row id, position) pairs -- positions stay ascending per row id;
//   2. one lane group per segment head walks its segment in order, fp32 adds in ascending position
//      == the CPU reference's summation order, so the result is bit-identical to it and run-to-run stable.
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

// one group of gw lanes per sorted slot; only segment heads work
__global__ void __launch_bounds__(256) segment_reduce_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                             const float* __restrict__ src, const float* __restrict__ coef, float alpha,
                                                             float* __restrict__ grad, int64_t n, int d4, int64_t padding_idx, int gw) {
  int lane = threadIdx.x % gw;
  int64_t j = (int64_t)blockIdx.x * (blockDim.x / gw) + threadIdx.x / gw;
  if (j >= n) return;
  uint32_t key = keys[j];
  if (j > 0 && keys[j - 1] == key) return;
  if ((int64_t)key == padding_idx) return;
  for (int c4 = lane; c4 < d4; c4 += gw) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool first = true;
    for (int64_t t = j; t < n && keys[t] == key; ++t) {
      uint32_t row = vals[t];
      float4 v = ld4(src + ((int64_t)row * d4 + c4) * 4);
      if (coef) {
        float cf = coef[row];
        v.x = __fmul_rn(cf, v.x); v.y = __fmul_rn(cf, v.y); v.z = __fmul_rn(cf, v.z); v.w = __fmul_rn(cf, v.w);
      }
      if (alpha != 1.f) {
        v.x = __fmul_rn(v.x, alpha); v.y = __fmul_rn(v.y, alpha); v.z = __fmul_rn(v.z, alpha); v.w = __fmul_rn(v.w, alpha);
      }
      if (first) {
        acc = v;
        first = false;
      } else {
        acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
      }
    }
    float* g = grad + ((int64_t)key * d4 + c4) * 4;
    float4 o = ld4(g);
    st4(g, make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w));
  }
}

int sort_blocks(int64_t n) { return (int)rbm_cdiv(n, TILE); }

}  // namespace

extern "C" size_t rbm_scatter_ws_bytes(int64_t n, int64_t vocab) {
  (void)vocab;
  size_t nblk = (size_t)sort_blocks(n < 1 ? 1 : n);
  return (size_t)4 * (size_t)(n < 1 ? 1 : n) * sizeof(uint32_t) + nblk * RADIX * sizeof(uint32_t) + 256;
}

extern "C" int rbm_scatter_add_sorted(const int64_t* idx, const float* src, const float* coef, float alpha, float* grad,
                                      int64_t n, int d, int64_t vocab, int64_t padding_idx, void* ws, size_t ws_bytes,
                                      rbm_stream_t stream) {
  RBM_REQUIRE(idx && src && grad && ws, "rbm_scatter_add_sorted: null pointer");
  RBM_REQUIRE(n >= 0 && n < ((int64_t)1 << 32) && vocab > 0 && vocab < ((int64_t)1 << 31), "rbm_scatter_add_sorted: n/vocab out of range");
  RBM_REQUIRE(d >= 4 && d % 4 == 0, "rbm_scatter_add_sorted: d=%d must be a multiple of 4", d);
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
  segment_reduce_kernel<<<(unsigned)rbm_cdiv(n, 256 / gw), 256, 0, st>>>(kin, vin, src, coef, alpha, grad, n, d4, padding_idx, gw);
  RBM_LAUNCH_CHECK("rbm_scatter_add_sorted(reduce)");
  return 0;
}
