// attention.cu -- short-sequence (L <= 256) multi-head attention, forward and backward, fp32.
// One CTA per (sequence, head): the whole K/V (or Q/dO) panel of the head lives in shared memory, every warp owns
// groups of R query (or key) rows, scores stay in registers, softmax via warp shuffles, dropout regenerated from
// Philox -- the [B,h,L,L] probability tensor the reference materialises (NN/models/bert_modules/attention/single.py:14-33,
// torch MHA inside NN/models/sas_model/sas.py:75-76) never touches HBM.  Backward is recompute-based and
// atomics-free (pass A: dQ by query rows; pass B: dK,dV by key rows) so results are bit-deterministic.
#include "common.cuh"

namespace {

constexpr int WARPS = 8;
constexpr int DCMAX = 4;  // dk <= 128

struct AttnArgs {
  const float *q, *k, *v, *o, *dout, *stats_in;
  float *out, *stats, *dq, *dk_, *dv, *delta;
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  const int64_t* tok;
  int L, h, dk, mask_mode;
  float scale;
  uint32_t thr;
  float inv_keep;
  uint64_t seed, site;
};

// dst[LP][KS] <- src rows (b*L + j), columns [hh*dk, hh*dk+dk); rows >= L zero-filled.  All threads of the CTA.
__device__ __forceinline__ void load_panel(float* dst, const float* __restrict__ src, int64_t ld, int64_t row0, int col0,
                                           int L, int LP, int dk, int KS, float mul) {
  int dk4 = dk >> 2;
  for (int idx = threadIdx.x; idx < LP * dk4; idx += blockDim.x) {
    int j = idx / dk4, c4 = idx - j * dk4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < L) v = ld4(src + (row0 + j) * ld + col0 + c4 * 4);
    float* d = dst + j * KS + c4 * 4;
    d[0] = v.x * mul; d[1] = v.y * mul; d[2] = v.z * mul; d[3] = v.w * mul;
  }
}

// per-warp: dst[c*R + r] <- src row (row0 + i0 + r), r < R (zero beyond L)
template <int R>
__device__ __forceinline__ void load_rows_cr(float* dst, const float* __restrict__ src, int64_t ld, int64_t row0, int col0,
                                             int i0, int L, int dk, float mul, int lane) {
  for (int idx = lane; idx < R * dk; idx += 32) {
    int r = idx / dk, c = idx - r * dk;
    int i = i0 + r;
    dst[c * R + r] = i < L ? src[(row0 + i) * ld + col0 + c] * mul : 0.f;
  }
}

// acc[r][jj] += sum_c rows_cr[c][r] * panel[(lane + 32*jj)][c]
template <int NJ, int R>
__device__ __forceinline__ void rows_dot_panel(const float* rows_cr, const float* panel, int KS, int dk, int lane,
                                               float (&acc)[R][NJ]) {
  for (int c = 0; c < dk; ++c) {
    float rv[R];
#pragma unroll
    for (int r4 = 0; r4 < R / 4; ++r4) {
      float4 t = ld4(rows_cr + c * R + r4 * 4);
      rv[r4 * 4] = t.x; rv[r4 * 4 + 1] = t.y; rv[r4 * 4 + 2] = t.z; rv[r4 * 4 + 3] = t.w;
    }
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      float pv = panel[(lane + 32 * jj) * KS + c];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r][jj] = fmaf(rv[r], pv, acc[r][jj]);
    }
  }
}

// o[r][cc] = sum_{j in [jb, je)} w_jr[j*R + r] * panel[j][cc*32 + lane]
template <int R>
__device__ __forceinline__ void weights_times_panel(const float* w_jr, const float* panel, int KS, int dk, int jb, int je,
                                                    int lane, float (&o)[R][DCMAX]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int cc = 0; cc < DCMAX; ++cc) o[r][cc] = 0.f;
  for (int j = jb; j < je; ++j) {
    float wv[R];
#pragma unroll
    for (int r4 = 0; r4 < R / 4; ++r4) {
      float4 t = ld4(w_jr + j * R + r4 * 4);
      wv[r4 * 4] = t.x; wv[r4 * 4 + 1] = t.y; wv[r4 * 4 + 2] = t.z; wv[r4 * 4 + 3] = t.w;
    }
#pragma unroll
    for (int cc = 0; cc < DCMAX; ++cc) {
      if (cc * 32 < dk) {
        int c = cc * 32 + lane;
        float pv = c < dk ? panel[j * KS + c] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) o[r][cc] = fmaf(wv[r], pv, o[r][cc]);
      }
    }
  }
}

__device__ __forceinline__ bool key_masked_inf(int mask_mode, int i, int j, int L) {
  return j >= L || (mask_mode == RBM_MASK_CAUSAL && j > i);
}

// ------------------------------------------------------------------------------------------------ forward
template <int NJ, int R>
__global__ void __launch_bounds__(32 * WARPS, 1) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, dk = a.dk, KS = dk + 1, LP = NJ * 32;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)b * L;
  const int col0 = hh * dk;
  float* Ks = sm;
  float* Vs = Ks + LP * KS;
  float* padk = Vs + LP * KS;  // [LP] 1.0 where the key token is padding
  float* wbase = padk + LP + (size_t)warp * (R * dk + LP * R);
  float* Qs = wbase;           // [dk][R]
  float* Ps = wbase + R * dk;  // [LP][R]

  load_panel(Ks, a.k, a.ldk, row0, col0, L, LP, dk, KS, 1.f);
  load_panel(Vs, a.v, a.ldv, row0, col0, L, LP, dk, KS, 1.f);
  for (int j = threadIdx.x; j < LP; j += blockDim.x)
    padk[j] = (a.mask_mode == RBM_MASK_KEYPAD && j < L && a.tok[row0 + j] == 0) ? 1.f : 0.f;
  __syncthreads();

  for (int i0 = warp * R; i0 < L; i0 += WARPS * R) {
    load_rows_cr<R>(Qs, a.q, a.ldq, row0, col0, i0, L, dk, a.scale, lane);
    __syncwarp();
    float s[R][NJ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) s[r][jj] = 0.f;
    rows_dot_panel<NJ, R>(Qs, Ks, KS, dk, lane, s);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r;
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        int j = lane + 32 * jj;
        float v = s[r][jj];
        if (padk[j] != 0.f) v = -1e9f;
        if (key_masked_inf(a.mask_mode, i, j, L)) v = -INFINITY;
        s[r][jj] = v;
        mx = fmaxf(mx, v);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        float e = expf(s[r][jj] - mx);
        s[r][jj] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      float inv = 1.f / sum;
      if (lane == 0 && i < L && a.stats) {
        int64_t sr = ((int64_t)blockIdx.x * L + i) * 2;
        a.stats[sr] = mx;
        a.stats[sr + 1] = inv;
      }
      const uint64_t Rrow = (uint64_t)blockIdx.x * L + i;
#pragma unroll
      for (int g = 0; g < (NJ + 3) / 4; ++g) {
        uint4 rnd = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (a.thr) rnd = rbm_philox(a.seed, a.site, rbm_attn_call(Rrow, lane + 128 * g));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          int jj = g * 4 + t;
          if (jj < NJ) {
            float p = s[r][jj] * inv;
            if (a.thr) p = rbm_u4_get(rnd, t) >= a.thr ? p * a.inv_keep : 0.f;
            Ps[(lane + 32 * jj) * R + r] = p;
          }
        }
      }
    }
    __syncwarp();
    float o[R][DCMAX];
    int je = a.mask_mode == RBM_MASK_CAUSAL ? (i0 + R < L ? i0 + R : L) : L;
    weights_times_panel<R>(Ps, Vs, KS, dk, 0, je, lane, o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int i = i0 + r;
      if (i < L) {
#pragma unroll
        for (int cc = 0; cc < DCMAX; ++cc) {
          int c = cc * 32 + lane;
          if (c < dk) a.out[(row0 + i) * a.ldo + col0 + c] = o[r][cc];
        }
      }
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------- backward pass A: dQ and delta
template <int NJ, int R>
__global__ void __launch_bounds__(32 * WARPS, 1) attn_bwd_dq_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, dk = a.dk, KS = dk + 1, LP = NJ * 32;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)b * L;
  const int col0 = hh * dk;
  float* Ks = sm;
  float* Vs = Ks + LP * KS;
  float* padk = Vs + LP * KS;
  float* wbase = padk + LP + (size_t)warp * (2 * R * dk + LP * R);
  float* Qs = wbase;            // [dk][R]  (scaled q)
  float* dOs = wbase + R * dk;  // [dk][R]
  float* Ps = dOs + R * dk;     // [LP][R]  dS

  load_panel(Ks, a.k, a.ldk, row0, col0, L, LP, dk, KS, 1.f);
  load_panel(Vs, a.v, a.ldv, row0, col0, L, LP, dk, KS, 1.f);
  for (int j = threadIdx.x; j < LP; j += blockDim.x)
    padk[j] = (a.mask_mode == RBM_MASK_KEYPAD && j < L && a.tok[row0 + j] == 0) ? 1.f : 0.f;
  __syncthreads();

  for (int i0 = warp * R; i0 < L; i0 += WARPS * R) {
    load_rows_cr<R>(Qs, a.q, a.ldq, row0, col0, i0, L, dk, a.scale, lane);
    load_rows_cr<R>(dOs, a.dout, a.lddo, row0, col0, i0, L, dk, 1.f, lane);
    float delta[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int i = i0 + r;
      float part = 0.f;
      if (i < L)
        for (int c = lane; c < dk; c += 32)
          part = fmaf(a.dout[(row0 + i) * a.lddo + col0 + c], a.o[(row0 + i) * a.ldo + col0 + c], part);
      delta[r] = warp_sum(part);
      if (lane == 0 && i < L) a.delta[(int64_t)blockIdx.x * L + i] = delta[r];
    }
    __syncwarp();
    float s[R][NJ], dp[R][NJ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) s[r][jj] = dp[r][jj] = 0.f;
    rows_dot_panel<NJ, R>(Qs, Ks, KS, dk, lane, s);
    rows_dot_panel<NJ, R>(dOs, Vs, KS, dk, lane, dp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r;
      float mx = 0.f, inv = 0.f;
      if (i < L) {
        int64_t sr = ((int64_t)blockIdx.x * L + i) * 2;
        mx = a.stats_in[sr];
        inv = a.stats_in[sr + 1];
      }
      const uint64_t Rrow = (uint64_t)blockIdx.x * L + i;
#pragma unroll
      for (int g = 0; g < (NJ + 3) / 4; ++g) {
        uint4 rnd = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (a.thr) rnd = rbm_philox(a.seed, a.site, rbm_attn_call(Rrow, lane + 128 * g));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          int jj = g * 4 + t;
          if (jj < NJ) {
            int j = lane + 32 * jj;
            float v = s[r][jj];
            if (padk[j] != 0.f) v = -1e9f;
            float p = (i >= L || key_masked_inf(a.mask_mode, i, j, L)) ? 0.f : expf(v - mx) * inv;
            float mk = 1.f;
            if (a.thr) mk = rbm_u4_get(rnd, t) >= a.thr ? a.inv_keep : 0.f;
            // masked_fill(-1e9) replaces the score by a constant: no gradient reaches q.k through a padded key
            Ps[j * R + r] = padk[j] != 0.f ? 0.f : p * (mk * dp[r][jj] - delta[r]);
          }
        }
      }
    }
    __syncwarp();
    float o[R][DCMAX];
    int je = a.mask_mode == RBM_MASK_CAUSAL ? (i0 + R < L ? i0 + R : L) : L;
    weights_times_panel<R>(Ps, Ks, KS, dk, 0, je, lane, o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int i = i0 + r;
      if (i < L) {
#pragma unroll
        for (int cc = 0; cc < DCMAX; ++cc) {
          int c = cc * 32 + lane;
          if (c < dk) a.dq[(row0 + i) * a.lddq + col0 + c] = o[r][cc] * a.scale;
        }
      }
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------- backward pass B: dK and dV
template <int NJ, int R>
__global__ void __launch_bounds__(32 * WARPS, 1) attn_bwd_dkv_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, dk = a.dk, KS = dk + 1, LP = NJ * 32;
  const int b = blockIdx.x / a.h, hh = blockIdx.x % a.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row0 = (int64_t)b * L;
  const int col0 = hh * dk;
  float* Qa = sm;               // [LP][KS] scaled q
  float* dOa = Qa + LP * KS;    // [LP][KS]
  float* st_m = dOa + LP * KS;  // [LP]
  float* st_i = st_m + LP;      // [LP]
  float* st_d = st_i + LP;      // [LP]
  float* wbase = st_d + LP + (size_t)warp * (2 * R * dk + 2 * LP * R);
  float* Kg = wbase;            // [dk][R]
  float* Vg = Kg + R * dk;      // [dk][R]
  float* P1 = Vg + R * dk;      // [LP][R] dS^T
  float* P2 = P1 + LP * R;      // [LP][R] P~^T

  load_panel(Qa, a.q, a.ldq, row0, col0, L, LP, dk, KS, a.scale);
  load_panel(dOa, a.dout, a.lddo, row0, col0, L, LP, dk, KS, 1.f);
  for (int i = threadIdx.x; i < LP; i += blockDim.x) {
    bool in = i < L;
    int64_t sr = ((int64_t)blockIdx.x * L + i) * 2;
    st_m[i] = in ? a.stats_in[sr] : 0.f;
    st_i[i] = in ? a.stats_in[sr + 1] : 0.f;
    st_d[i] = in ? a.delta[(int64_t)blockIdx.x * L + i] : 0.f;
  }
  __syncthreads();

  for (int j0 = warp * R; j0 < L; j0 += WARPS * R) {
    load_rows_cr<R>(Kg, a.k, a.ldk, row0, col0, j0, L, dk, 1.f, lane);
    load_rows_cr<R>(Vg, a.v, a.ldv, row0, col0, j0, L, dk, 1.f, lane);
    __syncwarp();
    float s[R][NJ], dp[R][NJ];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int ii = 0; ii < NJ; ++ii) s[r][ii] = dp[r][ii] = 0.f;
    rows_dot_panel<NJ, R>(Kg, Qa, KS, dk, lane, s);
    rows_dot_panel<NJ, R>(Vg, dOa, KS, dk, lane, dp);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = j0 + r;
      const bool jpad = a.mask_mode == RBM_MASK_KEYPAD && j < L && a.tok[row0 + j] == 0;
#pragma unroll
      for (int ii = 0; ii < NJ; ++ii) {
        int i = lane + 32 * ii;
        float v = jpad ? -1e9f : s[r][ii];
        float p = (i >= L || key_masked_inf(a.mask_mode, i, j, L)) ? 0.f : expf(v - st_m[i]) * st_i[i];
        float mk = 1.f;
        if (a.thr) {
          uint4 rnd = rbm_philox(a.seed, a.site, rbm_attn_call((uint64_t)blockIdx.x * L + i, j));
          mk = rbm_u4_get(rnd, (j >> 5) & 3) >= a.thr ? a.inv_keep : 0.f;
        }
        P1[i * R + r] = jpad ? 0.f : p * (mk * dp[r][ii] - st_d[i]);  // no score gradient through a padded key
        P2[i * R + r] = p * mk;
      }
    }
    __syncwarp();
    float o[R][DCMAX];
    int ib = a.mask_mode == RBM_MASK_CAUSAL ? j0 : 0;
    weights_times_panel<R>(P1, Qa, KS, dk, ib, L, lane, o);  // Qa already carries `scale`
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int j = j0 + r;
      if (j < L) {
#pragma unroll
        for (int cc = 0; cc < DCMAX; ++cc) {
          int c = cc * 32 + lane;
          if (c < dk) a.dk_[(row0 + j) * a.lddk + col0 + c] = o[r][cc];
        }
      }
    }
    weights_times_panel<R>(P2, dOa, KS, dk, ib, L, lane, o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int j = j0 + r;
      if (j < L) {
#pragma unroll
        for (int cc = 0; cc < DCMAX; ++cc) {
          int c = cc * 32 + lane;
          if (c < dk) a.dv[(row0 + j) * a.lddv + col0 + c] = o[r][cc];
        }
      }
    }
    __syncwarp();
  }
}

int pick_nj(int L) { return L <= 32 ? 1 : L <= 64 ? 2 : L <= 128 ? 4 : L <= 224 ? 7 : 8; }

size_t smem_fwd(int NJ, int R, int dk) { return sizeof(float) * ((size_t)2 * NJ * 32 * (dk + 1) + NJ * 32 + (size_t)WARPS * (R * dk + NJ * 32 * R)); }
size_t smem_dq(int NJ, int R, int dk) { return sizeof(float) * ((size_t)2 * NJ * 32 * (dk + 1) + NJ * 32 + (size_t)WARPS * (2 * R * dk + NJ * 32 * R)); }
size_t smem_dkv(int NJ, int R, int dk) { return sizeof(float) * ((size_t)2 * NJ * 32 * (dk + 1) + 3 * NJ * 32 + (size_t)WARPS * (2 * R * dk + 2 * NJ * 32 * R)); }

template <typename Kern>
int launch(Kern kern, const AttnArgs& a, int B, size_t smem, cudaStream_t st, const char* name) {
  if (smem > 227 * 1024) {
    rbm_set_error("%s: L=%d dk=%d needs %zu B of shared memory (> 227 KB): unsupported shape", name, a.L, a.dk, smem);
    return -1;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    rbm_set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
    return (int)e;
  }
  kern<<<B * a.h, 32 * WARPS, smem, st>>>(a);
  RBM_LAUNCH_CHECK(name);
  return 0;
}

int check_common(const char* name, int B, int L, int h, int dk, int mask_mode, float p, const int64_t* tok) {
  RBM_REQUIRE(B > 0 && L > 0 && L <= 256 && h > 0, "%s: need B>0, 0<L<=256, h>0 (B=%d L=%d h=%d)", name, B, L, h);
  RBM_REQUIRE(dk >= 4 && dk % 4 == 0 && dk <= 32 * DCMAX, "%s: unsupported head dim %d (need dk%%4==0, dk<=128)", name, dk);
  RBM_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "%s: bad mask_mode %d", name, mask_mode);
  RBM_REQUIRE(mask_mode != RBM_MASK_KEYPAD || tok, "%s: key-padding mask needs tok", name);
  RBM_REQUIRE(p >= 0.f && p < 1.f, "%s: dropout p out of [0,1)", name);
  return 0;
}

}  // namespace

#define ATTN_DISPATCH(KERNEL, R, SMEMFN, NAME)                                                       \
  switch (nj) {                                                                                      \
    case 1: rc = launch(KERNEL<1, R>, a, B, SMEMFN(1, R, dk), st, NAME); break;                      \
    case 2: rc = launch(KERNEL<2, R>, a, B, SMEMFN(2, R, dk), st, NAME); break;                      \
    case 4: rc = launch(KERNEL<4, R>, a, B, SMEMFN(4, R, dk), st, NAME); break;                      \
    case 7: rc = launch(KERNEL<7, R>, a, B, SMEMFN(7, R, dk), st, NAME); break;                      \
    default: rc = launch(KERNEL<8, R>, a, B, SMEMFN(8, R, dk), st, NAME); break;                     \
  }

extern "C" int rbm_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                            const int64_t* tok, float* out, int64_t ldo, float* stats, int B, int L, int h, int dk,
                            int mask_mode, float scale, float p, uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(q && k && v && out, "rbm_attn_fwd: null pointer");
  if (check_common("rbm_attn_fwd", B, L, h, dk, mask_mode, p, tok)) return -1;
  RBM_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && rbm_aligned16(q) && rbm_aligned16(k) && rbm_aligned16(v),
              "rbm_attn_fwd: q/k/v must be 16B aligned with strides %% 4 == 0");
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.out = out; a.stats = stats; a.tok = tok;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.L = L; a.h = h; a.dk = dk; a.mask_mode = mask_mode; a.scale = scale;
  a.thr = rbm_drop_threshold(p); a.inv_keep = 1.f / (1.f - p); a.seed = seed; a.site = site;
  cudaStream_t st = (cudaStream_t)stream;
  int nj = pick_nj(L), rc;
  if (L > 64) { ATTN_DISPATCH(attn_fwd_kernel, 8, smem_fwd, "rbm_attn_fwd") }
  else { ATTN_DISPATCH(attn_fwd_kernel, 4, smem_fwd, "rbm_attn_fwd") }
  return rc;
}

extern "C" size_t rbm_attn_bwd_ws_bytes(int B, int L, int h) { return (size_t)B * L * h * sizeof(float); }

extern "C" int rbm_attn_bwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                            const int64_t* tok, const float* out, int64_t ldo, const float* dout, int64_t lddo,
                            const float* stats, float* dq, int64_t lddq, float* dk_, int64_t lddk, float* dv,
                            int64_t lddv, int B, int L, int h, int dk, int mask_mode, float scale, float p, uint64_t seed,
                            uint64_t site, void* ws, size_t ws_bytes, rbm_stream_t stream) {
  RBM_REQUIRE(q && k && v && out && dout && stats && dq && dk_ && dv && ws, "rbm_attn_bwd: null pointer");
  if (check_common("rbm_attn_bwd", B, L, h, dk, mask_mode, p, tok)) return -1;
  RBM_REQUIRE(ws_bytes >= rbm_attn_bwd_ws_bytes(B, L, h), "rbm_attn_bwd: workspace too small");
  RBM_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && lddo % 4 == 0 && rbm_aligned16(q) && rbm_aligned16(k) &&
                  rbm_aligned16(v) && rbm_aligned16(dout),
              "rbm_attn_bwd: q/k/v/dout must be 16B aligned with strides %% 4 == 0");
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.o = out; a.dout = dout; a.stats_in = stats; a.tok = tok;
  a.dq = dq; a.dk_ = dk_; a.dv = dv; a.delta = (float*)ws;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.L = L; a.h = h; a.dk = dk; a.mask_mode = mask_mode; a.scale = scale;
  a.thr = rbm_drop_threshold(p); a.inv_keep = 1.f / (1.f - p); a.seed = seed; a.site = site;
  cudaStream_t st = (cudaStream_t)stream;
  int nj = pick_nj(L), rc;
  ATTN_DISPATCH(attn_bwd_dq_kernel, 4, smem_dq, "rbm_attn_bwd(dq)")
  if (rc) return rc;
  ATTN_DISPATCH(attn_bwd_dkv_kernel, 4, smem_dkv, "rbm_attn_bwd(dkv)")
  return rc;
}
