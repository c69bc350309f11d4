// attention_live.cu -- SASRec attention on the live-row (compact) layout (csrc/rows.cu, models/sas.py).  The rows of padding
// positions are exactly zero at every block input, so their keys / values equal the projection bias (W.0 + b): ALL padding keys of a
// sequence are the same vector.  For a live query at position i of sequence b the causal softmax of nn.MultiheadAttention
// (NN/models/sas_model/sas.py:69-76) therefore runs over
//     the live keys at positions j <= i   (rows of the compact q / kv tensors), and
//     n_dead(i) copies of ONE key / value pair (b_k, b_v): the padding positions j <= i,
// which is exact, not an approximation: softmax denominators take n_dead(i) * exp(s_dead), the output takes
// kept(i) * p_dead * b_v, kept(i) = the padding keys whose dropout field keeps them -- the Philox fields are still indexed by
// (sequence-head, i, j) over the [L x L] positions, exactly like the dense kernels (common.cuh, rbm_attn_keep).  With Amazon-Beauty-like
// histories (9 of 50 positions live) a (sequence, head) item is ~ 9 x 10 scores instead of 50 x 50, and the [B*L, d] q / k / v / dO
// tensors of the dense layout are never built.
// One warp per (sequence, head, chunk of QC queries or keys) -- a batch holds a few full-length histories whose L^2 / 2 pairs would
// otherwise be one warp's serial tail; lanes own the d_k columns, the other side is walked in position order (deterministic).
// Backward: kernel A (query-major) delta, dq and the padding-key terms (per-query rows of d b_k | d b_v, column-summed afterwards
// in a fixed order); kernel B (key-major) dk, dv, with delta from kernel A and the keep words recomputed per query.
#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int QC = 4;   // queries (keys in the key-major kernel) per warp
constexpr int U = 4;    // independent (query, key) pairs in flight per warp: their loads and warp reductions overlap

struct LiveArgs {
  const float* q; int64_t ldq;
  const float* kv; int64_t ldkv;   // k at column 0, v at column d
  const float* bkv;                // [2 d]: key / value of every padding position
  const int32_t* rows; const int32_t* seq_start; const int64_t* tok;
  float* out;                      // fwd: [cap, d] context; bwd: the saved context (read)
  float* stats;                    // [cap, h, 2] = {row maximum, 1 / row sum}
  const float* dout;               // bwd
  float* dq; float* dkv; float* dead;  // bwd: [cap, d], [cap, 2 d], [cap, 2 d] (per-query d b_k | d b_v)
  float* delta;                    // bwd: [cap, h] = <dO, O> per (query, head)
  int B, L, h, dk, d;
  float scale, inv_keep;
  uint32_t thr16;
  uint64_t seed, site;
};

// keep decisions of query i against the keys 0..63 (bit j), all lanes return the same word
__device__ __forceinline__ unsigned long long keep_mask(const LiveArgs& a, uint64_t site_e, uint64_t bh, int i, int lane) {
  if (a.thr16 == 0) return ~0ull;
  unsigned long long km = 0ull;
  if (lane < 16) {
    const int np = lane >> 2, t = lane & 3;
    const uint4 r = rbm_philox(a.seed, site_e, rbm_attn_call(bh, i >> 4, i & 7, t, np));
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi)
        if (rbm_attn_field(r, ((i >> 3) & 1) * 4 + e * 2 + hi) >= a.thr16) km |= 1ull << (16 * np + 8 * hi + 2 * t + e);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) km |= __shfl_xor_sync(0xffffffffu, km, off);
  return km;
}

struct SeqInfo {
  int b, hh, start, end, c0, c1;  // live rows [start, end) of the sequence; this warp's chunk [c0, c1)
  unsigned long long dead;        // bit j: position j of the sequence is padding
};
__device__ __forceinline__ bool seq_info(const LiveArgs& a, int lane, SeqInfo& s) {
  const int nch = (a.L + QC - 1) / QC;
  const int wid = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int w = wid / nch, ch = wid - w * nch;
  if (w >= a.B * a.h) return false;
  s.b = w / a.h;
  s.hh = w - s.b * a.h;
  s.start = a.seq_start[s.b];
  s.end = a.seq_start[s.b + 1];
  s.c0 = s.start + ch * QC;
  s.c1 = s.c0 + QC < s.end ? s.c0 + QC : s.end;
  if (s.c0 >= s.end) return false;
  const int64_t* tk = a.tok + (int64_t)s.b * a.L;
  const unsigned lo = __ballot_sync(0xffffffffu, lane < a.L && tk[lane] == 0);
  const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < a.L && tk[lane + 32] == 0);
  s.dead = (unsigned long long)lo | ((unsigned long long)hi << 32);
  return true;
}

template <int NT>
__device__ __forceinline__ void load_cols(float (&v)[NT], const float* p, int dk, int lane) {
#pragma unroll
  for (int t = 0; t < NT; ++t) v[t] = lane + 32 * t < dk ? p[lane + 32 * t] : 0.f;
}
template <int NT>
__device__ __forceinline__ float dot_part(const float (&x)[NT], const float (&y)[NT]) {
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) s = fmaf(x[t], y[t], s);
  return s;
}
// N independent warp sums, butterflies interleaved
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] += __shfl_xor_sync(0xffffffffu, v[n], off);
}

template <int NT>
__global__ void __launch_bounds__(32 * WARPS) attn_live_fwd_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31;
  SeqInfo s;
  if (!seq_info(a, lane, s)) return;
  const uint64_t site_e = rbm_site(a.site), bh = (uint64_t)s.b * a.h + s.hh;
  const int co = s.hh * a.dk;
  float bk[NT], bv[NT];
  load_cols(bk, a.bkv + co, a.dk, lane);
  load_cols(bv, a.bkv + a.d + co, a.dk, lane);
  for (int r = s.c0; r < s.c1; ++r) {
    const int i = a.rows[r] - s.b * a.L;
    float qv[NT];
    load_cols(qv, a.q + (int64_t)r * a.ldq + co, a.dk, lane);
    const unsigned long long km = keep_mask(a, site_e, bh, i, lane);
    const unsigned long long causal = i >= 63 ? ~0ull : ((1ull << (i + 1)) - 1ull);
    const int nd = __popcll(s.dead & causal), kd = __popcll(s.dead & causal & km);
    // the padding keys: n_dead copies of one score
    float m = -INFINITY, l = 0.f, o[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) o[t] = 0.f;
    if (nd > 0) {
      float sd[1] = {dot_part(qv, bk)};
      warp_sum_n(sd);
      m = a.scale * sd[0];
      l = (float)nd;
      const float w = (float)kd * a.inv_keep;
#pragma unroll
      for (int t = 0; t < NT; ++t) o[t] = w * bv[t];
    }
    for (int r0 = s.start; r0 <= r; r0 += U) {  // live keys at positions j <= i, ascending, U at a time
      float kk[U][NT], vv[U][NT], sc[U];
      int jj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r2 = r0 + u <= r ? r0 + u : r;
        jj[u] = a.rows[r2] - s.b * a.L;
        const float* kr = a.kv + (int64_t)r2 * a.ldkv + co;
        load_cols(kk[u], kr, a.dk, lane);
        load_cols(vv[u], kr + a.d, a.dk, lane);
        sc[u] = dot_part(qv, kk[u]);
      }
      warp_sum_n(sc);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r0 + u > r) break;
        const float x = a.scale * sc[u];
        const float mn = fmaxf(m, x);
        const float al = m == -INFINITY ? 0.f : __expf(m - mn), e = __expf(x - mn);
        l = fmaf(l, al, e);
        const float w = ((km >> jj[u]) & 1ull) ? e * a.inv_keep : 0.f;
#pragma unroll
        for (int t = 0; t < NT; ++t) o[t] = fmaf(o[t], al, w * vv[u][t]);
        m = mn;
      }
    }
    const float inv = 1.f / l;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane + 32 * t < a.dk) a.out[(int64_t)r * a.d + co + lane + 32 * t] = o[t] * inv;
    if (lane == 0) {
      a.stats[((int64_t)r * a.h + s.hh) * 2] = m;
      a.stats[((int64_t)r * a.h + s.hh) * 2 + 1] = inv;
    }
  }
}

template <int NT>
__global__ void __launch_bounds__(32 * WARPS) attn_live_bwd_q_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31;
  SeqInfo s;
  if (!seq_info(a, lane, s)) return;
  const uint64_t site_e = rbm_site(a.site), bh = (uint64_t)s.b * a.h + s.hh;
  const int co = s.hh * a.dk;
  float bk[NT], bv[NT];
  load_cols(bk, a.bkv + co, a.dk, lane);
  load_cols(bv, a.bkv + a.d + co, a.dk, lane);
  for (int r = s.c0; r < s.c1; ++r) {
    const int i = a.rows[r] - s.b * a.L;
    float qv[NT], dov[NT], ov[NT];
    load_cols(qv, a.q + (int64_t)r * a.ldq + co, a.dk, lane);
    load_cols(dov, a.dout + (int64_t)r * a.d + co, a.dk, lane);
    load_cols(ov, a.out + (int64_t)r * a.d + co, a.dk, lane);
    float red[3] = {dot_part(dov, ov), dot_part(qv, bk), dot_part(dov, bv)};
    warp_sum_n(red);
    const float delta = red[0];
    if (lane == 0) a.delta[(int64_t)r * a.h + s.hh] = delta;
    const float m = a.stats[((int64_t)r * a.h + s.hh) * 2], inv = a.stats[((int64_t)r * a.h + s.hh) * 2 + 1];
    const unsigned long long km = keep_mask(a, site_e, bh, i, lane);
    const unsigned long long causal = i >= 63 ? ~0ull : ((1ull << (i + 1)) - 1ull);
    const int nd = __popcll(s.dead & causal), kd = __popcll(s.dead & causal & km);
    float dq[NT], dbk[NT], dbv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) dq[t] = dbk[t] = dbv[t] = 0.f;
    if (nd > 0) {
      const float pd = __expf(a.scale * red[1] - m) * inv;
      const float wv = (float)kd * a.inv_keep * pd;
      const float dsd = a.scale * (wv * red[2] - (float)nd * pd * delta);  // sum over the padding keys of scale * dS
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        dq[t] = dsd * bk[t];
        dbk[t] = dsd * qv[t];
        dbv[t] = wv * dov[t];
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane + 32 * t < a.dk) {
        a.dead[(int64_t)r * 2 * a.d + co + lane + 32 * t] = dbk[t];
        a.dead[(int64_t)r * 2 * a.d + a.d + co + lane + 32 * t] = dbv[t];
      }
    constexpr int UB = U / 2;
    for (int r0 = s.start; r0 <= r; r0 += UB) {
      float kk[UB][NT], sc[2 * UB];
      int jj[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int r2 = r0 + u <= r ? r0 + u : r;
        jj[u] = a.rows[r2] - s.b * a.L;
        const float* kr = a.kv + (int64_t)r2 * a.ldkv + co;
        float vv[NT];
        load_cols(kk[u], kr, a.dk, lane);
        load_cols(vv, kr + a.d, a.dk, lane);
        sc[2 * u] = dot_part(qv, kk[u]);
        sc[2 * u + 1] = dot_part(dov, vv);
      }
      warp_sum_n(sc);
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (r0 + u > r) break;
        const float p = __expf(a.scale * sc[2 * u] - m) * inv;
        const float mk = ((km >> jj[u]) & 1ull) ? a.inv_keep : 0.f;
        const float ds = a.scale * p * (mk * sc[2 * u + 1] - delta);
#pragma unroll
        for (int t = 0; t < NT; ++t) dq[t] = fmaf(ds, kk[u][t], dq[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane + 32 * t < a.dk) a.dq[(int64_t)r * a.d + co + lane + 32 * t] = dq[t];
  }
}

template <int NT>
__global__ void __launch_bounds__(32 * WARPS) attn_live_bwd_kv_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31;
  SeqInfo s;
  if (!seq_info(a, lane, s)) return;
  const uint64_t site_e = rbm_site(a.site), bh = (uint64_t)s.b * a.h + s.hh;
  const int co = s.hh * a.dk;
  for (int r2 = s.c0; r2 < s.c1; ++r2) {
    const int j = a.rows[r2] - s.b * a.L;
    const float* kr = a.kv + (int64_t)r2 * a.ldkv + co;
    float kk[NT], vv[NT], dk_[NT], dv_[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) dk_[t] = dv_[t] = 0.f;
    load_cols(kk, kr, a.dk, lane);
    load_cols(vv, kr + a.d, a.dk, lane);
    for (int rb = r2; rb < s.end; rb += 32) {  // queries at positions i >= j, ascending, 32 at a time:
      // lane x fetches the scalars of query rb + x and takes its keep decision against key j (one Philox call per lane)
      const int rx = rb + lane;
      float dl = 0.f, ml = 0.f, il = 0.f;
      bool kp = true;
      if (rx < s.end) {
        dl = a.delta[(int64_t)rx * a.h + s.hh];
        ml = a.stats[((int64_t)rx * a.h + s.hh) * 2];
        il = a.stats[((int64_t)rx * a.h + s.hh) * 2 + 1];
        if (a.thr16) kp = rbm_attn_keep(a.seed, site_e, bh, a.rows[rx] - s.b * a.L, j, a.thr16);
      }
      const unsigned kbits = __ballot_sync(0xffffffffu, kp);
      const int nq = s.end - rb < 32 ? s.end - rb : 32;
      constexpr int UB = U / 2;
      for (int x0 = 0; x0 < nq; x0 += UB) {
        float qv[UB][NT], dov[UB][NT], sc[2 * UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int r = rb + (x0 + u < nq ? x0 + u : nq - 1);
          load_cols(qv[u], a.q + (int64_t)r * a.ldq + co, a.dk, lane);
          load_cols(dov[u], a.dout + (int64_t)r * a.d + co, a.dk, lane);
          sc[2 * u] = dot_part(qv[u], kk);
          sc[2 * u + 1] = dot_part(dov[u], vv);
        }
        warp_sum_n(sc);
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int x = x0 + u;
          if (x >= nq) break;
          const float delta = __shfl_sync(0xffffffffu, dl, x), m = __shfl_sync(0xffffffffu, ml, x), inv = __shfl_sync(0xffffffffu, il, x);
          const float p = __expf(a.scale * sc[2 * u] - m) * inv;
          const float mk = ((kbits >> x) & 1u) ? a.inv_keep : 0.f;
          const float ds = a.scale * p * (mk * sc[2 * u + 1] - delta), pw = mk * p;
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            dk_[t] = fmaf(ds, qv[u][t], dk_[t]);
            dv_[t] = fmaf(pw, dov[u][t], dv_[t]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane + 32 * t < a.dk) {
        a.dkv[(int64_t)r2 * 2 * a.d + co + lane + 32 * t] = dk_[t];
        a.dkv[(int64_t)r2 * 2 * a.d + a.d + co + lane + 32 * t] = dv_[t];
      }
  }
}

// first compact row of every sequence: seq_start[b] = lower_bound(rows[0..count), b * L), b = 0..B
__global__ void __launch_bounds__(256) seq_start_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ count, int B, int L,
                                                        int32_t* __restrict__ seq_start) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  int lo = 0, hi = *count;
  const int key = b * L;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (rows[mid] < key) lo = mid + 1; else hi = mid;
  }
  seq_start[b] = lo;
}

int check_args(const char* who, int B, int L, int h, int dk, float p) {
  if (B < 0 || L < 1 || L > 64 || h < 1 || dk < 1 || dk > 128) {
    rbm_set_error("%s: need 1 <= L <= 64 and d_k <= %d (L=%d d_k=%d)", who, 128, L, dk);
    return -1;
  }
  if (!(p >= 0.f && p < 1.f)) {
    rbm_set_error("%s: dropout p out of [0,1)", who);
    return -1;
  }
  return 0;
}

}  // namespace

extern "C" int rbm_rows_seq_start(const int32_t* rows, const int32_t* count, int B, int L, int32_t* seq_start, rbm_stream_t stream) {
  RBM_REQUIRE(rows && count && seq_start && B >= 0 && L >= 1, "rbm_rows_seq_start: bad arguments");
  seq_start_kernel<<<(unsigned)rbm_cdiv(B + 1, 256), 256, 0, (cudaStream_t)stream>>>(rows, count, B, L, seq_start);
  RBM_LAUNCH_CHECK("rbm_rows_seq_start");
  return 0;
}

extern "C" int rbm_attn_live_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* bkv, const int32_t* rows,
                                 const int32_t* seq_start, const int64_t* tok, float* out, float* stats, int B, int L, int h, int dk,
                                 float scale, float p, uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(q && kv && bkv && rows && seq_start && tok && out && stats, "rbm_attn_live_fwd: null pointer");
  if (check_args("rbm_attn_live_fwd", B, L, h, dk, p)) return -1;
  if (B == 0) return 0;
  LiveArgs a{};
  a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.bkv = bkv; a.rows = rows; a.seq_start = seq_start; a.tok = tok; a.out = out; a.stats = stats;
  a.B = B; a.L = L; a.h = h; a.dk = dk; a.d = h * dk; a.scale = scale; a.inv_keep = 1.f / (1.f - p); a.thr16 = rbm_drop_threshold16(p);
  a.seed = seed; a.site = site;
  const unsigned grid = (unsigned)rbm_cdiv((int64_t)B * h * ((L + QC - 1) / QC), WARPS);
  if (dk <= 32) attn_live_fwd_kernel<1><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk <= 64) attn_live_fwd_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_fwd_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_fwd");
  return 0;
}

extern "C" int rbm_attn_live_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* bkv, const int32_t* rows,
                                 const int32_t* seq_start, const int64_t* tok, const float* out, const float* stats, const float* dout,
                                 float* dq, float* dkv, float* dead, float* delta, int B, int L, int h, int dk, float scale, float p,
                                 uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(q && kv && bkv && rows && seq_start && tok && out && stats && dout && dq && dkv && dead && delta, "rbm_attn_live_bwd: null pointer");
  if (check_args("rbm_attn_live_bwd", B, L, h, dk, p)) return -1;
  if (B == 0) return 0;
  LiveArgs a{};
  a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.bkv = bkv; a.rows = rows; a.seq_start = seq_start; a.tok = tok;
  a.out = const_cast<float*>(out); a.stats = const_cast<float*>(stats); a.dout = dout; a.dq = dq; a.dkv = dkv; a.dead = dead; a.delta = delta;
  a.B = B; a.L = L; a.h = h; a.dk = dk; a.d = h * dk; a.scale = scale; a.inv_keep = 1.f / (1.f - p); a.thr16 = rbm_drop_threshold16(p);
  a.seed = seed; a.site = site;
  const unsigned grid = (unsigned)rbm_cdiv((int64_t)B * h * ((L + QC - 1) / QC), WARPS);
  if (dk <= 32) attn_live_bwd_q_kernel<1><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk <= 64) attn_live_bwd_q_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_bwd_q_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_bwd(dq)");
  if (dk <= 32) attn_live_bwd_kv_kernel<1><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk <= 64) attn_live_bwd_kv_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_bwd_kv_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_bwd(dkv)");
  return 0;
}

RBM_DEFINE_STEP_PTR_SETTER(rbm_step_ptr_set_attention_live)
