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
// A warp takes a chunk of four consecutive live rows of one (sequence, head); each 8-lane group owns one of them (a query, or a key
// in the key-major kernel) and d_k / 8 columns per lane, and walks the other side in position order (deterministic).  Measured on
// the first version (lanes over columns, one query at a time): 68 M warp instructions per launch, two thirds of them per-query
// overhead (Philox, five-step reductions) -- hence four rows per warp and three-step group reductions.
// Backward: kernel A (query-major) delta, keep words, dq and the padding-key terms (per-query rows of d b_k | d b_v, column-summed
// afterwards in a fixed order); kernel B (key-major) dk, dv.
#include "common.cuh"

namespace {

constexpr int WARPS = 4;
constexpr int QC = 4;   // rows per warp: one per 8-lane group

struct LiveArgs {
  const float* q; int64_t ldq;
  const float* kv; int64_t ldkv;   // k at column 0, v at column d
  const float* bkv;                // [2 d]: key / value of every padding position
  const int32_t* rows; const int32_t* seq_start; const int64_t* tok;
  float* out;                      // fwd: [cap, d] context; bwd: the saved context (read)
  float* stats;                    // [cap, h, 2] = {row maximum, 1 / row sum}
  const float* dout;               // bwd
  float* dq; float* dkv; float* dead;  // bwd: [cap, d], [cap, 2 d], [cap, 2 d] (per-query d b_k | d b_v)
  float* delta;                    // bwd scratch: [cap, h] = <dO, O> per (query, head)
  unsigned long long* keepw;       // bwd scratch: [cap, h] keep word of the query (bit j = key position j kept)
  int B, L, h, dk, d;
  float scale, inv_keep;
  uint32_t thr16;
  uint64_t seed, site;
};

// keep decisions of query i against the key positions 0..63 (bit j); the 8 lanes of a group share the 16 Philox calls
__device__ __forceinline__ unsigned long long keep_mask8(const LiveArgs& a, uint64_t site_e, uint64_t bh, int i, int sub) {
  if (a.thr16 == 0) return ~0ull;
  unsigned long long km = 0ull;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int call = sub + 8 * c, np = call >> 2, t = call & 3;
    const uint4 r = rbm_philox_drop(a.seed, site_e, rbm_attn_call(bh, i >> 4, i & 7, t, np));
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi)
        if (rbm_attn_field(r, ((i >> 3) & 1) * 4 + e * 2 + hi) >= a.thr16) km |= 1ull << (16 * np + 8 * hi + 2 * t + e);
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) km |= __shfl_xor_sync(0xffffffffu, km, off);
  return km;
}

struct SeqInfo {
  int b, hh, start, end, c0;  // live rows [start, end) of the sequence; this warp's chunk starts at c0
  unsigned long long dead;    // bit j: position j of the sequence is padding
};
__device__ __forceinline__ bool seq_info(const LiveArgs& a, int wid, int lane, SeqInfo& s) {
  const int nch = (a.L + QC - 1) / QC;
  const int w = wid / nch, ch = wid - w * nch;
  s.b = w / a.h;
  s.hh = w - s.b * a.h;
  s.start = a.seq_start[s.b];
  s.end = a.seq_start[s.b + 1];
  s.c0 = s.start + ch * QC;
  if (s.c0 >= s.end) return false;
  const int64_t* tk = a.tok + (int64_t)s.b * a.L;
  const unsigned lo = __ballot_sync(0xffffffffu, lane < a.L && tk[lane] == 0);
  const unsigned hi = __ballot_sync(0xffffffffu, lane + 32 < a.L && tk[lane + 32] == 0);
  s.dead = (unsigned long long)lo | ((unsigned long long)hi << 32);
  return true;
}

template <int CPL>
__device__ __forceinline__ void load_cols(float (&v)[CPL], const float* p) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int c = 0; c < CPL; c += 4) {
      const float4 x = ld4(p + c);
      v[c] = x.x; v[c + 1] = x.y; v[c + 2] = x.z; v[c + 3] = x.w;
    }
  } else {
    const float2 x = *reinterpret_cast<const float2*>(p);
    v[0] = x.x; v[1] = x.y;
  }
}
template <int CPL>
__device__ __forceinline__ void store_cols(float* p, const float (&v)[CPL]) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int c = 0; c < CPL; c += 4) st4(p + c, make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]));
  } else {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  }
}
template <int CPL>
__device__ __forceinline__ float dot_part(const float (&x)[CPL], const float (&y)[CPL]) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) s = fmaf(x[c], y[c], s);
  return s;
}
// sums over the 8 lanes of a group (every lane of the warp takes part)
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ int warp_max_int(int v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

template <int CPL>
__global__ void __launch_bounds__(32 * WARPS) attn_live_fwd_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int nwork = a.B * a.h * ((a.L + QC - 1) / QC), nwarps = gridDim.x * WARPS;
  for (int wid = blockIdx.x * WARPS + (threadIdx.x >> 5); wid < nwork; wid += nwarps) {
    SeqInfo s;
    if (!seq_info(a, wid, lane, s)) continue;
    const uint64_t site_e = rbm_site(a.site), bh = (uint64_t)s.b * a.h + s.hh;
    const int co = s.hh * a.dk + sub * CPL;
    const int r = s.c0 + grp;
    const bool valid = r < s.end;
    const int rq = valid ? r : s.end - 1;
    const int i = a.rows[rq] - s.b * a.L;
    float qv[CPL], bk[CPL], bv[CPL], o[CPL];
    load_cols(qv, a.q + (int64_t)rq * a.ldq + co);
    load_cols(bk, a.bkv + co);
    load_cols(bv, a.bkv + a.d + co);
    const unsigned long long km = keep_mask8(a, site_e, bh, i, sub);
    const unsigned long long causal = i >= 63 ? ~0ull : ((1ull << (i + 1)) - 1ull);
    const int nd = __popcll(s.dead & causal), kd = __popcll(s.dead & causal & km);
    // the padding keys: n_dead copies of one score
    const float sdead = a.scale * group_sum(dot_part(qv, bk));
    float m = nd > 0 ? sdead : -INFINITY, l = (float)nd;
    {
      const float w = (float)kd * a.inv_keep;
#pragma unroll
      for (int c = 0; c < CPL; ++c) o[c] = w * bv[c];
    }
    const int nk = valid ? rq - s.start + 1 : 0, nk_max = warp_max_int(nk);
    for (int x = 0; x < nk_max; ++x) {  // live keys at positions j <= i, ascending
      const bool act = x < nk;
      const int r2 = s.start + (act ? x : 0);
      const int j = a.rows[r2] - s.b * a.L;
      const float* kr = a.kv + (int64_t)r2 * a.ldkv + co;
      float kk[CPL], vv[CPL];
      load_cols(kk, kr);
      load_cols(vv, kr + a.d);
      const float sc = a.scale * group_sum(dot_part(qv, kk));
      if (act) {
        const float mn = fmaxf(m, sc);
        const float al = m == -INFINITY ? 0.f : __expf(m - mn), e = __expf(sc - mn);
        l = fmaf(l, al, e);
        const float w = ((km >> j) & 1ull) ? e * a.inv_keep : 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) o[c] = fmaf(o[c], al, w * vv[c]);
        m = mn;
      }
    }
    if (valid) {
      const float inv = 1.f / l;
#pragma unroll
      for (int c = 0; c < CPL; ++c) o[c] *= inv;
      store_cols(a.out + (int64_t)r * a.d + co, o);
      if (sub == 0) {
        a.stats[((int64_t)r * a.h + s.hh) * 2] = m;
        a.stats[((int64_t)r * a.h + s.hh) * 2 + 1] = inv;
      }
    }
  }
}

template <int CPL>
__global__ void __launch_bounds__(32 * WARPS) attn_live_bwd_q_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int nwork = a.B * a.h * ((a.L + QC - 1) / QC), nwarps = gridDim.x * WARPS;
  for (int wid = blockIdx.x * WARPS + (threadIdx.x >> 5); wid < nwork; wid += nwarps) {
    SeqInfo s;
    if (!seq_info(a, wid, lane, s)) continue;
    const uint64_t site_e = rbm_site(a.site), bh = (uint64_t)s.b * a.h + s.hh;
    const int co = s.hh * a.dk + sub * CPL;
    const int r = s.c0 + grp;
    const bool valid = r < s.end;
    const int rq = valid ? r : s.end - 1;
    const int i = a.rows[rq] - s.b * a.L;
    float qv[CPL], dov[CPL], bk[CPL], bv[CPL], dq[CPL];
    load_cols(qv, a.q + (int64_t)rq * a.ldq + co);
    load_cols(dov, a.dout + (int64_t)rq * a.d + co);
    load_cols(bk, a.bkv + co);
    load_cols(bv, a.bkv + a.d + co);
    float delta;
    {
      float ov[CPL];
      load_cols(ov, a.out + (int64_t)rq * a.d + co);
      delta = group_sum(dot_part(dov, ov));
    }
    const float m = a.stats[((int64_t)rq * a.h + s.hh) * 2], inv = a.stats[((int64_t)rq * a.h + s.hh) * 2 + 1];
    const unsigned long long km = keep_mask8(a, site_e, bh, i, sub);
    if (valid && sub == 0) {
      a.delta[(int64_t)r * a.h + s.hh] = delta;
      a.keepw[(int64_t)r * a.h + s.hh] = km;
    }
    const unsigned long long causal = i >= 63 ? ~0ull : ((1ull << (i + 1)) - 1ull);
    const int nd = __popcll(s.dead & causal), kd = __popcll(s.dead & causal & km);
    {  // the padding keys
      const float sd = a.scale * group_sum(dot_part(qv, bk)), dpd = group_sum(dot_part(dov, bv));
      const float pd = nd > 0 ? __expf(sd - m) * inv : 0.f;
      const float wv = (float)kd * a.inv_keep * pd;
      const float dsd = a.scale * (wv * dpd - (float)nd * pd * delta);  // sum over the padding keys of scale * dS
      float dbk[CPL], dbv[CPL];
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        dq[c] = dsd * bk[c];
        dbk[c] = dsd * qv[c];
        dbv[c] = wv * dov[c];
      }
      if (valid) {
        store_cols(a.dead + (int64_t)r * 2 * a.d + co, dbk);
        store_cols(a.dead + (int64_t)r * 2 * a.d + a.d + co, dbv);
      }
    }
    const int nk = valid ? rq - s.start + 1 : 0, nk_max = warp_max_int(nk);
    for (int x = 0; x < nk_max; ++x) {
      const bool act = x < nk;
      const int r2 = s.start + (act ? x : 0);
      const int j = a.rows[r2] - s.b * a.L;
      const float* kr = a.kv + (int64_t)r2 * a.ldkv + co;
      float kk[CPL], vv[CPL];
      load_cols(kk, kr);
      load_cols(vv, kr + a.d);
      const float sc = a.scale * group_sum(dot_part(qv, kk)), dp = group_sum(dot_part(dov, vv));
      if (act) {
        const float p = __expf(sc - m) * inv;
        const float mk = ((km >> j) & 1ull) ? a.inv_keep : 0.f;
        const float ds = a.scale * p * (mk * dp - delta);
#pragma unroll
        for (int c = 0; c < CPL; ++c) dq[c] = fmaf(ds, kk[c], dq[c]);
      }
    }
    if (valid) store_cols(a.dq + (int64_t)r * a.d + co, dq);
  }
}

template <int CPL>
__global__ void __launch_bounds__(32 * WARPS) attn_live_bwd_kv_kernel(const LiveArgs a) {
  const int lane = threadIdx.x & 31, grp = lane >> 3, sub = lane & 7;
  const int nwork = a.B * a.h * ((a.L + QC - 1) / QC), nwarps = gridDim.x * WARPS;
  for (int wid = blockIdx.x * WARPS + (threadIdx.x >> 5); wid < nwork; wid += nwarps) {
    SeqInfo s;
    if (!seq_info(a, wid, lane, s)) continue;
    const int co = s.hh * a.dk + sub * CPL;
    const int r2 = s.c0 + grp;
    const bool valid = r2 < s.end;
    const int rk = valid ? r2 : s.end - 1;
    const int j = a.rows[rk] - s.b * a.L;
    const float* kr = a.kv + (int64_t)rk * a.ldkv + co;
    float kk[CPL], vv[CPL], dk_[CPL], dv_[CPL];
    load_cols(kk, kr);
    load_cols(vv, kr + a.d);
#pragma unroll
    for (int c = 0; c < CPL; ++c) dk_[c] = dv_[c] = 0.f;
    const int nq = valid ? s.end - rk : 0, nq_max = warp_max_int(nq);
    for (int x = 0; x < nq_max; ++x) {  // queries at positions i >= j, ascending
      const bool act = x < nq;
      const int r = rk + (act ? x : 0);
      const float delta = a.delta[(int64_t)r * a.h + s.hh];
      const float m = a.stats[((int64_t)r * a.h + s.hh) * 2], inv = a.stats[((int64_t)r * a.h + s.hh) * 2 + 1];
      const unsigned long long km = a.keepw[(int64_t)r * a.h + s.hh];
      float qv[CPL], dov[CPL];
      load_cols(qv, a.q + (int64_t)r * a.ldq + co);
      load_cols(dov, a.dout + (int64_t)r * a.d + co);
      const float sc = a.scale * group_sum(dot_part(qv, kk)), dp = group_sum(dot_part(dov, vv));
      if (act) {
        const float p = __expf(sc - m) * inv;
        const float mk = ((km >> j) & 1ull) ? a.inv_keep : 0.f;
        const float ds = a.scale * p * (mk * dp - delta), pw = mk * p;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          dk_[c] = fmaf(ds, qv[c], dk_[c]);
          dv_[c] = fmaf(pw, dov[c], dv_[c]);
        }
      }
    }
    if (valid) {
      store_cols(a.dkv + (int64_t)r2 * 2 * a.d + co, dk_);
      store_cols(a.dkv + (int64_t)r2 * 2 * a.d + a.d + co, dv_);
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

unsigned live_grid(int B, int h, int L) {
  const int64_t blocks = rbm_cdiv((int64_t)B * h * ((L + QC - 1) / QC), WARPS), cap = (int64_t)RBM_NUM_SMS * 12;
  return (unsigned)(blocks < cap ? blocks : cap);
}

int check_args(const char* who, int B, int L, int h, int dk, float p) {
  if (B < 0 || L < 1 || L > 64 || h < 1 || !(dk == 16 || dk == 32 || dk == 64 || dk == 128)) {
    rbm_set_error("%s: need 1 <= L <= 64 and d_k in {16, 32, 64, 128} (L=%d d_k=%d)", who, L, dk);
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
  const unsigned grid = live_grid(B, h, L);
  if (dk == 16) attn_live_fwd_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 32) attn_live_fwd_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 64) attn_live_fwd_kernel<8><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_fwd_kernel<16><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_fwd");
  return 0;
}

extern "C" int rbm_attn_live_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* bkv, const int32_t* rows,
                                 const int32_t* seq_start, const int64_t* tok, const float* out, const float* stats, const float* dout,
                                 float* dq, float* dkv, float* dead, float* delta, uint64_t* keepw, int B, int L, int h, int dk, float scale,
                                 float p, uint64_t seed, uint64_t site, rbm_stream_t stream) {
  RBM_REQUIRE(q && kv && bkv && rows && seq_start && tok && out && stats && dout && dq && dkv && dead && delta && keepw, "rbm_attn_live_bwd: null pointer");
  if (check_args("rbm_attn_live_bwd", B, L, h, dk, p)) return -1;
  if (B == 0) return 0;
  LiveArgs a{};
  a.q = q; a.ldq = ldq; a.kv = kv; a.ldkv = ldkv; a.bkv = bkv; a.rows = rows; a.seq_start = seq_start; a.tok = tok;
  a.out = const_cast<float*>(out); a.stats = const_cast<float*>(stats); a.dout = dout; a.dq = dq; a.dkv = dkv; a.dead = dead; a.delta = delta; a.keepw = reinterpret_cast<unsigned long long*>(keepw);
  a.B = B; a.L = L; a.h = h; a.dk = dk; a.d = h * dk; a.scale = scale; a.inv_keep = 1.f / (1.f - p); a.thr16 = rbm_drop_threshold16(p);
  a.seed = seed; a.site = site;
  const unsigned grid = live_grid(B, h, L);
  if (dk == 16) attn_live_bwd_q_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 32) attn_live_bwd_q_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 64) attn_live_bwd_q_kernel<8><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_bwd_q_kernel<16><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_bwd(dq)");
  if (dk == 16) attn_live_bwd_kv_kernel<2><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 32) attn_live_bwd_kv_kernel<4><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else if (dk == 64) attn_live_bwd_kv_kernel<8><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  else attn_live_bwd_kv_kernel<16><<<grid, 32 * WARPS, 0, (cudaStream_t)stream>>>(a);
  RBM_LAUNCH_CHECK("rbm_attn_live_bwd(dkv)");
  return 0;
}

RBM_DEFINE_STEP_PTR_SETTER(rbm_step_ptr_set_attention_live)
