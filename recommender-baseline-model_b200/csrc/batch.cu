// batch.cu -- device-side construction of training batches from a CSR of user histories.
//   rbm_bert_cloze_batch   BertTrainDataset.__getitem__  NN/dataloaders/bert.py:77-110  (Cloze masking, 80/10/10, left padding)
//   rbm_sas_train_batch    sample_function / random_neq  NN/dataloaders/sas.py:65-86    (seq / pos shift, uniform negatives
//                                                                                        outside the user's own items)
//   rbm_negative_samples   Random/PopularNegativeSampler.generate_negative_samples
//                                                        NN/dataloaders/negative_samplers/{random.py:13-37,popular.py:15-44}
//   rbm_eval_batch         BertEvalDataset / SASEvalDataset.__getitem__  NN/dataloaders/bert.py:128-142, sas.py:136-153
// Integer work, HBM-bound and tiny next to the step; the point is that the python per-item loops, deepcopy and
// mp.Queue pickling of the reference disappear from the end-to-end path.  Randomness: Philox4x32-10 with
// key = seed, counter = (idx4, site) (common.cuh), 32-bit fields:
//   Cloze: output column p of batch row b uses call idx4 = b*128 + (p >> 1); words 2*(p & 1) (mask decision) and
//          2*(p & 1) + 1 (replacement item).  decision u: u < thr -> scored position (label = item) and
//          u < t80 -> [MASK], u < t90 -> item 1 + ((w * num_items) >> 32), else unchanged;  with X = mask_prob * 2^32:
//          thr = ceil(X), t80 = ceil(0.8 X), t90 = ceil(0.9 X)   (the reference's prob < p, prob/p < 0.8, < 0.9 chain
//          for prob = u / 2^32).
//   SAS:   column p of row b uses call idx4 = b*64 + (p >> 2), word p & 3: rank i = (w * len_a) >> 32 into the ASCENDING
//          list a of ids in [0, num_items] that are not in the user's window (the reference's random_neq with l = 0: id 0
//          is a legal negative there, and so it is here).
//   Negatives: user u's attempt t (t = 0, 1, ...) uses call idx4 = u * 65536 + t; r64 = (word0 << 32) | word1.
//          uniform:    item = 1 + mulhi64(r64, num_items)            (random.py:29: np.random.choice(item_count) + 1)
//          popularity: x = mulhi64(r64, cdf[num_items - 1]), item = 1 + (first i with cdf[i] > x), cdf = inclusive prefix sums
//                      of the interaction counts of items 1..num_items (popular.py:18-22: p = count / total)
//          an attempt is accepted unless the item is in the user's seen set or was already accepted; the first
//          `n_samples` accepted items, in order, are the row (the reference's while-loops keep exactly that sub-sequence of
//          their draws).  Rows that cannot be filled within 65536 attempts end in -1.
#include <math.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t word_of(const uint4& r, int w) { return w == 0 ? r.x : (w == 1 ? r.y : (w == 2 ? r.z : r.w)); }

__global__ void __launch_bounds__(128) bert_cloze_batch_kernel(const int64_t* __restrict__ hist_ptr, const int64_t* __restrict__ hist_items,
                                                               const int64_t* __restrict__ users, int B, int L, uint64_t thr, uint64_t t80,
                                                               uint64_t t90, int64_t mask_token, int64_t num_items, uint64_t seed,
                                                               uint64_t site, int64_t* __restrict__ tokens, int64_t* __restrict__ labels) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int64_t u = users[b];
  const int64_t beg = hist_ptr[u], n = hist_ptr[u + 1] - beg;
  const int nt = (int)(n < L ? n : L);  // the last nt items of the history
  const int pad = L - nt;
  for (int p = threadIdx.x; p < L; p += blockDim.x) {
    int64_t tok = 0, lab = 0;
    if (p >= pad) {
      const int64_t s = hist_items[beg + n - nt + (p - pad)];
      const uint4 r = rbm_philox(seed, site, (uint64_t)b * 128 + (uint64_t)(p >> 1));
      const uint64_t dec = word_of(r, 2 * (p & 1)), w = word_of(r, 2 * (p & 1) + 1);
      tok = s;
      if (dec < thr) {
        lab = s;
        if (dec < t80) tok = mask_token;
        else if (dec < t90) tok = 1 + (int64_t)((w * (uint64_t)num_items) >> 32);
      }
    }
    tokens[(int64_t)b * L + p] = tok;
    labels[(int64_t)b * L + p] = lab;
  }
}

constexpr int SAS_MAXL = 256;

__global__ void __launch_bounds__(128) sas_train_batch_kernel(const int64_t* __restrict__ hist_ptr, const int64_t* __restrict__ hist_items,
                                                              const int64_t* __restrict__ users, int B, int L, int64_t num_items,
                                                              uint64_t seed, uint64_t site, int64_t* __restrict__ seq,
                                                              int64_t* __restrict__ pos, int64_t* __restrict__ neg) {
  __shared__ int64_t train[SAS_MAXL];   // the window in history order
  __shared__ int64_t sorted[SAS_MAXL];  // ... sorted, then unique-compacted in place
  __shared__ int m_uniq;
  const int b = blockIdx.x;
  if (b >= B) return;
  const int64_t u = users[b];
  const int64_t beg = hist_ptr[u], n = hist_ptr[u + 1] - beg;
  const int nt = (int)(n < L ? n : L);
  for (int k = threadIdx.x; k < SAS_MAXL; k += blockDim.x) {
    const int64_t v = k < nt ? hist_items[beg + n - nt + k] : INT64_MAX;
    train[k] = v;
    sorted[k] = v;
  }
  __syncthreads();
  // bitonic sort of 256 keys by 128 threads
  for (int size = 2; size <= SAS_MAXL; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int t = threadIdx.x;
      const int i = 2 * t - (t & (stride - 1));  // lower index of the pair
      const int j = i + stride;
      const bool up = (i & size) == 0;
      const int64_t a = sorted[i], c = sorted[j];
      if ((a > c) == up) { sorted[i] = c; sorted[j] = a; }
      __syncthreads();
    }
  if (threadIdx.x == 0) {  // unique-compact (<= 256 keys)
    int m = 0;
    for (int k = 0; k < nt; ++k)
      if (k == 0 || sorted[k] != sorted[k - 1]) sorted[m++] = sorted[k];
    m_uniq = m;
  }
  __syncthreads();
  const int m = m_uniq;
  // ids of the window that lie inside [0, num_items] shrink the candidate list (ids are sorted: count by scan)
  int inside = 0;
  for (int k = 0; k < m; ++k) inside += (sorted[k] >= 0 && sorted[k] <= num_items) ? 1 : 0;
  const uint64_t len_a = (uint64_t)(num_items + 1 - inside);
  const int pad = L - nt + 1;  // NN/dataloaders/sas.py:68: one more than the free slots, the last item is only a target
  for (int p = threadIdx.x; p < L; p += blockDim.x) {
    int64_t sv = 0, pv = 0, nv = 0;
    if (nt >= 2 && p >= pad) {
      sv = train[p - pad];
      pv = train[p - pad + 1];
      if (len_a > 0) {
        const uint4 r = rbm_philox(seed, site, (uint64_t)b * 64 + (uint64_t)(p >> 2));
        int64_t x = (int64_t)(((uint64_t)word_of(r, p & 3) * len_a) >> 32);  // rank into the ascending complement
        for (int k = 0; k < m; ++k) {
          if (sorted[k] <= x) ++x; else break;
        }
        nv = x;
      }
    }
    seq[(int64_t)b * L + p] = sv;
    pos[(int64_t)b * L + p] = pv;
    neg[(int64_t)b * L + p] = nv;
  }
}

constexpr int NEG_MAX_ATTEMPTS = 65536;

__global__ void __launch_bounds__(128) negative_samples_kernel(const int64_t* __restrict__ seen_ptr, const int64_t* __restrict__ seen_items,
                                                               const uint64_t* __restrict__ cdf, int64_t num_users, int64_t user0,
                                                               int64_t num_items, int n_samples, uint64_t seed, uint64_t site,
                                                               int64_t* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= num_users) return;
  const int64_t u = user0 + row;
  const int64_t* seen = seen_items + seen_ptr[u];
  const int64_t n_seen = seen_ptr[u + 1] - seen_ptr[u];
  int64_t* mine = out + row * n_samples;
  const uint64_t total = cdf ? cdf[num_items - 1] : 0;
  int got = 0;
  for (int t = 0; t < NEG_MAX_ATTEMPTS && got < n_samples; ++t) {
    const uint4 r = rbm_philox(seed, site, (uint64_t)u * NEG_MAX_ATTEMPTS + (uint64_t)t);
    const uint64_t r64 = ((uint64_t)r.x << 32) | (uint64_t)r.y;
    int64_t item;
    if (cdf) {
      const uint64_t x = __umul64hi(r64, total);
      int64_t lo = 0, hi = num_items - 1;  // first index with cdf[i] > x (exists: x < total = cdf[num_items - 1])
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (cdf[mid] > x) hi = mid; else lo = mid + 1;
      }
      item = 1 + lo;
    } else {
      item = 1 + (int64_t)__umul64hi(r64, (uint64_t)num_items);
    }
    // in the (ascending) seen set?
    int64_t lo = 0, hi = n_seen;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (seen[mid] < item) lo = mid + 1; else hi = mid;
    }
    bool reject = lo < n_seen && seen[lo] == item;
    for (int k = 0; k < got && !reject; ++k) reject = mine[k] == item;
    if (!reject) mine[got++] = item;
  }
  for (; got < n_samples; ++got) mine[got] = -1;
}

__global__ void __launch_bounds__(128) eval_batch_kernel(const int64_t* __restrict__ hist_ptr, const int64_t* __restrict__ hist_items,
                                                         const int64_t* __restrict__ answers, const int64_t* __restrict__ negatives,
                                                         const int64_t* __restrict__ users, int B, int L, int n_neg, int64_t mask_token,
                                                         int64_t* __restrict__ seq, int64_t* __restrict__ cand, int64_t* __restrict__ labels) {
  const int b = blockIdx.x;
  if (b >= B) return;
  const int64_t u = users[b];
  const int64_t beg = hist_ptr[u], n = hist_ptr[u + 1] - beg;
  const int extra = mask_token >= 0 ? 1 : 0;  // BERT appends [MASK] before the cut to the last L entries
  const int64_t total = n + extra;
  const int nt = (int)(total < L ? total : L);
  const int pad = L - nt;
  for (int p = threadIdx.x; p < L; p += blockDim.x) {
    int64_t v = 0;
    if (p >= pad) {
      const int64_t k = total - nt + (p - pad);  // index into history ++ [MASK]
      v = k < n ? hist_items[beg + k] : mask_token;
    }
    seq[(int64_t)b * L + p] = v;
  }
  for (int c = threadIdx.x; c <= n_neg; c += blockDim.x) {
    cand[(int64_t)b * (n_neg + 1) + c] = c == 0 ? answers[u] : negatives[u * n_neg + (c - 1)];
    labels[(int64_t)b * (n_neg + 1) + c] = c == 0 ? 1 : 0;
  }
}

}  // namespace

extern "C" int rbm_bert_cloze_batch(const int64_t* hist_ptr, const int64_t* hist_items, const int64_t* users, int B, int L,
                                    double mask_prob, int64_t mask_token, int64_t num_items, uint64_t seed, uint64_t site,
                                    int64_t* tokens, int64_t* labels, rbm_stream_t stream) {
  RBM_REQUIRE(hist_ptr && hist_items && users && tokens && labels, "rbm_bert_cloze_batch: null pointer");
  RBM_REQUIRE(B >= 0 && L >= 1 && L <= 256, "rbm_bert_cloze_batch: need B >= 0, 1 <= L <= 256 (got B=%d L=%d)", B, L);
  RBM_REQUIRE(mask_prob >= 0.0 && mask_prob <= 1.0 && num_items >= 1, "rbm_bert_cloze_batch: mask_prob in [0,1], num_items >= 1");
  if (B == 0) return 0;
  // integer thresholds of the reference's float chain  prob < p,  prob/p < 0.8,  prob/p < 0.9  with prob = u / 2^32
  const double X = mask_prob * 4294967296.0;
  const uint64_t thr = (uint64_t)ceil(X), t80 = (uint64_t)ceil(0.8 * X), t90 = (uint64_t)ceil(0.9 * X);
  bert_cloze_batch_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(hist_ptr, hist_items, users, B, L, thr, t80, t90, mask_token, num_items, seed,
                                                              site, tokens, labels);
  RBM_LAUNCH_CHECK("rbm_bert_cloze_batch");
  return 0;
}

extern "C" int rbm_sas_train_batch(const int64_t* hist_ptr, const int64_t* hist_items, const int64_t* users, int B, int L,
                                   int64_t num_items, uint64_t seed, uint64_t site, int64_t* seq, int64_t* pos, int64_t* neg,
                                   rbm_stream_t stream) {
  RBM_REQUIRE(hist_ptr && hist_items && users && seq && pos && neg, "rbm_sas_train_batch: null pointer");
  RBM_REQUIRE(B >= 0 && L >= 1 && L <= SAS_MAXL, "rbm_sas_train_batch: need B >= 0, 1 <= L <= 256 (got B=%d L=%d)", B, L);
  RBM_REQUIRE(num_items >= 1, "rbm_sas_train_batch: num_items >= 1");
  if (B == 0) return 0;
  sas_train_batch_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(hist_ptr, hist_items, users, B, L, num_items, seed, site, seq, pos, neg);
  RBM_LAUNCH_CHECK("rbm_sas_train_batch");
  return 0;
}

extern "C" int rbm_negative_samples(const int64_t* seen_ptr, const int64_t* seen_items, const uint64_t* pop_cdf, int64_t user_begin,
                                    int64_t num_users, int64_t num_items, int n_samples, uint64_t seed, uint64_t site, int64_t* out,
                                    rbm_stream_t stream) {
  RBM_REQUIRE(seen_ptr && seen_items && out, "rbm_negative_samples: null pointer");
  RBM_REQUIRE(num_users >= 0 && user_begin >= 0 && num_items >= 1 && n_samples >= 1,
              "rbm_negative_samples: need num_users >= 0, num_items >= 1, n_samples >= 1");
  RBM_REQUIRE((uint64_t)(user_begin + num_users) <= (UINT64_MAX >> 16), "rbm_negative_samples: user index out of the counter range");
  if (num_users == 0) return 0;
  negative_samples_kernel<<<(unsigned)rbm_cdiv(num_users, 128), 128, 0, (cudaStream_t)stream>>>(seen_ptr, seen_items, pop_cdf, num_users,
                                                                                              user_begin, num_items, n_samples, seed, site, out);
  RBM_LAUNCH_CHECK("rbm_negative_samples");
  return 0;
}

extern "C" int rbm_eval_batch(const int64_t* hist_ptr, const int64_t* hist_items, const int64_t* answers, const int64_t* negatives,
                              const int64_t* users, int B, int L, int n_neg, int64_t mask_token, int64_t* seq, int64_t* cand,
                              int64_t* labels, rbm_stream_t stream) {
  RBM_REQUIRE(hist_ptr && hist_items && answers && users && seq && cand && labels, "rbm_eval_batch: null pointer");
  RBM_REQUIRE(n_neg == 0 || negatives, "rbm_eval_batch: negatives missing");
  RBM_REQUIRE(B >= 0 && L >= 1 && n_neg >= 0, "rbm_eval_batch: need B >= 0, L >= 1, n_neg >= 0");
  if (B == 0) return 0;
  eval_batch_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(hist_ptr, hist_items, answers, negatives, users, B, L, n_neg, mask_token, seq, cand,
                                                        labels);
  RBM_LAUNCH_CHECK("rbm_eval_batch");
  return 0;
}
