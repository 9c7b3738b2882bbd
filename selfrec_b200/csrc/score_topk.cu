// (iv) Full-catalog scoring + rated-item mask + top-k   (impl 1: CUDA-core fp32).
//
// Replaces the per-user loop of GraphRecommender.test() base/graph_recommender.py:38-58:
//   candidates = predict(user)            XSimGCL.py:57-60  (user_emb[u] @ item_emb.T)
//   candidates[rated] = -10e8             graph_recommender.py:48-50
//   find_k_largest(max_N, candidates)     util/algorithm.py:144-156
//
// One CTA scores 32 users against the whole catalogue in tiles of 128 items.  Each score is
// an fp32 fma chain over k = 0..d-1 (exactly the oracle's loop, so scores are bit-identical).
// The warp that computed a user's scores also owns that user's top-k list: one list entry
// per lane, kept sorted by (score desc, id desc); a candidate enters iff score > current
// k-th score (strict, like the reference's heapreplace test) and the last entry -- the
// lexicographically smallest (score, id), i.e. what heapq would pop -- is evicted.  Items
// are visited in id order, so the final SET equals find_k_largest's, ties included.
#include "common.cuh"

namespace srb {

constexpr int TK_TM = 32;   // users per CTA
constexpr int TK_TN = 128;  // items per tile
constexpr float TK_MASKED = -1e9f;  // -10e8

struct TopkArgs {
  const float* user_emb;
  const float* item_emb;
  int32_t n_items;
  const int32_t* users;
  int32_t n_q;
  const int32_t* rated_ptr;
  const int32_t* rated_idx;
  int32_t k;
  int32_t* out_ids;
  float* out_scores;
  const int32_t* q_map;    // optional: output row of query q (fallback path of impl 2)
  const int32_t* n_q_dev;  // optional: device-side query count (<= n_q)
  int32_t q_skip;          // first q_skip queries are handled elsewhere (fast fallback)
};

template <int D>
__global__ void __launch_bounds__(256) score_topk_kernel(const TopkArgs a) {
  extern __shared__ __align__(16) float tk_smem[];
  float (*Us)[TK_TM] = reinterpret_cast<float (*)[TK_TM]>(tk_smem);                   // [D][32] k-major
  float (*Is)[D + 1] = reinterpret_cast<float (*)[D + 1]>(tk_smem + D * TK_TM);       // [128][D+1]
  const int lane = threadIdx.x & 31;
  const int ty = threadIdx.x >> 5;  // warp id: users ty*4 .. ty*4+3 of the CTA tile
  const int q0 = a.q_skip + blockIdx.x * TK_TM;
  const int n_q = a.n_q_dev ? min(*a.n_q_dev, a.n_q) : a.n_q;
  if (q0 >= n_q) return;

  // user tile (gathered by id), transposed to k-major
  for (int e = threadIdx.x; e < TK_TM * (D / 4); e += blockDim.x) {
    const int u = e % TK_TM, k4 = e / TK_TM;
    float4 v = f4_zero();
    if (q0 + u < n_q) v = ldg4(a.user_emb + (size_t)a.users[q0 + u] * D + k4 * 4);
    Us[k4 * 4 + 0][u] = v.x;
    Us[k4 * 4 + 1][u] = v.y;
    Us[k4 * 4 + 2][u] = v.z;
    Us[k4 * 4 + 3][u] = v.w;
  }

  // per-user state of this warp: sorted list entry per lane, mask cursor
  float ls[4];
  int li[4];
  int cur[4], cend[4];
  int ev[4];  // lane l holds rated_idx[cur + l] of user r (reloaded only when the cursor moves)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    ls[r] = -INFINITY;
    li[r] = -1;
    const int q = q0 + ty * 4 + r;
    if (q < n_q && a.rated_ptr) {
      const int u = a.users[q];
      cur[r] = a.rated_ptr[u];
      cend[r] = a.rated_ptr[u + 1];
    } else {
      cur[r] = cend[r] = 0;
    }
    ev[r] = (cur[r] + lane < cend[r]) ? a.rated_idx[cur[r] + lane] : 0x7fffffff;
  }
  const int K = a.k;

  for (int n0 = 0; n0 < a.n_items; n0 += TK_TN) {
    __syncthreads();
    for (int e = threadIdx.x; e < TK_TN * (D / 4); e += blockDim.x) {
      const int row = e / (D / 4), c4 = e % (D / 4);
      float4 v = f4_zero();
      if (n0 + row < a.n_items) v = ldg4(a.item_emb + (size_t)(n0 + row) * D + c4 * 4);
      Is[row][c4 * 4 + 0] = v.x;
      Is[row][c4 * 4 + 1] = v.y;
      Is[row][c4 * 4 + 2] = v.z;
      Is[row][c4 * 4 + 3] = v.w;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
#pragma unroll 8
    for (int k = 0; k < D; ++k) {
      const float4 uv = *reinterpret_cast<const float4*>(&Us[k][ty * 4]);
      const float ur[4] = {uv.x, uv.y, uv.z, uv.w};
      float iv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) iv[c] = Is[lane + 32 * c][k];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[r][c] = fmaf(ur[r], iv[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      // items past the end of the catalogue can never enter
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (n0 + lane + 32 * c >= a.n_items) s[r][c] = -INFINITY;
      // rated-item mask: walk this user's sorted rated list through the tile
      while (true) {
        const unsigned in = __ballot_sync(SRB_FULL_MASK, ev[r] < n0 + TK_TN);
        const int cnt = __popc(in);
        if (cnt == 0) break;
        for (int t = 0; t < cnt; ++t) {
          const int et = __shfl_sync(SRB_FULL_MASK, ev[r], t) - n0;
          if (et >= 0 && (et & 31) == lane) {
            const int c = et >> 5;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              if (cc == c) s[r][cc] = TK_MASKED;
          }
        }
        cur[r] += cnt;
        ev[r] = (cur[r] + lane < cend[r]) ? a.rated_idx[cur[r] + lane] : 0x7fffffff;
        if (cnt < 32) break;
      }
      // sequential (id-ordered) insertion, 32 candidates per round
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float sc = s[r][c];
        const int id = n0 + lane + 32 * c;
        float thr = __shfl_sync(SRB_FULL_MASK, ls[r], K - 1);
        unsigned m = __ballot_sync(SRB_FULL_MASK, sc > thr);
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float cs = __shfl_sync(SRB_FULL_MASK, sc, src);
          const int cid = __shfl_sync(SRB_FULL_MASK, id, src);
          thr = __shfl_sync(SRB_FULL_MASK, ls[r], K - 1);
          if (cs > thr) {
            const int pos = __popc(__ballot_sync(SRB_FULL_MASK, lane < K && ls[r] > cs));
            const float ps = __shfl_up_sync(SRB_FULL_MASK, ls[r], 1);
            const int pi = __shfl_up_sync(SRB_FULL_MASK, li[r], 1);
            if (lane > pos && lane < K) ls[r] = ps, li[r] = pi;
            if (lane == pos) ls[r] = cs, li[r] = cid;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = q0 + ty * 4 + r;
    if (q < n_q && lane < K) {
      const size_t orow = a.q_map ? (size_t)a.q_map[q] : (size_t)q;
      a.out_ids[orow * K + lane] = li[r];
      a.out_scores[orow * K + lane] = ls[r];
    }
  }
}

// top-k of precomputed score rows (models whose predict() is not a single dot product,
// SURVEY 8b): one warp per row, same sequential insertion rule as above.
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* scores, int n_q, int n_items, int k, int32_t* out_ids,
                                                        float* out_scores) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_q) return;
  float ls = -INFINITY;
  int li = -1;
  const float* row = scores + (size_t)q * n_items;
  for (int n0 = 0; n0 < n_items; n0 += 32) {
    const int id = n0 + lane;
    const float sc = (id < n_items) ? row[id] : -INFINITY;
    float thr = __shfl_sync(SRB_FULL_MASK, ls, k - 1);
    unsigned m = __ballot_sync(SRB_FULL_MASK, sc > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float cs = __shfl_sync(SRB_FULL_MASK, sc, src);
      const int cid = __shfl_sync(SRB_FULL_MASK, id, src);
      thr = __shfl_sync(SRB_FULL_MASK, ls, k - 1);
      if (cs > thr) {
        const int pos = __popc(__ballot_sync(SRB_FULL_MASK, lane < k && ls > cs));
        const float ps = __shfl_up_sync(SRB_FULL_MASK, ls, 1);
        const int pi = __shfl_up_sync(SRB_FULL_MASK, li, 1);
        if (lane > pos && lane < k) ls = ps, li = pi;
        if (lane == pos) ls = cs, li = cid;
      }
    }
  }
  if (lane < k) {
    out_ids[(size_t)q * k + lane] = li;
    out_scores[(size_t)q * k + lane] = ls;
  }
}

// dense score rows out[q][i] = <user_emb[users[q]], item_emb[i]> (same fma chain as above):
// the reference's predict() (XSimGCL.py:57-60) for callers that want the raw vector.
template <int D>
__global__ void __launch_bounds__(256) score_rows_kernel(const float* user_emb, const float* item_emb, const int32_t* users,
                                                         int n_items, float* out) {
  __shared__ float us[D];
  const int q = blockIdx.y;
  for (int k = threadIdx.x; k < D; k += blockDim.x) us[k] = user_emb[(size_t)users[q] * D + k];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const float* it = item_emb + (size_t)i * D;
  float acc = 0.f;
#pragma unroll 8
  for (int k4 = 0; k4 < D / 4; ++k4) {
    const float4 v = ldg4(it + k4 * 4);
    acc = fmaf(us[k4 * 4 + 0], v.x, acc);
    acc = fmaf(us[k4 * 4 + 1], v.y, acc);
    acc = fmaf(us[k4 * 4 + 2], v.z, acc);
    acc = fmaf(us[k4 * 4 + 3], v.w, acc);
  }
  out[(size_t)q * n_items + i] = acc;
}

int score_topk_tc(const srb_topk_desc* d, cudaStream_t st);  // score_topk_tc.cu

template <int D>
static int launch_topk(const TopkArgs& a, cudaStream_t st) {
  const size_t smem = sizeof(float) * (D * TK_TM + TK_TN * (D + 1));
  static bool attr_done = false;
  if (!attr_done) {
    SRB_TRY(check_cuda(cudaFuncSetAttribute(score_topk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "topk smem attr"));
    attr_done = true;
  }
  const int blocks = (a.n_q - a.q_skip + TK_TM - 1) / TK_TM;
  if (blocks <= 0) return SRB_OK;
  score_topk_kernel<D><<<blocks, 256, smem, st>>>(a);
  return post_launch("score_topk_kernel");
}

// ---- fallback of impl 2: exact re-run of the users it could not certify (device-side list) ----
// Fast path for the first `cap` of them: exact score rows spread over many CTAs + one warp per user
// for the sequential top-k (a handful of users must not cost a full impl-1 block pass, ~1.6 ms).
template <int D>
__global__ void __launch_bounds__(256) fb_score_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
                                                      const int32_t* __restrict__ fb_users, const int32_t* __restrict__ fb_count,
                                                      const int32_t* __restrict__ rated_ptr, const int32_t* __restrict__ rated_idx,
                                                      int n_items, float* __restrict__ scratch, int cap) {
  // the usual case is zero fallback users: a small grid.y that strides over the slots keeps that case cheap
  const int count = min(*fb_count, cap);
  __shared__ float us[D];
  for (int slot = blockIdx.y; slot < count; slot += gridDim.y) {
    const int u = fb_users[slot];
    __syncthreads();
    for (int k = threadIdx.x; k < D; k += blockDim.x) us[k] = user_emb[(size_t)u * D + k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) continue;
    const float* it = item_emb + (size_t)i * D;
    float acc = 0.f;
#pragma unroll 8
    for (int k4 = 0; k4 < D / 4; ++k4) {
      const float4 v = ldg4(it + k4 * 4);
      acc = fmaf(us[k4 * 4 + 0], v.x, acc);
      acc = fmaf(us[k4 * 4 + 1], v.y, acc);
      acc = fmaf(us[k4 * 4 + 2], v.z, acc);
      acc = fmaf(us[k4 * 4 + 3], v.w, acc);
    }
    if (rated_ptr) {  // binary search of item i in the user's sorted rated list
      int lo = rated_ptr[u], hi = rated_ptr[u + 1];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int v = rated_idx[mid];
        if (v < i) lo = mid + 1; else hi = mid;
      }
      if (lo < rated_ptr[u + 1] && rated_idx[lo] == i) acc = TK_MASKED;
    }
    scratch[(size_t)slot * n_items + i] = acc;
  }
}

__global__ void __launch_bounds__(256) fb_topk_kernel(const float* scratch, const int32_t* fb_rows, const int32_t* fb_count, int cap,
                                                      int n_items, int k, int32_t* out_ids, float* out_scores) {
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (slot >= cap || slot >= *fb_count) return;
  float ls = -INFINITY;
  int li = -1;
  const float* row = scratch + (size_t)slot * n_items;
  for (int n0 = 0; n0 < n_items; n0 += 32) {
    const int id = n0 + lane;
    const float sc = (id < n_items) ? row[id] : -INFINITY;
    float thr = __shfl_sync(SRB_FULL_MASK, ls, k - 1);
    unsigned m = __ballot_sync(SRB_FULL_MASK, sc > thr);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const float cs = __shfl_sync(SRB_FULL_MASK, sc, src);
      const int cid = __shfl_sync(SRB_FULL_MASK, id, src);
      thr = __shfl_sync(SRB_FULL_MASK, ls, k - 1);
      if (cs > thr) {
        const int pos = __popc(__ballot_sync(SRB_FULL_MASK, lane < k && ls > cs));
        const float ps = __shfl_up_sync(SRB_FULL_MASK, ls, 1);
        const int pi = __shfl_up_sync(SRB_FULL_MASK, li, 1);
        if (lane > pos && lane < k) ls = ps, li = pi;
        if (lane == pos) ls = cs, li = cid;
      }
    }
  }
  if (lane < k) {
    const size_t orow = (size_t)fb_rows[slot];
    out_ids[orow * k + lane] = li;
    out_scores[orow * k + lane] = ls;
  }
}

int score_topk_fallback(const srb_topk_desc* d, const int32_t* fb_users, const int32_t* fb_rows, const int32_t* fb_count,
                        float* scratch, int fb_cap, cudaStream_t st) {
  // fast path: up to fb_cap users
  dim3 grid((d->n_items + 255) / 256, fb_cap < 8 ? fb_cap : 8);
  fb_score_kernel<64><<<grid, 256, 0, st>>>(d->user_emb, d->item_emb, fb_users, fb_count, d->rated_ptr, d->rated_idx, d->n_items,
                                            scratch, fb_cap);
  SRB_TRY(post_launch("fb_score_kernel"));
  fb_topk_kernel<<<(fb_cap + 7) / 8, 256, 0, st>>>(scratch, fb_rows, fb_count, fb_cap, d->n_items, d->k, d->out_ids, d->out_scores);
  SRB_TRY(post_launch("fb_topk_kernel"));
  // slow path: everyone beyond fb_cap goes through the impl-1 kernel (CTAs without work exit at once)
  if (d->n_q <= fb_cap) return SRB_OK;
  TopkArgs a;
  a.user_emb = d->user_emb;
  a.item_emb = d->item_emb;
  a.n_items = d->n_items;
  a.users = fb_users;
  a.n_q = d->n_q;
  a.rated_ptr = d->rated_ptr;
  a.rated_idx = d->rated_idx;
  a.k = d->k;
  a.out_ids = d->out_ids;
  a.out_scores = d->out_scores;
  a.q_map = fb_rows;
  a.n_q_dev = fb_count;
  a.q_skip = fb_cap;
  return launch_topk<64>(a, st);
}

}  // namespace srb

extern "C" int srb_score_topk(const srb_topk_desc* d, void* stream) {
  SRB_REQUIRE(d != nullptr, "topk: null desc");
  SRB_REQUIRE(d->n_q >= 0, "topk: negative n_q");
  if (d->n_q == 0) return SRB_OK;  // empty query list: nothing to launch (pointers may be null)
  SRB_REQUIRE(d->user_emb && d->item_emb && d->users && d->out_ids && d->out_scores, "topk: null pointer");
  SRB_REQUIRE((d->rated_ptr == nullptr) == (d->rated_idx == nullptr), "topk: rated_ptr/rated_idx must both be set or both null");
  SRB_REQUIRE(d->k >= 1 && d->k <= 32, "topk: k=%d unsupported (1..32)", d->k);
  SRB_REQUIRE(d->n_items >= 1 && d->n_q >= 0, "topk: bad shape");
  SRB_REQUIRE(d->impl >= 0 && d->impl <= 2, "topk: bad impl");
  if (d->n_q == 0) return SRB_OK;
  if (d->impl == 2 || (d->impl == 0 && d->d == 64 && d->workspace != nullptr && d->n_items >= 1024))
    return srb::score_topk_tc(d, (cudaStream_t)stream);
  srb::TopkArgs a;
  a.user_emb = d->user_emb;
  a.item_emb = d->item_emb;
  a.n_items = d->n_items;
  a.users = d->users;
  a.n_q = d->n_q;
  a.rated_ptr = d->rated_ptr;
  a.rated_idx = d->rated_idx;
  a.k = d->k;
  a.out_ids = d->out_ids;
  a.out_scores = d->out_scores;
  a.q_map = nullptr;
  a.n_q_dev = nullptr;
  a.q_skip = 0;
  switch (d->d) {
    case 32: return srb::launch_topk<32>(a, (cudaStream_t)stream);
    case 64: return srb::launch_topk<64>(a, (cudaStream_t)stream);
    case 128: return srb::launch_topk<128>(a, (cudaStream_t)stream);
    default: srb::set_error("topk: unsupported d=%d (32, 64, 128)", d->d); return SRB_ERR_ARG;
  }
}

extern "C" int srb_topk_rows(const float* scores, int32_t n_q, int32_t n_items, int32_t k, int32_t* out_ids,
                             float* out_scores, void* stream) {
  SRB_REQUIRE(scores && out_ids && out_scores, "topk_rows: null pointer");
  SRB_REQUIRE(k >= 1 && k <= 32, "topk_rows: k=%d unsupported (1..32)", k);
  SRB_REQUIRE(n_q >= 0 && n_items >= 1, "topk_rows: bad shape");
  if (n_q == 0) return SRB_OK;
  srb::topk_rows_kernel<<<(n_q + 7) / 8, 256, 0, (cudaStream_t)stream>>>(scores, n_q, n_items, k, out_ids, out_scores);
  return srb::post_launch("topk_rows_kernel");
}

extern "C" int srb_score_rows(const float* user_emb, const float* item_emb, int32_t d, const int32_t* users, int32_t n_q,
                              int32_t n_items, float* out, void* stream) {
  SRB_REQUIRE(user_emb && item_emb && users && out, "score_rows: null pointer");
  SRB_REQUIRE(n_q >= 0 && n_q <= 65535 && n_items >= 1, "score_rows: bad shape (n_q <= 65535)");
  if (n_q == 0) return SRB_OK;
  dim3 grid((n_items + 255) / 256, n_q);
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 32: srb::score_rows_kernel<32><<<grid, 256, 0, st>>>(user_emb, item_emb, users, n_items, out); break;
    case 64: srb::score_rows_kernel<64><<<grid, 256, 0, st>>>(user_emb, item_emb, users, n_items, out); break;
    case 128: srb::score_rows_kernel<128><<<grid, 256, 0, st>>>(user_emb, item_emb, users, n_items, out); break;
    default: srb::set_error("score_rows: unsupported d=%d (32, 64, 128)", d); return SRB_ERR_ARG;
  }
  return srb::post_launch("score_rows_kernel");
}

namespace srb {
// one warp per query row: lane r (and r + 32) looks its recommended id up in the user's sorted test list
__global__ void __launch_bounds__(256) rank_hit_masks_kernel(const int32_t* __restrict__ ids, int n_q, int k, const int32_t* __restrict__ users,
                                                             const int32_t* __restrict__ test_ptr, const int32_t* __restrict__ test_idx,
                                                             unsigned long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= n_q) return;
  const int u = users[q];
  const int beg = test_ptr[u], end = test_ptr[u + 1];
  unsigned long long mask = 0;
  for (int half = 0; half < 2; ++half) {
    const int r = half * 32 + lane;
    bool hit = false;
    if (r < k) {
      const int id = ids[(size_t)q * k + r];
      int lo = beg, hi = end;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (test_idx[mid] < id) lo = mid + 1;
        else hi = mid;
      }
      hit = lo < end && test_idx[lo] == id;
    }
    mask |= (unsigned long long)__ballot_sync(SRB_FULL_MASK, hit) << (32 * half);
  }
  if (lane == 0) out[q] = mask;
}
}  // namespace srb

extern "C" int srb_rank_hit_masks(const int32_t* topk_ids, int32_t n_q, int32_t k, const int32_t* users, const int32_t* test_ptr,
                                  const int32_t* test_idx, uint64_t* hit_mask, void* stream) {
  SRB_REQUIRE(n_q >= 0 && k >= 1 && k <= 64, "rank_hit_masks: k must be 1..64");
  if (n_q == 0) return SRB_OK;
  SRB_REQUIRE(topk_ids && users && test_ptr && test_idx && hit_mask, "rank_hit_masks: null pointer");
  srb::rank_hit_masks_kernel<<<(n_q + 7) / 8, 256, 0, (cudaStream_t)stream>>>(topk_ids, n_q, k, users, test_ptr, test_idx,
                                                                              reinterpret_cast<unsigned long long*>(hit_mask));
  return srb::post_launch("rank_hit_masks_kernel");
}
