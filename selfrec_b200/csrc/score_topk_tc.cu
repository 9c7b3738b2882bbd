// (iv) impl 2: full-catalog scoring on the 5th-gen tensor cores (tcgen05, TF32) with fused
// rated-item mask + candidate selection, followed by exact fp32 re-scoring.
//
// Replaces GraphRecommender.test() base/graph_recommender.py:38-58 (predict XSimGCL.py:57-60,
// mask :48-50, find_k_largest util/algorithm.py:144-156), same contract as impl 1.
//
// Stage 1  tc_gather_kernel     users[q] rows -> contiguous [n_q, d] table (TMA cannot gather),
//                               ||u_q||, max_i ||item_i||
// Stage 2  tc_score_kernel      one persistent CTA per block of UB <= 256 users, streaming the whole
//                               catalogue in tiles of 128 items:
//            warp 0   TMA producer: item tiles [128 x 64] fp32, 128B-swizzled, 3-stage mbarrier ring
//            warp 1   MMA issuer : tcgen05.mma kind::tf32, M=128 (users) x N=128 (items) x K=8,
//                                  2 user halves x 8 k-steps per tile, accumulators in TMEM
//                                  (2 stages x 2 halves x 128 columns = all 512 columns)
//            warps 2-17 epilogue : 16 warps (4 per scheduler: the select loop is latency-bound), thread =
//                                  one user row x one 64-column half of the tile: tcgen05.ld (lane = user
//                                  row), rated-item cursor, per-thread top-24 candidate list (3 buckets
//                                  of 8, minima in registers) in shared memory -> 2 x 24 candidates per user
//          The raw fp32 tables are fed to the tensor core, which reads them as TF32 (low 13
//          mantissa bits ignored): scores carry <= 2^-9 ||u|| ||i|| error -- candidates only.
// Stage 3  tc_rescore_kernel    warp per user: exact fp32 fma-chain scores of the 2 x 24 candidates
//                               (bit-identical to impl 1 / the oracle), find_k_largest's sequential
//                               insertion in id order, and a safety test: every non-candidate of a column
//                               half has approx score <= that half's 24th best, so the result is exact iff
//                               thr32 := max of the two + E < exact k-th score.  Users failing it (or with fewer than
//                               k unrated items) are re-run by the exact CUDA-core kernel (impl 1).
#include "common.cuh"
#include "tc_common.cuh"

namespace srb {

using namespace tc;

constexpr int TC_D = 64;          // embedding size handled by this kernel
constexpr int TC_TN = 128;        // items per tile (UMMA N)
constexpr int TC_STAGES = 2;      // smem ring depth (a tile's select takes ~2 us: one tile of prefetch is enough)
constexpr int TC_LIST = 24;       // candidates per list; every user has two lists, one per column half of the tiles
constexpr int TC_CAND = 2 * TC_LIST;
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr uint32_t TC_TILE_BYTES = TC_TN * TC_D * 4;     // 32 KB: 2 k-chunks x [128][32] fp32
constexpr uint32_t TC_USER_BYTES = 2 * 128 * TC_D * 4;   // 64 KB: 2 halves x 2 k-chunks x [128][32]

struct TcSmem {
  // dynamic shared memory, 1024-byte aligned base:
  //   [0, 64K)          user tiles   half h, chunk c at (h*2 + c) * 16 KB
  //   [64K, 64K+64K)    item stages  stage s, chunk c at 64K + s*32K + c*16K
  //   then candidate lists: scores [24][512] f32, ids [24][512] i32   (96 KB)
  //   then barriers
  static constexpr uint32_t users_off = 0;
  static constexpr uint32_t items_off = TC_USER_BYTES;
  static constexpr uint32_t cand_s_off = items_off + TC_STAGES * TC_TILE_BYTES;
  static constexpr uint32_t cand_i_off = cand_s_off + TC_LIST * 512 * 4;
  static constexpr uint32_t bar_off = cand_i_off + TC_LIST * 512 * 4;
  static constexpr uint32_t total = bar_off + 256;
};

struct TcArgs {
  const int32_t* users;      // original user ids per query row (for the rated CSR)
  const int32_t* rated_ptr;
  const int32_t* rated_idx;
  int32_t n_q;
  int32_t n_items;
  int32_t ub;                // users per CTA (<= 256)
  float* cand_s;             // [n_q][2][24] approx scores
  int32_t* cand_i;           // [n_q][2][24]
  int32_t* cand_n;           // [n_q][2]
  float* cand_thr;           // [n_q][2] min approx score of a full list, else -inf
};

__global__ void __launch_bounds__(256) tc_gather_kernel(const float* __restrict__ user_emb, const int32_t* __restrict__ users, int n_q,
                                                       int n_q_pad, float* __restrict__ ug, float* __restrict__ unorm,
                                                       const float* __restrict__ item_emb, int n_items, unsigned int* bmax_bits) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  // rows [0, n_q_pad): gathered user rows; rows [n_q_pad, n_q_pad + n_items): item norms
  if (w < n_q_pad) {
    float2 v = make_float2(0.f, 0.f);
    if (w < n_q) v = *reinterpret_cast<const float2*>(user_emb + (size_t)users[w] * TC_D + lane * 2);
    *reinterpret_cast<float2*>(ug + (size_t)w * TC_D + lane * 2) = v;
    const float ss = warp_sum(v.x * v.x + v.y * v.y);
    if (lane == 0 && w < n_q) unorm[w] = sqrtf(ss);
  } else if (w < n_q_pad + n_items) {
    const int i = w - n_q_pad;
    const float2 v = *reinterpret_cast<const float2*>(item_emb + (size_t)i * TC_D + lane * 2);
    const float ss = warp_sum(v.x * v.x + v.y * v.y);
    // non-negative floats order like uints; 38 k atomics on one word serialise, so look before touching it
    const unsigned int bits = __float_as_uint(sqrtf(ss));
    if (lane == 0 && bits > *reinterpret_cast<volatile unsigned int*>(bmax_bits)) atomicMax(bmax_bits, bits);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap tm_users, const __grid_constant__ CUtensorMap tm_items, const TcArgs a) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TcSmem::bar_off);
  uint64_t* bar_full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* bar_empty = bars + TC_STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* bar_tfull = bars + 2 * TC_STAGES;   // [2]       MMA -> epilogue
  uint64_t* bar_tempty = bar_tfull + 2;         // [2]       epilogue -> MMA
  uint64_t* bar_users = bar_tempty + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_users + 1);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * a.ub;             // first query row of this CTA
  const int n_tiles = (a.n_items + TC_TN - 1) / TC_TN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + s, 1);
      mbar_init(bar_tempty + s, TC_EPI_WARPS);  // one arrive per epilogue warp
    }
    mbar_init(bar_users, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      tma_prefetch_desc(&tm_users);
      tma_prefetch_desc(&tm_items);
      mbar_arrive_expect_tx(bar_users, TC_USER_BYTES);
      for (int h = 0; h < 2; ++h)
        for (int c = 0; c < 2; ++c)
          tma_load_2d(sm + TcSmem::users_off + (h * 2 + c) * 16384, &tm_users, bar_users, c * 32, q0 + h * 128);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % TC_STAGES;
        const uint32_t ph = (t / TC_STAGES) & 1;
        mbar_wait(bar_empty + s, ph ^ 1);  // first pass through the ring passes immediately
        mbar_arrive_expect_tx(bar_full + s, TC_TILE_BYTES);
        for (int c = 0; c < 2; ++c)
          tma_load_2d(sm + TcSmem::items_off + s * TC_TILE_BYTES + c * 16384, &tm_items, bar_full + s, c * 32, t * TC_TN);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, TC_TN);
      mbar_wait(bar_users, 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % TC_STAGES;
        const int acc = t & 1;
        mbar_wait(bar_tempty + acc, ((t >> 1) & 1) ^ 1);  // epilogue drained this accumulator stage
        mbar_wait(bar_full + s, (t / TC_STAGES) & 1);
        fence_after_sync();
        const uint32_t items_base = smem_u32(sm + TcSmem::items_off + s * TC_TILE_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t d_tmem = tmem + acc * 256 + h * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t ua = smem_u32(sm + TcSmem::users_off + (h * 2 + c) * 16384);
            const uint32_t ib = items_base + c * 16384;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32_ss(d_tmem, make_smem_desc_k_sw128(ua + k * 32), make_smem_desc_k_sw128(ib + k * 32), idesc, (c | k) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + s);     // smem stage reusable once these MMAs have read it
        umma_commit(bar_tfull + acc);   // accumulators complete
      }
    }
  } else {
    // ===== epilogue: 16 warps, thread = one user row x one 64-column half of every tile =====
    const int e = warp - 2;
    const int half = (e >> 2) & 1;  // user half (accumulator)
    const int chalf = e >> 3;       // column half
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int row = half * 128 + quarter * 32 + lane;
    const int q = q0 + row;
    const bool active = row < a.ub && q < a.n_q;
    const int tix = chalf * 256 + row;  // column in the candidate arrays
    float* cs = reinterpret_cast<float*>(sm + TcSmem::cand_s_off);
    int32_t* ci = reinterpret_cast<int32_t*>(sm + TcSmem::cand_i_off);
    float thr = -INFINITY;
    int cnt = 0;
    // cursor into this user's sorted rated list; the id after next is prefetched so that advancing the
    // cursor never makes the warp wait for a global load
    int cur = 0, cend = 0, next_rated = 0x7fffffff, after_next = 0x7fffffff;
    if (active && a.rated_ptr) {
      const int u = a.users[q];
      cur = a.rated_ptr[u];
      cend = a.rated_ptr[u + 1];
      if (cur < cend) next_rated = a.rated_idx[cur];
      if (cur + 1 < cend) after_next = a.rated_idx[cur + 1];
    }
    // candidate list: 24 (score, id) slots in shared memory (column `tix`), organised as 3 buckets of 8.
    // Registers keep each bucket's minimum, so replacing the global minimum re-scans only one bucket
    // (8 independent shared-memory loads) instead of the whole list.
    // Each stored score carries its slot-in-bucket in the 3 low mantissa bits (the scores only rank candidates;
    // the certificate in tc_rescore_kernel accounts for the 2^-20 relative perturbation), so a bucket's minimum
    // names its own slot and 7 FMNMX replace a compare/select scan.
    float bm0 = INFINITY, bm1 = INFINITY, bm2 = INFINITY;  // bucket minima (valid once full)
    auto process_group = [&](const uint32_t (&r)[32], int g0) {
      uint32_t mask = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j)  // two instructions per score: FSETP + predicated LOP3
        asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}"
            : "+r"(mask)
            : "f"(__uint_as_float(r[j])), "f"(thr), "r"(1u << j));
      if (g0 + 32 > a.n_items) mask &= (g0 < a.n_items) ? (0xffffffffu >> (g0 + 32 - a.n_items)) : 0u;  // zero-filled OOB rows
      if (!active) mask = 0;
      // rated items never become candidates: walk the sorted rated list through this group
      while (next_rated < g0 + 32) {
        if (next_rated >= g0) mask &= ~(1u << (next_rated - g0));
        ++cur;
        next_rated = after_next;
        after_next = (cur + 1 < cend) ? a.rated_idx[cur + 1] : 0x7fffffff;
      }
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1;
        // r[j] with a per-lane j: a 5-level select tree on the bits of j (31 selects)
        uint32_t t16[16], t8[8], t4[4];
#pragma unroll
        for (int i = 0; i < 16; ++i) t16[i] = (j & 1) ? r[2 * i + 1] : r[2 * i];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = (j & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
        for (int i = 0; i < 4; ++i) t4[i] = (j & 4) ? t8[2 * i + 1] : t8[2 * i];
        const uint32_t t2a = (j & 8) ? t4[1] : t4[0], t2b = (j & 8) ? t4[3] : t4[2];
        float sc = __uint_as_float((j & 16) ? t2b : t2a);
        if (!(sc > thr)) continue;  // thr may have risen inside this group
        const int id = g0 + j;
        if (cnt < TC_LIST) {
          cs[cnt * 512 + tix] = sc;
          ci[cnt * 512 + tix] = id;
          ++cnt;
          if (cnt == TC_LIST) {  // list full: tag every slot and establish the bucket minima
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              float mn = INFINITY;
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const float v = __uint_as_float((__float_as_uint(cs[(b * 8 + qq) * 512 + tix]) & ~7u) | (uint32_t)qq);
                cs[(b * 8 + qq) * 512 + tix] = v;
                mn = fminf(mn, v);
              }
              if (b == 0) bm0 = mn;
              if (b == 1) bm1 = mn;
              if (b == 2) bm2 = mn;
            }
            thr = fminf(fminf(bm0, bm1), bm2);
          }
        } else {
          // evict the global minimum: it sits in the bucket whose minimum equals thr, in the slot its tag names
          const int b = (bm0 == thr) ? 0 : ((bm1 == thr) ? 1 : 2);
          const int pos = b * 8 + (int)(__float_as_uint(thr) & 7u);
          cs[pos * 512 + tix] = __uint_as_float((__float_as_uint(sc) & ~7u) | (__float_as_uint(thr) & 7u));
          ci[pos * 512 + tix] = id;
          float mn = cs[(b * 8) * 512 + tix];
#pragma unroll
          for (int qq = 1; qq < 8; ++qq) mn = fminf(mn, cs[(b * 8 + qq) * 512 + tix]);
          if (b == 0) bm0 = mn;
          else if (b == 1) bm1 = mn;
          else bm2 = mn;
          thr = fminf(fminf(bm0, bm1), bm2);
        }
      }
    };
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      mbar_wait(bar_tfull + acc, (t >> 1) & 1);
      fence_after_sync();
      const int n0 = t * TC_TN;
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + acc * 256 + half * 128 + chalf * 64;
      const int c0 = n0 + chalf * 64;
      // two register buffers: the load of the second group is in flight while the first is processed
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(tbase, r0);
      tmem_ld_wait();
      tmem_ld_32x32(tbase + 32, r1);
      process_group(r0, c0);
      tmem_ld_wait();
      fence_before_sync();  // my share of the accumulator is in registers: hand the stage back before selecting
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + acc);
      process_group(r1, c0 + 32);
    }
    if (active) {
      const size_t o = ((size_t)q * 2 + chalf) * TC_LIST;
      for (int p = 0; p < TC_LIST; ++p) {
        a.cand_s[o + p] = (p < cnt) ? cs[p * 512 + tix] : -INFINITY;
        a.cand_i[o + p] = (p < cnt) ? ci[p * 512 + tix] : -1;
      }
      a.cand_n[(size_t)q * 2 + chalf] = cnt;
      a.cand_thr[(size_t)q * 2 + chalf] = (cnt == TC_LIST) ? thr : -INFINITY;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

struct RescoreArgs {
  const float* ug;          // gathered user rows [n_q_pad][64]
  const float* item_emb;
  const float* unorm;
  const unsigned int* bmax_bits;
  const int32_t* users;
  const int32_t* rated_ptr;
  const float* cand_s;
  const int32_t* cand_i;
  const int32_t* cand_n;
  const float* cand_thr;
  int32_t n_q, n_items, k;
  int32_t* out_ids;
  float* out_scores;
  int32_t* fb_count;        // device counter of users needing the exact fallback
  int32_t* fb_rows;         // their query rows
  int32_t* fb_users;        // their user ids
};

__global__ void __launch_bounds__(256) tc_rescore_kernel(const RescoreArgs a) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.n_q) return;
  const int cnt_a = a.cand_n[(size_t)q * 2], cnt_b = a.cand_n[(size_t)q * 2 + 1];
  const int K = a.k;
  const float bmax = __uint_as_float(*a.bmax_bits);
  const float E = (1.0f / 512.0f + 1.0f / 65536.0f + 1.0f / 262144.0f) * a.unorm[q] * bmax;  // TF32 truncation + slot tags
  // ---- prune by approximate score before any exact work ----
  // Every exact score lies within E of its approximate score.  Let a_K be the K-th largest approximate score of the
  // candidates: K candidates have an exact score >= a_K - E, so one whose approximate score is below a_K - 2E is beaten
  // by at least K others and can never end in the top-K, whatever the insertion order.  Typically half of the 48 go,
  // the survivors fit one lane each, and the second exact pass and half of the sequential insertions disappear.
  __shared__ int32_t surv[8][TC_CAND];
  int32_t* sv = surv[threadIdx.x >> 5];
  int cnt;
  {
    float ap[2];
    int cid[2];
    bool have[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = lane + 32 * h;
      have[h] = c < TC_CAND && (c % TC_LIST) < ((c < TC_LIST) ? cnt_a : cnt_b);
      ap[h] = have[h] ? a.cand_s[(size_t)q * TC_CAND + c] : -INFINITY;
      cid[h] = have[h] ? a.cand_i[(size_t)q * TC_CAND + c] : 0x7fffffff;
    }
    float cut = -INFINITY;
    if (cnt_a + cnt_b > K) {
      // descending rank of my approximate scores (ties broken by slot), then the value of rank K-1
      int rk[2] = {0, 0};
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        const float o0 = __shfl_sync(SRB_FULL_MASK, ap[0], l);
        const float o1 = __shfl_sync(SRB_FULL_MASK, ap[1], l);
        rk[0] += (o0 > ap[0] || (o0 == ap[0] && l < lane)) + (o1 > ap[0]);
        rk[1] += (o0 > ap[1] || o0 == ap[1]) + (o1 > ap[1] || (o1 == ap[1] && l < lane));
      }
      const unsigned b0 = __ballot_sync(SRB_FULL_MASK, have[0] && rk[0] == K - 1);
      const unsigned b1 = __ballot_sync(SRB_FULL_MASK, have[1] && rk[1] == K - 1);
      const float ak = b0 ? __shfl_sync(SRB_FULL_MASK, ap[0], __ffs(b0) - 1) : __shfl_sync(SRB_FULL_MASK, ap[1], b1 ? __ffs(b1) - 1 : 0);
      if (b0 | b1) cut = ak - 2.0f * E;
    }
    const unsigned k0 = __ballot_sync(SRB_FULL_MASK, have[0] && ap[0] >= cut);
    const unsigned k1 = __ballot_sync(SRB_FULL_MASK, have[1] && ap[1] >= cut);
    const unsigned lt = (1u << lane) - 1u;
    if ((k0 >> lane) & 1u) sv[__popc(k0 & lt)] = cid[0];
    if ((k1 >> lane) & 1u) sv[__popc(k0) + __popc(k1 & lt)] = cid[1];
    cnt = __popc(k0) + __popc(k1);
    __syncwarp();
  }
  // survivor slots on 32 lanes: lane l owns slots l and l + 32 (the second pass only runs when more than 32 survive)
  int id[2];
  float s[2];
  bool mine[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c = lane + 32 * h;
    mine[h] = c < cnt;
    id[h] = 0x7fffffff;
    s[h] = -INFINITY;
    if (h == 1 && cnt <= 32) continue;  // warp-uniform
    if (mine[h]) id[h] = sv[c];
    // exact score: the same fp32 fma chain over k = 0..63 as impl 1 and the oracle
    if (mine[h]) {
      const float* u = a.ug + (size_t)q * TC_D;
      const float4* it = reinterpret_cast<const float4*>(a.item_emb + (size_t)id[h] * TC_D);
      float4 iv[TC_D / 4];  // the whole row in flight at once: this kernel is bound by the latency of these gathers
#pragma unroll
      for (int k4 = 0; k4 < TC_D / 4; ++k4) iv[k4] = __ldg(it + k4);
      float acc = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < TC_D / 4; ++k4) {
        const float4 uv = *reinterpret_cast<const float4*>(u + k4 * 4);
        acc = fmaf(uv.x, iv[k4].x, acc);
        acc = fmaf(uv.y, iv[k4].y, acc);
        acc = fmaf(uv.z, iv[k4].z, acc);
        acc = fmaf(uv.w, iv[k4].w, acc);
      }
      s[h] = acc;
    }
  }
  // rank of my candidates by item id (ids are distinct; empty slots carry INT_MAX and never count)
  int rank[2] = {0, 0};
#pragma unroll 8
  for (int l = 0; l < 32; ++l) {
    const int o0 = __shfl_sync(SRB_FULL_MASK, id[0], l);
    const int o1 = __shfl_sync(SRB_FULL_MASK, id[1], l);
    rank[0] += (o0 < id[0]) + (o1 < id[0]);
    rank[1] += (o0 < id[1]) + (o1 < id[1]);
  }
  // find_k_largest's sequential process over the candidates in id order (see score_topk.cu).  The candidates
  // are first laid out in id order in shared memory, so that each step is one broadcast load and a
  // warp-uniform compare instead of a ballot / find-first-set / shuffle chain
  __shared__ float2 ord[8][TC_CAND];
  float2* mo = ord[threadIdx.x >> 5];
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (mine[h]) mo[rank[h]] = make_float2(s[h], __int_as_float(id[h]));
  __syncwarp();
  float ls = -INFINITY;
  int li = -1;
  float thr = -INFINITY;  // score of list slot K-1 (warp-uniform)
  for (int t = 0; t < cnt; ++t) {
    const float2 e = mo[t];
    const float cs = e.x;
    if (cs > thr) {
      const int cid = __float_as_int(e.y);
      const int pos = __popc(__ballot_sync(SRB_FULL_MASK, lane < K && ls > cs));
      const float ps = __shfl_up_sync(SRB_FULL_MASK, ls, 1);
      const int pi = __shfl_up_sync(SRB_FULL_MASK, li, 1);
      if (lane > pos && lane < K) ls = ps, li = pi;
      if (lane == pos) ls = cs, li = cid;
      thr = __shfl_sync(SRB_FULL_MASK, ls, K - 1);
    }
  }
  // exactness test
  const float kth = __shfl_sync(SRB_FULL_MASK, ls, K - 1);
  const float thr32 = fmaxf(a.cand_thr[(size_t)q * 2], a.cand_thr[(size_t)q * 2 + 1]);
  int deg = 0;
  if (a.rated_ptr) {
    const int u = a.users[q];
    deg = a.rated_ptr[u + 1] - a.rated_ptr[u];
  }
  const bool unsafe = (a.n_items - deg < K) || !(thr32 + E < kth);
  if (unsafe) {
    if (lane == 0) {
      const int slot = atomicAdd(a.fb_count, 1);
      a.fb_rows[slot] = q;
      a.fb_users[slot] = a.users[q];
    }
    return;
  }
  if (lane < K) {
    a.out_ids[(size_t)q * K + lane] = li;
    a.out_scores[(size_t)q * K + lane] = ls;
  }
}

static int64_t tc_align(int64_t x) { return (x + 255) / 256 * 256; }

struct TcWorkspace {
  float* ug;
  float* unorm;
  unsigned int* bmax;
  float* cand_s;
  int32_t* cand_i;
  int32_t* cand_n;
  float* cand_thr;
  int32_t* fb_count;
  int32_t* fb_rows;
  int32_t* fb_users;
  float* fb_scratch;  // [fb_cap][n_items] exact score rows of the users re-run by the fast fallback
  int32_t fb_cap;
  int64_t bytes;
};

static int tc_fb_cap(int n_items) {
  long long cap = (64ll << 20) / ((long long)n_items * 4);  // at most 64 MB of scratch
  if (cap > 256) cap = 256;
  if (cap < 8) cap = 8;
  return (int)cap;
}

static TcWorkspace tc_carve(char* base, int n_q, int n_items) {
  TcWorkspace w;
  const int64_t n_q_pad = ((int64_t)n_q + 255) / 256 * 256 + 256;
  int64_t off = 0;
  auto take = [&](int64_t b) {
    char* p = base ? base + off : nullptr;
    off += tc_align(b);
    return p;
  };
  w.ug = (float*)take(n_q_pad * TC_D * 4);
  w.unorm = (float*)take(n_q_pad * 4);
  w.bmax = (unsigned int*)take(16);
  w.cand_s = (float*)take((int64_t)n_q * TC_CAND * 4);
  w.cand_i = (int32_t*)take((int64_t)n_q * TC_CAND * 4);
  w.cand_n = (int32_t*)take((int64_t)n_q * 2 * 4);
  w.cand_thr = (float*)take((int64_t)n_q * 2 * 4);
  w.fb_count = (int32_t*)take(16);
  w.fb_rows = (int32_t*)take((int64_t)n_q * 4);
  w.fb_users = (int32_t*)take((int64_t)n_q * 4);
  w.fb_cap = tc_fb_cap(n_items);
  w.fb_scratch = (float*)take((int64_t)w.fb_cap * n_items * 4);
  w.bytes = off;
  return w;
}

int score_topk_fallback(const srb_topk_desc* d, const int32_t* fb_users, const int32_t* fb_rows, const int32_t* fb_count,
                        float* scratch, int fb_cap, cudaStream_t st);  // score_topk.cu

int score_topk_tc(const srb_topk_desc* d, cudaStream_t st) {
  SRB_REQUIRE(d->d == TC_D, "topk impl 2 (tcgen05) supports d=64 only (got %d)", d->d);
  const int n_q = d->n_q;
  const TcWorkspace need = tc_carve(nullptr, n_q, d->n_items);
  SRB_REQUIRE(d->workspace && d->workspace_bytes >= need.bytes, "topk impl 2: workspace too small (%lld < %lld)",
              (long long)d->workspace_bytes, (long long)need.bytes);
  SRB_REQUIRE(((uintptr_t)d->workspace & 255) == 0, "topk impl 2: workspace must be 256-byte aligned");
  SRB_REQUIRE(((uintptr_t)d->item_emb & 15) == 0, "topk impl 2: item_emb must be 16-byte aligned");
  TcWorkspace w = tc_carve((char*)d->workspace, n_q, d->n_items);
  const int n_q_pad = (n_q + 255) / 256 * 256 + 256;
  SRB_TRY(check_cuda(cudaMemsetAsync(w.bmax, 0, 16, st), "tc memset"));
  SRB_TRY(check_cuda(cudaMemsetAsync(w.fb_count, 0, 16, st), "tc memset"));
  {
    const long long rows = (long long)n_q_pad + d->n_items;
    tc_gather_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(d->user_emb, d->users, n_q, n_q_pad, w.ug, w.unorm, d->item_emb,
                                                                d->n_items, w.bmax);
    SRB_TRY(post_launch("tc_gather_kernel"));
  }
  // users per CTA: spread the queries over one wave of SMs, multiple of 32, at most 256
  const int sms = sm_count();
  int ub = (n_q + sms - 1) / sms;
  ub = (ub + 31) / 32 * 32;
  if (ub > 256) ub = 256;
  if (ub < 32) ub = 32;
  const int blocks = (n_q + ub - 1) / ub;
  CUtensorMap tm_users, tm_items;
  SRB_REQUIRE(make_tmap_f32_rows(&tm_users, w.ug, (uint64_t)n_q_pad, TC_D, 128) == 0, "topk impl 2: cuTensorMapEncodeTiled(users) failed");
  SRB_REQUIRE(make_tmap_f32_rows(&tm_items, d->item_emb, (uint64_t)d->n_items, TC_D, TC_TN) == 0,
              "topk impl 2: cuTensorMapEncodeTiled(items) failed");
  TcArgs a;
  a.users = d->users;
  a.rated_ptr = d->rated_ptr;
  a.rated_idx = d->rated_idx;
  a.n_q = n_q;
  a.n_items = d->n_items;
  a.ub = ub;
  a.cand_s = w.cand_s;
  a.cand_i = w.cand_i;
  a.cand_n = w.cand_n;
  a.cand_thr = w.cand_thr;
  const size_t smem = TcSmem::total + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    SRB_TRY(check_cuda(cudaFuncSetAttribute(tc_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tc smem attr"));
    attr_done = true;
  }
  tc_score_kernel<<<blocks, TC_THREADS, smem, st>>>(tm_users, tm_items, a);
  SRB_TRY(post_launch("tc_score_kernel"));
  RescoreArgs r;
  r.ug = w.ug;
  r.item_emb = d->item_emb;
  r.unorm = w.unorm;
  r.bmax_bits = w.bmax;
  r.users = d->users;
  r.rated_ptr = d->rated_ptr;
  r.cand_s = w.cand_s;
  r.cand_i = w.cand_i;
  r.cand_n = w.cand_n;
  r.cand_thr = w.cand_thr;
  r.n_q = n_q;
  r.n_items = d->n_items;
  r.k = d->k;
  r.out_ids = d->out_ids;
  r.out_scores = d->out_scores;
  r.fb_count = w.fb_count;
  r.fb_rows = w.fb_rows;
  r.fb_users = w.fb_users;
  tc_rescore_kernel<<<(n_q + 7) / 8, 256, 0, st>>>(r);
  SRB_TRY(post_launch("tc_rescore_kernel"));
  return score_topk_fallback(d, w.fb_users, w.fb_rows, w.fb_count, w.fb_scratch, w.fb_cap, st);
}

}  // namespace srb

// byte offset of the int32 fallback counter inside the workspace (diagnostics: how many users the
// exact kernel had to re-run)
extern "C" int64_t srb_topk_fallback_count_offset(int32_t n_q, int32_t n_items) {
  if (n_q <= 0 || n_items <= 0) return -1;
  const srb::TcWorkspace w = srb::tc_carve((char*)256, n_q, n_items);
  return (int64_t)((char*)w.fb_count - (char*)256);
}

extern "C" int64_t srb_topk_workspace_bytes(int32_t n_q, int32_t n_items, int32_t d, int32_t k) {
  (void)d;
  (void)k;
  if (n_q <= 0 || n_items <= 0) return 0;
  return srb::tc_carve(nullptr, n_q, n_items).bytes;
}
