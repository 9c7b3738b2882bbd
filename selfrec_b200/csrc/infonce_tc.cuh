// (iii) InfoNCE forward/backward on the 5th-gen tensor cores (tcgen05, TF32), d = 64.
//
// Same contract as infonce.cu (util/loss_torch.py:35-50 + autograd backward); the n x n logit matrix
// lives only in TMEM / shared memory.  Per problem, with V1, V2 the L2-normalised gathered views:
//   pass LSE     S = V1 V2^T / tau, tile by tile;  per-row running (max, sum-exp) partials
//   pass GRAD-A  rows = view-1 rows i:   G = exp(S - lse_i) * w/(n tau), j != i;   dV1 += G  V2
//   pass GRAD-B  rows = view-2 rows j:   G'= exp(S^T - lse_i) * w/(n tau), i != j;  dV2 += G' V1
//   finish       adds the diagonal term (P_ii - 1) w/(n tau) v_i in exact fp32, then the normalisation backward
// Each CTA owns a block of 128 rows and a strided subset of the 128-column tiles:
//   warp 0    TMA producer: column-operand tile [128 x 64] (K-major, for S) and its transposed copy
//             [64 x 128] (K-major over the column index, for G V), 128-byte swizzle, 2 stages
//   warp 1    MMA issuer:   S tile  = tcgen05.mma kind::tf32  M=128 N=128 K=64   (8 k-steps)
//                           D[128 x 64] += G[128 x 128] Vcol[128 x 64]           (16 k-steps, N=64)
//   warps 2-5 epilogue:     tcgen05.ld S row -> G in registers -> shared memory in the UMMA K-major
//                           128B-swizzle layout (the layout TMA would have produced), fence.proxy.async,
//                           hand the tile to the MMA warp; finally D -> global (atomicAdd over splits)
// fp32 accuracy on a TF32 pipe: every operand is split x = hi + lo (both TF32, round-to-nearest) and each
// product is evaluated as hi*hi + hi*lo + lo*hi ("3xTF32", error ~2^-22 instead of 2^-11): V1, V2 and their
// transposes arrive as hi/lo pairs, G is split by the epilogue.  S_ii and the normalisation backward use the
// exact fp32 rows.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace srb {

using namespace tc;

constexpr int NT_D = 64;
constexpr int NT_T = 128;        // rows per CTA (UMMA M)
constexpr int NT_C = 64;         // columns per tile (UMMA N of the S product, K of the G V product)
constexpr int NT_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int NT_MAX_SPLITS = 8;
constexpr uint32_t NT_ATILE = NT_T * NT_D * 4;   // 32 KB: 2 k-chunks x [128][32]      (x2: hi, lo)
constexpr uint32_t NT_BTILE = NT_C * NT_D * 4;   // 16 KB: 2 k-chunks x [64][32]       (x2: hi, lo)
constexpr uint32_t NT_TTILE = NT_D * NT_C * 4;   // 16 KB: 2 column chunks x [64][32]  (x2: hi, lo)
constexpr uint32_t NT_GTILE = NT_T * NT_C * 4;   // 32 KB: 2 column chunks x [128][32] (x2: hi, lo)

struct NtSmem {
  static constexpr uint32_t a_off = 0;                              // row-operand tile hi|lo (fixed per CTA)  64 KB
  static constexpr uint32_t b_off = a_off + 2 * NT_ATILE;           // 2 stages x (hi|lo) column tile          64 KB
  static constexpr uint32_t bt_off = b_off + 2 * 2 * NT_BTILE;      // 1 stage  x (hi|lo) transposed tile      32 KB
  static constexpr uint32_t g_off = bt_off + 2 * NT_TTILE;          // G tile hi|lo                            64 KB
  static constexpr uint32_t bar_off = g_off + 2 * NT_GTILE;
  static constexpr uint32_t total = bar_off + 1024;                 // barriers (128 B) + lse of the current column tile
};

struct NtProblem {
  int32_t n;
  const int32_t* n_dev;
  float weight;
  const float* diag;     // [NP] exact S_ii
  float* part_m;         // [SPLITS][NP]
  float* part_l;
  float* lse;            // [NP] combined log-sum-exp of every view-1 row (written by GRAD-A, read by finish)
  float* dV1;            // [NP][64] accumulators (zeroed by prep)
  float* dV2;
  float* loss_acc;
};

struct NtArgs {
  int32_t np;      // padded capacity, multiple of 128
  int32_t splits;
  float inv_tau;
  NtProblem p[2];
};

struct NtMaps {
  // per problem (<= 2), hi and lo parts [2]: row-major views as row operand (box [128][32]) and as column
  // operand (box [64][32]), and the transposed views ([64 rows][NP], box [64][32])
  CUtensorMap v1r[2][2], v2r[2][2], v1c[2][2], v2c[2][2], v1t[2][2], v2t[2][2];
};

__device__ __forceinline__ int nt_n(const NtProblem& p) { return p.n_dev ? min(*p.n_dev, p.n) : p.n; }

__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// mode 0: LSE (rows = view 1)   mode 1: GRAD-A (rows = view 1)   mode 2: GRAD-B (rows = view 2)
template <int MODE>
__global__ void __launch_bounds__(NT_THREADS, 1) nce_tc_kernel(const __grid_constant__ NtMaps maps, const NtArgs a) {
  extern __shared__ __align__(1024) uint8_t nt_smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(nt_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + NtSmem::bar_off);
  uint64_t* bar_full = bars;          // [2] TMA -> MMA (column tiles)
  uint64_t* bar_empty = bars + 2;     // [2] MMA2 done -> TMA
  uint64_t* bar_sfull = bars + 4;     // [2] S accumulator ready -> epilogue
  uint64_t* bar_sempty = bars + 6;    // [2] epilogue drained S -> MMA
  uint64_t* bar_gfull = bars + 8;     // [1] G tile written -> MMA
  uint64_t* bar_gempty = bars + 9;    // [1] MMA2 consumed G -> epilogue
  uint64_t* bar_a = bars + 10;        // [1] row tile loaded
  uint64_t* bar_dfull = bars + 11;    // [1] final D accumulator ready
  uint64_t* bar_tfull = bars + 13;    // [1] transposed column tile loaded
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* lse_col = reinterpret_cast<float*>(bars + 16);  // [128] lse of the current column tile (GRAD-B)

  const int prob = blockIdx.z;
  const NtProblem& P = a.p[prob];
  const int n = nt_n(P);
  const int r0 = blockIdx.x * NT_T;
  if (r0 >= n) return;
  const int split = blockIdx.y;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (n + NT_C - 1) / NT_C;
  const int my_tiles = (n_tiles - split + a.splits - 1) / a.splits;  // tiles split, split+S, ...
  if (my_tiles <= 0) {
    if (MODE == 0 && threadIdx.x < NT_T) {  // an empty split still has to publish neutral partials
      P.part_m[(size_t)split * a.np + r0 + threadIdx.x] = -INFINITY;
      P.part_l[(size_t)split * a.np + r0 + threadIdx.x] = 0.f;
    }
    return;
  }
  const CUtensorMap* map_row = (MODE == 2) ? maps.v2r[prob] : maps.v1r[prob];    // [2]: hi, lo
  const CUtensorMap* map_col = (MODE == 2) ? maps.v1c[prob] : maps.v2c[prob];
  const CUtensorMap* map_colt = (MODE == 2) ? maps.v1t[prob] : maps.v2t[prob];

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
      mbar_init(bar_sfull + s, 1);
      mbar_init(bar_sempty + s, 4);
    }
    mbar_init(bar_gfull, 4);
    mbar_init(bar_gempty, 1);
    mbar_init(bar_a, 1);
    mbar_init(bar_dfull, 1);
    mbar_init(bar_tfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem + 128;  // columns [128, 192): the G V accumulator; S stages at [0,64), [64,128)

  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar_a, 2 * NT_ATILE);
      for (int h = 0; h < 2; ++h)
        for (int c = 0; c < 2; ++c) tma_load_2d(sm + NtSmem::a_off + h * NT_ATILE + c * 16384, &map_row[h], bar_a, c * 32, r0);
      for (int k = 0; k < my_tiles; ++k) {
        const int t = split + k * a.splits;
        const int s = k & 1;
        mbar_wait(bar_empty + s, ((k >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(bar_full + s, 2 * NT_BTILE);
        for (int h = 0; h < 2; ++h)
          for (int c = 0; c < 2; ++c)
            tma_load_2d(sm + NtSmem::b_off + (s * 2 + h) * NT_BTILE + c * 8192, &map_col[h], bar_full + s, c * 32, t * NT_C);
        if (MODE != 0) {  // single-stage transposed tile: reusable once the previous G V product has read it
          mbar_wait(bar_gempty, (k & 1) ^ 1);
          mbar_arrive_expect_tx(bar_tfull, 2 * NT_TTILE);
          for (int h = 0; h < 2; ++h)
            for (int c = 0; c < 2; ++c)
              tma_load_2d(sm + NtSmem::bt_off + h * NT_TTILE + c * 8192, &map_colt[h], bar_tfull, t * NT_C + c * 32, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, NT_C);  // both products: M = 128, N = 64
      const uint32_t a_base = smem_u32(sm + NtSmem::a_off);
      const uint32_t g_base = smem_u32(sm + NtSmem::g_off);
      const uint32_t bt_base = smem_u32(sm + NtSmem::bt_off);
      mbar_wait(bar_a, 0);
      auto issue_s = [&](int k) {
        const int s = k & 1;
        mbar_wait(bar_sempty + s, ((k >> 1) & 1) ^ 1);
        mbar_wait(bar_full + s, (k >> 1) & 1);
        fence_after_sync();
        const uint32_t b_base = smem_u32(sm + NtSmem::b_off + s * 2 * NT_BTILE);
        // S = Ahi Bhi + Ahi Blo + Alo Bhi
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const uint32_t ab = a_base + (pr == 2 ? NT_ATILE : 0);
          const uint32_t bb = b_base + (pr == 1 ? NT_BTILE : 0);
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_tf32_ss(tmem + s * 64, make_smem_desc_k_sw128(ab + c * 16384 + kk * 32),
                           make_smem_desc_k_sw128(bb + c * 8192 + kk * 32), idesc, (pr | c | kk) ? 1u : 0u);
        }
        umma_commit(bar_sfull + s);
        umma_commit(bar_empty + s);  // the column tile (K-major copy) is free once S is computed
      };
      issue_s(0);
      for (int k = 0; k < my_tiles; ++k) {
        if (k + 1 < my_tiles) issue_s(k + 1);  // S of the next tile overlaps the epilogue of this one
        if (MODE != 0) {
          mbar_wait(bar_gfull, k & 1);
          mbar_wait(bar_tfull, k & 1);
          fence_after_sync();
          // D += Ghi Vhi + Glo Vhi + Ghi Vlo      (K = 64 columns of this tile)
#pragma unroll
          for (int pr = 0; pr < 3; ++pr) {
            const uint32_t gb = g_base + (pr == 1 ? NT_GTILE : 0);
            const uint32_t vb = bt_base + (pr == 2 ? NT_TTILE : 0);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_tf32_ss(tmem_d, make_smem_desc_k_sw128(gb + c * 16384 + kk * 32),
                             make_smem_desc_k_sw128(vb + c * 8192 + kk * 32), idesc, (k | pr | c | kk) ? 1u : 0u);
          }
          umma_commit(bar_gempty);  // G buffer and the transposed tile are reusable
        }
      }
      if (MODE != 0) umma_commit(bar_dfull);
    }
  } else {
    // ===== epilogue warps: thread = one row of the block =====
    const int quarter = warp & 3;
    const int row_l = quarter * 32 + lane;
    const int row = r0 + row_l;
    const bool row_ok = row < n;
    const float L2E = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;  // LSE mode
    float lse_row = 0.f;                   // GRAD-A: lse of my row
    const float gscale = P.weight * a.inv_tau / (float)n;
    if (MODE == 1) {
      float M = -INFINITY;
      for (int s = 0; s < a.splits; ++s) M = fmaxf(M, P.part_m[(size_t)s * a.np + row]);
      float Ls = 0.f;
      for (int s = 0; s < a.splits; ++s) {
        const float ms = P.part_m[(size_t)s * a.np + row];
        if (ms > -INFINITY) Ls += P.part_l[(size_t)s * a.np + row] * exp2f((ms - M) * L2E);
      }
      lse_row = row_ok ? M + logf(Ls) : 0.f;
      if (split == 0) {  // loss = mean(lse_i - S_ii), S_ii exact
        P.lse[row] = lse_row;
        float contrib = row_ok ? lse_row - P.diag[row] : 0.f;
        contrib = warp_sum(contrib);
        if (lane == 0) atomicAdd(P.loss_acc, contrib);
      }
    }
    uint8_t* gsm = sm + NtSmem::g_off;
    for (int k = 0; k < my_tiles; ++k) {
      const int t = split + k * a.splits;
      const int s = k & 1;
      const int c0 = t * NT_C;
      if (MODE == 2) {
        // lse of the 128 columns of this tile (view-1 rows), combined from the partials
        asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with the previous tile's values
        for (int cc = threadIdx.x - 64; cc < NT_C; cc += 128) {
          const int col = c0 + cc;
          float M = -INFINITY;
          for (int sp = 0; sp < a.splits; ++sp) M = fmaxf(M, P.part_m[(size_t)sp * a.np + col]);
          float Ls = 0.f;
          for (int sp = 0; sp < a.splits; ++sp) {
            const float ms = P.part_m[(size_t)sp * a.np + col];
            if (ms > -INFINITY) Ls += P.part_l[(size_t)sp * a.np + col] * exp2f((ms - M) * L2E);
          }
          lse_col[cc] = (col < n) ? M + logf(Ls) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the 4 epilogue warps
      }
      mbar_wait(bar_sfull + s, (k >> 1) & 1);
      fence_after_sync();
      if (MODE != 0) mbar_wait(bar_gempty, (k & 1) ^ 1);  // previous G consumed by the MMA
#pragma unroll 1
      for (int g = 0; g < NT_C / 32; ++g) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + s * 64 + g * 32, r);
        tmem_ld_wait();
        if (MODE == 0) {
          float tm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = (c0 + g * 32 + j < n) ? __uint_as_float(r[j]) * a.inv_tau : -INFINITY;
            r[j] = __float_as_uint(v);
            tm = fmaxf(tm, v);
          }
          const float mn = fmaxf(m_run, tm);
          if (mn > -INFINITY) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += exp2f((__uint_as_float(r[j]) - mn) * L2E);
            l_run = l_run * exp2f((m_run - mn) * L2E) + acc;
            m_run = mn;
          }
        } else {
          // G row chunk -> shared memory, UMMA K-major 128B-swizzle layout:
          //   chunk g (32 columns) at g*16 KB, row r at r*128 B, 16-byte unit u stored at u ^ (r & 7)
          uint8_t* dst = gsm + g * 16384 + row_l * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float gv[4], gl[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int j = u * 4 + q;
              const int col = c0 + g * 32 + j;
              const float lse = (MODE == 1) ? lse_row : lse_col[g * 32 + j];
              float x = 0.f;
              // the diagonal term (P_ii - 1) v_i is added in exact fp32 by the finish kernel: through the
              // tensor core its TF32 rounding (|G_ii| ~ 1) would dominate the error of the whole row
              if (row_ok && col < n && col != row) x = exp2f((__uint_as_float(r[j]) * a.inv_tau - lse) * L2E) * gscale;
              gv[q] = to_tf32_rna(x);
              gl[q] = to_tf32_rna(x - gv[q]);
            }
            *reinterpret_cast<float4*>(dst + ((u ^ (row_l & 7)) << 4)) = make_float4(gv[0], gv[1], gv[2], gv[3]);
            *reinterpret_cast<float4*>(dst + NT_GTILE + ((u ^ (row_l & 7)) << 4)) = make_float4(gl[0], gl[1], gl[2], gl[3]);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + s);
      if (MODE != 0) {
        fence_proxy_async();  // generic-proxy stores of G -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gfull);
      }
    }
    if (MODE == 0) {
      if (row < a.np) {
        P.part_m[(size_t)split * a.np + row] = m_run;
        P.part_l[(size_t)split * a.np + row] = l_run;
      }
    } else {
      mbar_wait(bar_dfull, 0);
      fence_after_sync();
      float* out = (MODE == 1) ? P.dV1 : P.dV2;
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_d + ((uint32_t)(quarter * 32) << 16) + g * 32, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(out + (size_t)row * NT_D + g * 32 + j, __uint_as_float(r[j]));
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace srb
