// Device-side assembly of the normalised bipartite adjacency  D^-1/2 [[0, R], [R^T, 0]] D^-1/2  as CSR.
//
// Replaces, for a set of kept interaction edges,
//   Interaction.__create_sparse_bipartite_adjacency / convert_to_laplacian_mat   data/ui_graph.py:47-65
//   Graph.normalize_graph_mat                                                   data/graph.py:10-24
// i.e. what SGL rebuilds twice per epoch after GraphAugmentor.edge_dropout / node_dropout (SGL.py:27-29,
// 80-96, data/augmentor.py:11-40), and what a config-5 sized graph (10 M x 2 M x 200 M) cannot afford to
// build with scipy on the host.
//
// Everything is edge-parallel (a power-law graph has rows with millions of entries; nothing here walks a
// row with one thread or one warp):
//   1. pack (kept?, weight) per edge in the users-x-items order and in the items-x-users order;
//   2. exclusive scan of both (64-bit: count in the high word, weight sum in the low word);
//   3. per row: count and row sum from the scanned values at the row boundaries,
//      d = table[rowsum]  (table = numpy's float32 power(k, -0.5), inf -> 0, computed by the host so that the
//      rounding is the reference's);
//   4. scan of the counts -> rowptr;
//   5. per kept edge: position = rowptr[row] + (edges kept before it in its row) -- columns stay ascending --
//      and value (d[r] * w) * d[c], the two fp32 products of graph.py:16-18 in the reference's order.
// Integer work + two fp32 multiplies per entry: HBM-bound streaming, no tensor cores.
#include "common.cuh"

namespace srb {

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------------
// exclusive scan of u64 values, in place, n elements (three passes; SCAN_TILE elements per block)
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_PER = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_PER;

__device__ __forceinline__ u64 block_exclusive_scan(u64 v, u64* total, u64* sh) {
  // sh: [SCAN_THREADS / 32 + 1]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  u64 x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u64 y = __shfl_up_sync(SRB_FULL_MASK, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) sh[wib] = x;
  __syncthreads();
  if (wib == 0) {
    u64 w = lane < SCAN_THREADS / 32 ? sh[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u64 y = __shfl_up_sync(SRB_FULL_MASK, w, o);
      if (lane >= o) w += y;
    }
    if (lane < SCAN_THREADS / 32) sh[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const u64 warp_off = wib ? sh[wib - 1] : 0;
  if (total) *total = sh[SCAN_THREADS / 32 - 1];
  return warp_off + x - v;  // exclusive
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const u64* x, long long n, u64* sums) {
  __shared__ u64 sh[SCAN_THREADS / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) {
    const long long i = base + (long long)k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += x[i];
  }
  u64 tot;
  block_exclusive_scan(s, &tot, sh);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// one block scans the tile sums in place (n_tiles is at most a few 100 k)
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(u64* sums, int n_tiles) {
  __shared__ u64 sh[SCAN_THREADS / 32 + 1];
  __shared__ u64 carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
    const int i = base + threadIdx.x;
    const u64 v = i < n_tiles ? sums[i] : 0;
    u64 tot;
    const u64 ex = block_exclusive_scan(v, &tot, sh);
    const u64 carry = carry_s;
    if (i < n_tiles) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(u64* x, long long n, const u64* sums) {
  __shared__ u64 sh[SCAN_THREADS / 32 + 1];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER;  // blocked: thread owns a run
  u64 v[SCAN_PER];
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) {
    v[k] = (base + k < n) ? x[base + k] : 0;
    s += v[k];
  }
  u64 run = sums[blockIdx.x] + block_exclusive_scan(s, nullptr, sh);
#pragma unroll
  for (int k = 0; k < SCAN_PER; ++k) {
    if (base + k < n) x[base + k] = run;
    run += v[k];
  }
}

static int exclusive_scan_u64(u64* x, long long n, u64* sums, cudaStream_t st) {
  if (n <= 0) return SRB_OK;
  const int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, st>>>(x, n, sums);
  SRB_TRY(post_launch("scan_tile_sums_kernel"));
  scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(sums, tiles);
  SRB_TRY(post_launch("scan_sums_kernel"));
  scan_apply_kernel<<<tiles, SCAN_THREADS, 0, st>>>(x, n, sums);
  return post_launch("scan_apply_kernel");
}

// ---------------------------------------------------------------------------------------
struct GbArgs {
  const int32_t* ui_ptr;
  const int32_t* ui_col;
  const float* ui_val;  // optional multiplicities
  const int32_t* iu_ptr;
  const int32_t* iu_col;
  const int32_t* iu_perm;
  const uint8_t* keep;  // optional flags over the ui order
  int reset;            // kept edges get weight 1 (augmentor.py:36)
  int n_users, n_items;
  long long nnz;
  const float* dinv_table;
  int table_n;
  u64* s_ui;   // [nnz + 1]
  u64* s_iu;   // [nnz + 1]
  u64* s_row;  // [N + 1]
  float* dinv; // [N]
  int32_t* rowptr;
  int32_t* colidx;
  float* vals;
  long long cap;
  long long* nnz_out;
};

__device__ __forceinline__ u64 pack_edge(const GbArgs& a, long long p) {
  if (a.keep && !a.keep[p]) return 0ull;
  const unsigned w = (a.reset || !a.ui_val) ? 1u : (unsigned)a.ui_val[p];
  return (1ull << 32) | w;
}

__global__ void __launch_bounds__(256) gb_pack_kernel(const GbArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p <= a.nnz; p += stride) {
    a.s_ui[p] = p < a.nnz ? pack_edge(a, p) : 0ull;
    a.s_iu[p] = p < a.nnz ? pack_edge(a, a.iu_perm[p]) : 0ull;
  }
}

__global__ void __launch_bounds__(256) gb_rows_kernel(const GbArgs a) {
  const int n = a.n_users + a.n_items;
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += stride) {
    u64 c = 0;
    if (r < n) {
      const bool user = r < a.n_users;
      const int32_t* ptr = user ? a.ui_ptr : a.iu_ptr;
      const u64* s = user ? a.s_ui : a.s_iu;
      const int k = user ? r : r - a.n_users;
      const u64 d = s[ptr[k + 1]] - s[ptr[k]];  // no borrow across the halves: both fields are monotone
      c = d >> 32;
      const unsigned w = (unsigned)(d & 0xffffffffull);
      a.dinv[r] = a.dinv_table[w < (unsigned)a.table_n ? w : a.table_n - 1];
    }
    a.s_row[r] = c;
  }
}

__global__ void __launch_bounds__(256) gb_rowptr_kernel(const GbArgs a) {
  const int n = a.n_users + a.n_items;
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n; r += stride) {
    a.rowptr[r] = (int32_t)a.s_row[r];
    if (r == n && a.nnz_out) *a.nnz_out = (long long)a.s_row[r];
  }
}

// row of entry p in a CSR row-pointer array (largest r with ptr[r] <= p)
__device__ __forceinline__ int row_of(const int32_t* ptr, int n_rows, long long p) {
  int lo = 0, hi = n_rows;  // invariant: ptr[lo] <= p < ptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)__ldg(ptr + mid) <= p) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) gb_fill_kernel(const GbArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * a.nnz; t += stride) {
    const bool user_half = t < a.nnz;
    const long long p = user_half ? t : t - a.nnz;       // position in this half's own order
    const long long pe = user_half ? p : a.iu_perm[p];   // the edge, in the ui order
    if (a.keep && !a.keep[pe]) continue;
    int r, c;
    long long before;
    if (user_half) {
      const int u = row_of(a.ui_ptr, a.n_users, p);
      r = u;
      c = a.n_users + a.ui_col[p];
      before = (long long)((a.s_ui[p] - a.s_ui[a.ui_ptr[u]]) >> 32);
    } else {
      const int i = row_of(a.iu_ptr, a.n_items, p);
      r = a.n_users + i;
      c = a.iu_col[p];
      before = (long long)((a.s_iu[p] - a.s_iu[a.iu_ptr[i]]) >> 32);
    }
    const long long pos = (long long)a.rowptr[r] + before;
    if (pos >= a.cap) continue;  // (srb_graph_assemble checks the capacity up front)
    const float w = (a.reset || !a.ui_val) ? 1.0f : a.ui_val[pe];
    a.colidx[pos] = c;
    a.vals[pos] = __fmul_rn(__fmul_rn(a.dinv[r], w), a.dinv[c]);  // (d_r * a) * d_c, graph.py:16-18
  }
}

__global__ void __launch_bounds__(256) gb_flags_kernel(const long long* idx, long long n, uint8_t* flags, long long nnz) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const long long p = idx[t];
    if (p >= 0 && p < nnz) flags[p] = 1;
  }
}

static int64_t gb_align(int64_t x) { return (x + 255) / 256 * 256; }

}  // namespace srb

extern "C" int64_t srb_graph_assemble_workspace_bytes(int32_t n_users, int32_t n_items, int64_t nnz) {
  using namespace srb;
  const int64_t n = (int64_t)n_users + n_items;
  const int64_t tiles = (nnz + 1 + SCAN_TILE - 1) / SCAN_TILE + (n + 1 + SCAN_TILE - 1) / SCAN_TILE + 8;
  return 2 * gb_align((nnz + 1) * 8) + gb_align((n + 1) * 8) + gb_align(n * 4) + gb_align(tiles * 8) + gb_align(nnz) + 1024;
}

extern "C" int srb_graph_assemble(const srb_graph_assemble_desc* g, void* stream) {
  using namespace srb;
  SRB_REQUIRE(g != nullptr, "graph_assemble: null desc");
  SRB_REQUIRE(g->n_users > 0 && g->n_items > 0 && g->nnz >= 0, "graph_assemble: bad shape");
  SRB_REQUIRE(g->nnz < (1ll << 30), "graph_assemble: 2*nnz must fit in int32");
  SRB_REQUIRE(g->ui_ptr && g->ui_col && g->iu_ptr && g->iu_col && g->iu_perm, "graph_assemble: null interaction CSR");
  SRB_REQUIRE(g->dinv_table && g->dinv_table_n > 0, "graph_assemble: degree table missing");
  SRB_REQUIRE(g->rowptr && g->colidx && g->vals && g->dinv, "graph_assemble: null output");
  SRB_REQUIRE(!(g->keep_flags && g->keep_idx), "graph_assemble: give keep_flags or keep_idx, not both");
  const int64_t kept = g->keep_idx ? g->n_keep : g->nnz;
  SRB_REQUIRE(!g->keep_idx || (g->n_keep >= 0 && g->n_keep <= g->nnz), "graph_assemble: bad n_keep");
  SRB_REQUIRE(g->out_cap >= 2 * kept || g->keep_flags, "graph_assemble: output capacity %lld < %lld", (long long)g->out_cap, (long long)(2 * kept));
  const int64_t need = srb_graph_assemble_workspace_bytes(g->n_users, g->n_items, g->nnz);
  SRB_REQUIRE(g->workspace && g->workspace_bytes >= need, "graph_assemble: workspace too small (%lld < %lld)",
              (long long)g->workspace_bytes, (long long)need);
  SRB_REQUIRE(((uintptr_t)g->workspace & 255) == 0, "graph_assemble: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)g->n_users + g->n_items;
  char* w = (char*)g->workspace;
  GbArgs a = {};
  a.s_ui = (u64*)w;
  w += gb_align((g->nnz + 1) * 8);
  a.s_iu = (u64*)w;
  w += gb_align((g->nnz + 1) * 8);
  a.s_row = (u64*)w;
  w += gb_align((n + 1) * 8);
  w += gb_align(n * 4);  // (reserved)
  u64* sums = (u64*)w;
  w += gb_align(((g->nnz + 1 + SCAN_TILE - 1) / SCAN_TILE + (n + 1 + SCAN_TILE - 1) / SCAN_TILE + 8) * 8);
  uint8_t* flags = (uint8_t*)w;
  a.ui_ptr = g->ui_ptr;
  a.ui_col = g->ui_col;
  a.ui_val = g->ui_val;
  a.iu_ptr = g->iu_ptr;
  a.iu_col = g->iu_col;
  a.iu_perm = g->iu_perm;
  a.keep = g->keep_flags;
  a.reset = g->reset_weights;
  a.n_users = g->n_users;
  a.n_items = g->n_items;
  a.nnz = g->nnz;
  a.dinv_table = g->dinv_table;
  a.table_n = g->dinv_table_n;
  a.dinv = g->dinv;
  a.rowptr = g->rowptr;
  a.colidx = g->colidx;
  a.vals = g->vals;
  a.cap = g->out_cap;
  a.nnz_out = (long long*)g->nnz_out;
  const int grid = sm_count() * 8;
  if (g->keep_idx) {
    SRB_TRY(check_cuda(cudaMemsetAsync(flags, 0, (size_t)g->nnz, st), "graph_assemble memset"));
    gb_flags_kernel<<<grid, 256, 0, st>>>((const long long*)g->keep_idx, g->n_keep, flags, g->nnz);
    SRB_TRY(post_launch("gb_flags_kernel"));
    a.keep = flags;
  }
  gb_pack_kernel<<<grid, 256, 0, st>>>(a);
  SRB_TRY(post_launch("gb_pack_kernel"));
  SRB_TRY(exclusive_scan_u64(a.s_ui, g->nnz + 1, sums, st));
  SRB_TRY(exclusive_scan_u64(a.s_iu, g->nnz + 1, sums, st));
  gb_rows_kernel<<<grid, 256, 0, st>>>(a);
  SRB_TRY(post_launch("gb_rows_kernel"));
  SRB_TRY(exclusive_scan_u64(a.s_row, n + 1, sums, st));
  gb_rowptr_kernel<<<grid, 256, 0, st>>>(a);
  SRB_TRY(post_launch("gb_rowptr_kernel"));
  gb_fill_kernel<<<grid, 256, 0, st>>>(a);
  return post_launch("gb_fill_kernel");
}
