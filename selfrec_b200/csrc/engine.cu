// One whole training step as a single stream-ordered call (graph-capturable).
//
// Replaces the body of the batch loop of <Model>.train():
//   MF.py:17-25  LightGCN.py:21-29  SimGCL.py:25-36  XSimGCL.py:27-37  SGL.py:30-41
// i.e. encoder forward (R4) -> gather + BPR + L2 (R5-R7) -> InfoNCE (R8) -> autograd
// backward -> Adam (R10).
//
// Backward through the propagation uses the fact that every encoder is linear in E0 and the
// SimGCL/XSimGCL noise has zero gradient (sign() and the noise are constants):
//   final = c * sum_k A^k E0   =>   dE0 = c * sum_k A^k G      (A symmetric)
// evaluated by Horner's rule with L SpMMs and no saved activations; the row-sparse loss
// gradients G (<= 3B + 2B rows) are kept compact and re-scattered at every level instead of
// being materialised as dense [N, d] tensors.  SimGCL's three encoders share A, so their
// three backward chains collapse into one.  The last SpMM applies Adam in its epilogue.
#include <mutex>
#include "common.cuh"

namespace srb {

struct Ws {
  float* final_;  // [N,d] main encoder output
  float* cl;      // [N,d] XSimGCL CL view / SimGCL,SGL view-1 output
  float* v2;      // [N,d] SimGCL,SGL view-2 output
  float* work0;
  float* work1;
  float* acc0;
  float* acc1;
  float* gd;      // dense gradient accumulator / last-layer addend
  float* rsum;    // running layer sum of the encoder being evaluated (training forwards)
  float* g_emb;   // [3,B,d]
  float* g_l2;    // [3,B,d]
  float* g_nce;   // 4 x [2B,d]
  float* bpr_scratch;  // [8]
  float* bpr_losses;   // [2]
  float* nce_losses;   // [4]
  int32_t* idx_cat;    // [2B] SGL concatenated unique ids
  int32_t* n_cat;      // [1]
  int32_t* batch_rows; // [4][3B] distinct table rows of the batch (u, U+i, U+j) by class: split, CTA, warp, lane group
  int32_t* n_hub;      // [8] rows per class [0..3], chunks of the split rows [4]
  uint32_t* row_mask;  // [(N+31)/32] bitmap of the batch's table rows (the rows the gradient seed touches)
  int32_t* hub_first;  // [3B] first chunk slot of each split batch row
  int32_t* hub_work;   // [hub_cap][2]
  float* hub_part;     // [hub_cap, d]
  int32_t hub_cap;
  void* nce_ws;
  int64_t nce_ws_bytes;
};

static int64_t al(int64_t x) { return (x + 255) / 256 * 256; }

static int64_t carve(const srb_step_desc* s, Ws* w, char* base) {
  const int64_t N = (int64_t)s->n_users + s->n_items, d = s->d, B = s->batch_cap;
  const int64_t nd = al(N * d * 4);
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    char* p = base ? base + off : nullptr;
    off += al(bytes);
    return p;
  };
  const bool graph = s->model != SRB_MODEL_MF;
  const bool two_views = s->model == SRB_MODEL_SIMGCL || s->model == SRB_MODEL_SGL;
  const bool has_cl = s->model == SRB_MODEL_XSIMGCL || two_views;
  float* f_final = (float*)take(graph ? nd : 0);
  float* f_cl = (float*)take(has_cl ? nd : 0);
  float* f_v2 = (float*)take(two_views ? nd : 0);
  float* f_w0 = (float*)take(graph ? nd : 0);
  float* f_w1 = (float*)take(graph ? nd : 0);
  float* f_a0 = (float*)take(nd);
  float* f_a1 = (float*)take(graph ? nd : 0);
  float* f_gd = (float*)take(graph ? nd : 0);
  float* f_rsum = (float*)take(graph ? nd : 0);
  float* f_gemb = (float*)take(3 * B * d * 4);
  float* f_gl2 = (float*)take(3 * B * d * 4);
  float* f_gnce = (float*)take(has_cl ? 4 * 2 * B * d * 4 : 0);
  float* f_bs = (float*)take(8 * 4);
  float* f_bl = (float*)take(2 * 4);
  float* f_nl = (float*)take(4 * 4);
  int32_t* i_cat = (int32_t*)take(2 * B * 4);
  int32_t* i_ncat = (int32_t*)take(4);
  int32_t* i_brows = (int32_t*)take(4 * 3 * B * 4);
  // [class counters (8 words) | row bitmap]: one memset clears all of it
  int32_t* i_nhub = (int32_t*)take(graph ? 32 + ((N + 31) / 32) * 4 : 32);
  uint32_t* u_mask = (uint32_t*)(i_nhub + 8);
  const int64_t hub_cap = graph ? s->adj.hub.n_work : 0;  // distinct batch rows: never more chunks than the whole graph has
  int32_t* i_hfirst = (int32_t*)take(hub_cap ? 3 * B * 4 : 0);
  int32_t* i_hwork = (int32_t*)take(hub_cap * 2 * 4);
  float* f_hpart = (float*)take(hub_cap * d * 4);
  const int64_t nws = has_cl ? srb_infonce_workspace_bytes((int32_t)(2 * B), (int32_t)d, 2) : 0;
  void* v_nws = take(nws);
  if (w) {
    w->final_ = f_final;
    w->cl = f_cl;
    w->v2 = f_v2;
    w->work0 = f_w0;
    w->work1 = f_w1;
    w->acc0 = f_a0;
    w->acc1 = f_a1;
    w->gd = f_gd;
    w->rsum = f_rsum;
    w->g_emb = f_gemb;
    w->g_l2 = f_gl2;
    w->g_nce = f_gnce;
    w->bpr_scratch = f_bs;
    w->bpr_losses = f_bl;
    w->nce_losses = f_nl;
    w->idx_cat = i_cat;
    w->n_cat = i_ncat;
    w->batch_rows = i_brows;
    w->n_hub = i_nhub;
    w->row_mask = u_mask;
    w->hub_first = i_hfirst;
    w->hub_work = i_hwork;
    w->hub_part = f_hpart;
    w->hub_cap = (int32_t)hub_cap;
    w->nce_ws = v_nws;
    w->nce_ws_bytes = nws;
  }
  return off;
}

// SGL: one InfoNCE over cat(users, items) (SGL.py:120-125) needs one combined id list.
__global__ void build_cat_idx_kernel(const int32_t* batch, int cap, int n_users, int32_t* idx_cat, int32_t* n_cat) {
  const int nu = min(batch[1], cap), ni = min(batch[2], cap);
  const int32_t* uu = batch + SRB_BATCH_HEADER + 3 * cap;
  const int32_t* ui = uu + cap;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nu + ni; t += gridDim.x * blockDim.x)
    idx_cat[t] = (t < nu) ? uu[t] : n_users + ui[t - nu];
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_cat = nu + ni;
}

// rows of the [N, d] tables a batch touches: u, U + i, U + j, each listed ONCE (the bitmap de-duplicates: a hub user
// sits in a batch many times) and classified by degree on the fly for the last-layer SpMM -- split rows (their
// chunks go to hub_work), a CTA per long row, a warp per other row; the lane-group class stays empty -- into four
// segments of capacity 3*cap; counters[c] ends up as the size of class c, counters[4] as the number of chunks.
// counters[0..7] and row_mask are zeroed by the caller.
// With a row range [row_begin, row_begin + n_local) (row-sharded tables) only the rows of that range are listed,
// as LOCAL row ids of the rank's CSR slice; the bitmap always covers all batch rows (global ids).
__global__ void __launch_bounds__(256) build_batch_rows_kernel(const int32_t* batch, int cap, int n_users, const int32_t* rowptr,
                                                               int row_begin, int n_local, int32_t* rows, int32_t* counters,
                                                               uint32_t* row_mask, int32_t* hub_first, int32_t* hub_work,
                                                               int hub_cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = min(batch[0], cap);
  const int sec = t / cap, k = t % cap;
  if (sec >= 3 || k >= b) return;
  const int32_t* u = batch + SRB_BATCH_HEADER;
  const int grow = (sec == 0) ? u[k] : n_users + u[sec * cap + k];
  const uint32_t bit = 1u << (grow & 31);
  if (atomicOr(row_mask + (grow >> 5), bit) & bit) return;  // listed already
  const int row = grow - row_begin;
  if (row < 0 || row >= n_local) return;
  const int deg = rowptr[row + 1] - rowptr[row];
  // only ~3B rows: parallelism is scarce, so no row shares a warp and rows above 4 warp-iterations get a CTA
  const int cls = (hub_first && deg >= SRB_HUB_MIN_NNZ) ? 0 : (deg >= 128 ? 1 : 2);
  // warp-aggregated slot allocation per class
  const unsigned mine = __match_any_sync(__activemask(), cls);
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(mine) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(counters + cls, __popc(mine));
  base = __shfl_sync(mine, base, leader);
  const int slot = base + __popc(mine & ((1u << lane) - 1));
  rows[cls * 3 * cap + slot] = row;
  if (cls == 0) {
    const int nch = (deg + SRB_HUB_CHUNK - 1) / SRB_HUB_CHUNK;
    const int first = atomicAdd(counters + 4, nch);
    hub_first[slot] = first;
    for (int c = 0; c < nch && first + c < hub_cap; ++c) {
      hub_work[2 * (first + c)] = row;
      hub_work[2 * (first + c) + 1] = c;
    }
  }
}

__global__ void finalize_losses_kernel(const float* bpr_losses, const float* nce_losses, int n_nce, float cl_rate, float* out) {
  float cl = 0.f;
  for (int q = 0; q < n_nce; ++q) cl += nce_losses[q];
  cl *= cl_rate;
  out[0] = bpr_losses[0];
  out[1] = bpr_losses[1];
  out[2] = cl;
  out[3] = bpr_losses[0] + bpr_losses[1] + cl;
}

struct Chain {
  const srb_graph_csr* adj;
  ScatterSegs final_segs;  // gradient w.r.t. the encoder's mean output (scale folded in)
  ScatterSegs cl_segs;     // gradient w.r.t. the output of layer `layer_cl`
  ScatterSegs ego_segs;    // gradient that lands on E0 directly
  int layer_cl;            // 1..L, or 0 when none / at the ego layer
  bool include_ego;
};

static int spmm_simple(const srb_step_desc* s, const srb_graph_csr* g, const float* x, float* y, const float* extra,
                       bool adam, cudaStream_t st, const uint32_t* col_mask = nullptr) {
  srb_spmm_desc p = {};
  p.col_mask = col_mask;
  p.rowptr = g->rowptr;
  p.colidx = g->colidx;
  p.vals = g->vals;
  p.row_order = g->row_order;
  p.n_long_rows = g->n_long_rows;
  p.n_vlong_rows = g->n_vlong_rows;
  p.hub = g->hub;
  p.n_rows = p.n_cols = s->n_users + s->n_items;
  p.d = s->d;
  p.X = x;
  p.Y = y;
  p.extra = extra;
  p.extra_scale = 1.f;
  if (adam) {
    p.adam_p = s->params;
    p.adam_m = s->adam_m;
    p.adam_v = s->adam_v;
    p.adam_scalars = s->scalars;
    p.beta1 = s->beta1;
    p.beta2 = s->beta2;
    p.adam_eps = s->adam_eps;
  }
  return srb_spmm_csr(&p, st);
}

// Horner backward of one encoder.  `gd` accumulates the E0 gradient across chains; the last
// chain applies Adam.  gd_live says whether gd already holds earlier chains' contributions.
// BPR + L2 and InfoNCE only read the encoder outputs and write disjoint buffers: the step forks BPR onto a
// side stream (event dependencies, so a stream capture records the fork and join) and joins before the
// losses are combined.
struct ForkRes {
  cudaStream_t side;
  cudaEvent_t fork, join;
  bool ok;
};
// Default resources: one side stream + two events per device, created on first use (before any capture: capture()
// warms up eagerly).  An engine that may run beside another one on the same device brings its own through
// srb_step_desc.fork_stream / fork_event / join_event, so that two engines never re-record each other's events.
static ForkRes* fork_res(const srb_step_desc* s, ForkRes* own) {
  if (s->fork_stream && s->fork_event && s->join_event) {
    own->side = (cudaStream_t)s->fork_stream;
    own->fork = (cudaEvent_t)s->fork_event;
    own->join = (cudaEvent_t)s->join_event;
    own->ok = true;
    return own;
  }
  static ForkRes res[64] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  ForkRes& r = res[dev];
  if (!r.ok) {
    if (cudaStreamCreateWithFlags(&r.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&r.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    r.ok = true;
  }
  return &r;
}

static ScatterSegs merged(const ScatterSegs& a, const ScatterSegs* b) {  // one launch instead of two
  ScatterSegs m = a;
  if (b)
    for (int q = 0; q < b->count && m.count < 8; ++q) m.s[m.count++] = b->s[q];
  return m;
}

static int run_chain(const srb_step_desc* s, const Ws& w, const Chain& c, bool* gd_live, bool last, cudaStream_t st) {
  const int L = s->n_layers, d = s->d;
  const size_t bytes = (size_t)(s->n_users + s->n_items) * d * 4;
  // seed: gradient w.r.t. the output of layer L
  SRB_TRY(check_cuda(cudaMemsetAsync(w.acc0, 0, bytes, st), "chain memset"));
  SRB_TRY(scatter_segments(w.acc0, d, merged(c.final_segs, c.layer_cl == L ? &c.cl_segs : nullptr), st));
  float* x = w.acc0;
  for (int k = L - 1; k >= 1; --k) {  // acc_k = A acc_{k+1} + (direct gradient of layer k)
    float* y = (x == w.acc0) ? w.acc1 : w.acc0;
    SRB_TRY(spmm_simple(s, c.adj, x, y, nullptr, false, st, k == L - 1 ? w.row_mask : nullptr));  // seed: batch rows only
    SRB_TRY(scatter_segments(y, d, merged(c.final_segs, c.layer_cl == k ? &c.cl_segs : nullptr), st));
    x = y;
  }
  const bool ego_add = (c.include_ego && c.final_segs.count) || c.ego_segs.count;
  if (ego_add && !*gd_live) {
    SRB_TRY(check_cuda(cudaMemsetAsync(w.gd, 0, bytes, st), "chain memset gd"));
    *gd_live = true;
  }
  if (ego_add) {
    if (c.include_ego) SRB_TRY(scatter_segments(w.gd, d, merged(c.final_segs, &c.ego_segs), st));
    else SRB_TRY(scatter_segments(w.gd, d, c.ego_segs, st));
  }
  const float* extra = *gd_live ? w.gd : nullptr;
  const uint32_t* mask = (L == 1) ? w.row_mask : nullptr;
  if (last) return spmm_simple(s, c.adj, x, nullptr, extra, true, st, mask);
  SRB_TRY(spmm_simple(s, c.adj, x, w.gd, extra, false, st, mask));
  *gd_live = true;
  return SRB_OK;
}

static ScatterSeg seg(const float* src, const int32_t* rows, const int32_t* n_dev, int n, int row_off, float scale) {
  ScatterSeg g = {src, rows, n_dev, n, row_off, scale, 0, 0};
  return g;
}

static int encoder(const srb_step_desc* s, const Ws& w, const srb_graph_csr* g, bool include_ego, int noise_mode, int view,
                   int layer_cl, float* final_out, float* cl_out, cudaStream_t st, const float* x1 = nullptr) {
  // training forward: the final mean is only read at the batch rows, so the last layer skips the rest
  srb_encoder_desc e = {};
  e.rowptr = g->rowptr;
  e.colidx = g->colidx;
  e.vals = g->vals;
  e.row_order = g->row_order;
  e.n_long_rows = g->n_long_rows;
  e.n_vlong_rows = g->n_vlong_rows;
  e.hub = g->hub;
  e.n = s->n_users + s->n_items;
  e.d = s->d;
  e.n_layers = s->n_layers;
  e.include_ego = include_ego;
  e.layer_cl = layer_cl;
  e.noise_mode = noise_mode;
  if (noise_mode == 1) e.noise = s->noise + (size_t)view * s->n_layers * e.n * e.d;
  e.eps = s->eps;
  e.philox_seed = s->philox_seed;
  e.philox_offset = ((uint64_t)view << 32) | 0x10u;
  e.philox_step_dev = s->step_dev;
  e.E0 = s->params;
  if (cl_out && layer_cl == s->n_layers) {
    e.final_out = final_out;  // the last layer is the CL view and is needed in full
  } else {
    e.last_rows = w.batch_rows;
    e.n_last_rows = 3 * s->batch_cap;
    e.last_rows_nv_dev = w.n_hub;
    e.last_rows_hub = srb_hub_split{0, w.hub_cap, w.hub_first, w.hub_work, w.hub_part};
    e.last_rows_out = final_out;  // batch rows of the mean land here; the running sum lives in w.rsum
    e.final_out = w.rsum;
  }
  e.cl_out = cl_out;
  e.work0 = w.work0;
  e.work1 = w.work1;
  e.x1 = x1;
  return srb_encoder_forward(&e, st);
}

// out = x + sign(x) * normalize(noise) * eps, row by row: the noise one perturbed SimGCL encoder adds to the shared
// first product (SimGCL.py:87-88), drawn exactly as the fused SpMM epilogue of layer 1 of view `view` would draw it.
static int perturb_rows(const srb_step_desc* s, const float* x, float* out, int view, cudaStream_t st) {
  const size_t nd = (size_t)(s->n_users + s->n_items) * s->d;
  srb_spmm_desc p = {};
  p.rowptr = s->adj.rowptr;  // (not read by the epilogue-only kernel)
  p.colidx = s->adj.colidx;
  p.vals = s->adj.vals;
  p.n_rows = p.n_cols = s->n_users + s->n_items;
  p.d = s->d;
  p.X = x;
  p.Y = out;
  p.extra_scale = 1.f;
  p.sum_scale = 1.f;
  p.noise_mode = s->noise_mode;
  if (s->noise_mode == 1) p.noise = s->noise + (size_t)view * s->n_layers * nd;
  p.eps = s->eps;
  p.philox_seed = s->philox_seed;
  p.philox_offset = ((uint64_t)view << 32) | 0x10u;
  p.philox_step_dev = s->step_dev;
  return srb_spmm_epilogue_rows(&p, st);
}

}  // namespace srb

extern "C" int64_t srb_step_workspace_bytes(int32_t model, int32_t n, int32_t d, int32_t batch_cap, int32_t n_hub_work) {
  srb_step_desc s = {};
  s.adj.hub.n_work = n_hub_work > 0 ? n_hub_work : 0;
  s.model = model;
  s.n_users = n;
  s.n_items = 0;
  s.d = d;
  s.batch_cap = batch_cap;
  return srb::carve(&s, nullptr, nullptr);
}

extern "C" int srb_train_step(const srb_step_desc* s, void* stream) {
  using namespace srb;
  SRB_REQUIRE(s != nullptr, "step: null desc");
  SRB_REQUIRE(s->model >= SRB_MODEL_MF && s->model <= SRB_MODEL_SGL, "step: unknown model %d", s->model);
  SRB_REQUIRE(s->d == 32 || s->d == 64 || s->d == 128, "step: unsupported d=%d (32, 64, 128)", s->d);
  SRB_REQUIRE(s->n_users > 0 && s->n_items > 0 && s->batch_cap > 0, "step: bad sizes");
  SRB_REQUIRE(s->params && s->adam_m && s->adam_v && s->step_dev && s->scalars && s->losses && s->batch, "step: null pointer");
  SRB_REQUIRE(s->model == SRB_MODEL_MF || s->n_layers >= 1, "step: graph models need n_layers >= 1");
  SRB_REQUIRE(s->model == SRB_MODEL_MF || (s->adj.rowptr && s->adj.colidx && s->adj.vals), "step: null adjacency");
  const int N = s->n_users + s->n_items;
  const int64_t need = srb_step_workspace_bytes(s->model, N, s->d, s->batch_cap, s->model != SRB_MODEL_MF ? s->adj.hub.n_work : 0);
  SRB_REQUIRE(s->workspace && s->workspace_bytes >= need, "step: workspace too small (%lld < %lld)",
              (long long)s->workspace_bytes, (long long)need);
  SRB_REQUIRE(((uintptr_t)s->workspace & 255) == 0, "step: workspace must be 256-byte aligned");
  Ws w;
  carve(s, &w, (char*)s->workspace);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = s->batch_cap, d = s->d, U = s->n_users, L = s->n_layers;
  const int32_t* hdr = s->batch;
  const int32_t* u_idx = s->batch + SRB_BATCH_HEADER;
  const int32_t* i_idx = u_idx + B;
  const int32_t* j_idx = i_idx + B;
  const int32_t* uq_u = j_idx + B;
  const int32_t* uq_i = uq_u + B;
  const int32_t* b_dev = hdr + 0;
  const int32_t* nu_dev = hdr + 1;
  const int32_t* ni_dev = hdr + 2;

  SRB_TRY(srb_adam_prepare(s->step_dev, s->scalars, s->lr, s->beta1, s->beta2, stream));

  // ---- forward ----
  if (s->model != SRB_MODEL_MF) {
    SRB_TRY(check_cuda(cudaMemsetAsync(w.n_hub, 0, 32 + (size_t)((U + s->n_items + 31) / 32) * 4, st), "row mask memset"));
    build_batch_rows_kernel<<<(3 * B + 255) / 256, 256, 0, st>>>(s->batch, B, U, s->adj.rowptr, 0, U + s->n_items, w.batch_rows, w.n_hub,
                                                                   w.row_mask, w.hub_cap ? w.hub_first : nullptr, w.hub_work, w.hub_cap);
    SRB_TRY(post_launch("build_batch_rows_kernel"));
  }
  const float* table = s->params;  // table BPR gathers from
  int n_nce = 0;
  switch (s->model) {
    case SRB_MODEL_MF: break;
    case SRB_MODEL_LIGHTGCN:
      SRB_TRY(encoder(s, w, &s->adj, true, 0, 0, 0, w.final_, nullptr, st));
      table = w.final_;
      break;
    case SRB_MODEL_XSIMGCL:
      SRB_REQUIRE(s->noise_mode == 1 || s->noise_mode == 2, "step: XSimGCL needs noise_mode 1 or 2");
      SRB_REQUIRE(s->noise_mode != 1 || s->noise, "step: noise tensor missing");
      SRB_TRY(encoder(s, w, &s->adj, false, s->noise_mode, 0, s->layer_cl, w.final_, w.cl, st));
      table = w.final_;
      break;
    case SRB_MODEL_SIMGCL:
      SRB_REQUIRE(s->noise_mode == 1 || s->noise_mode == 2, "step: SimGCL needs noise_mode 1 or 2");
      SRB_REQUIRE(s->noise_mode != 1 || s->noise, "step: noise tensor missing");
      if (L >= 2) {
        // layer 1 of the three encoders is the same product A * E0 (SimGCL.py:85); only the noise added to it differs
        // (:87-88).  It is evaluated once; acc0 / acc1 / gd belong to the backward pass and are free until then.
        SRB_TRY(spmm_simple(s, &s->adj, s->params, w.acc0, nullptr, false, st));
        SRB_TRY(perturb_rows(s, w.acc0, w.acc1, 0, st));
        SRB_TRY(perturb_rows(s, w.acc0, w.gd, 1, st));
        SRB_TRY(encoder(s, w, &s->adj, false, 0, 0, 0, w.final_, nullptr, st, w.acc0));
        SRB_TRY(encoder(s, w, &s->adj, false, s->noise_mode, 0, 0, w.cl, nullptr, st, w.acc1));
        SRB_TRY(encoder(s, w, &s->adj, false, s->noise_mode, 1, 0, w.v2, nullptr, st, w.gd));
      } else {
        SRB_TRY(encoder(s, w, &s->adj, false, 0, 0, 0, w.final_, nullptr, st));
        SRB_TRY(encoder(s, w, &s->adj, false, s->noise_mode, 0, 0, w.cl, nullptr, st));
        SRB_TRY(encoder(s, w, &s->adj, false, s->noise_mode, 1, 0, w.v2, nullptr, st));
      }
      table = w.final_;
      break;
    case SRB_MODEL_SGL:
      SRB_REQUIRE(s->adj_view[0].rowptr && s->adj_view[1].rowptr, "step: SGL needs two view graphs");
      SRB_TRY(encoder(s, w, &s->adj, true, 0, 0, 0, w.final_, nullptr, st));
      SRB_TRY(encoder(s, w, &s->adj_view[0], true, 0, 0, 0, w.cl, nullptr, st));
      SRB_TRY(encoder(s, w, &s->adj_view[1], true, 0, 0, 0, w.v2, nullptr, st));
      table = w.final_;
      break;
  }

  // ---- BPR + L2 (on the side stream when an InfoNCE follows) ----
  ForkRes fk_own = {};
  ForkRes* fk = (s->model == SRB_MODEL_XSIMGCL || s->model == SRB_MODEL_SIMGCL || s->model == SRB_MODEL_SGL) ? fork_res(s, &fk_own) : nullptr;
  if (fk) {
    SRB_TRY(check_cuda(cudaEventRecord(fk->fork, st), "fork record"));
    SRB_TRY(check_cuda(cudaStreamWaitEvent(fk->side, fk->fork, 0), "fork wait"));
  }
  {
    srb_bpr_desc p = {};
    p.emb = table;
    p.l2_emb = (s->model == SRB_MODEL_LIGHTGCN) ? s->params : table;  // LightGCN.py:25 regularises raw params
    p.n_users = U;
    p.d = d;
    p.u_idx = u_idx;
    p.i_idx = i_idx;
    p.j_idx = j_idx;
    p.b_dev = b_dev;
    p.b = B;
    p.emb_scale = 1.f;
    p.reg = s->reg;
    // (u,p,n)/batch_size: MF.py:21, LightGCN.py:25 | (u,p): SimGCL.py:31, XSimGCL.py:33 | (u,p,n): SGL.py:36
    p.l2_terms = (s->model == SRB_MODEL_SIMGCL || s->model == SRB_MODEL_XSIMGCL) ? 2 : 3;
    p.l2_div = s->l2_div;
    p.grad_scale = 1.f;
    p.losses = w.bpr_losses;
    p.g_emb = w.g_emb;
    p.g_l2 = (s->model == SRB_MODEL_LIGHTGCN) ? w.g_l2 : nullptr;
    p.scratch = w.bpr_scratch;
    SRB_TRY(srb_bpr_l2_fwd_bwd(&p, fk ? (void*)fk->side : stream));
    if (fk) SRB_TRY(check_cuda(cudaEventRecord(fk->join, fk->side), "join record"));
  }

  // ---- InfoNCE ----
  const size_t gn = (size_t)2 * B * d;
  float* g1a = w.g_nce;
  float* g2a = w.g_nce + gn;
  float* g1b = w.g_nce + 2 * gn;
  float* g2b = w.g_nce + 3 * gn;
  if (s->model == SRB_MODEL_XSIMGCL || s->model == SRB_MODEL_SIMGCL) {
    srb_infonce_desc q = {};
    q.n_problems = 2;
    q.d = d;
    q.b_cos = 1;
    q.temperature = s->tau;
    const float* t1 = (s->model == SRB_MODEL_XSIMGCL) ? w.final_ : w.cl;
    const float* t2 = (s->model == SRB_MODEL_XSIMGCL) ? w.cl : w.v2;
    q.prob[0] = {t1, t2, 0, 0, 1.f, 1.f, uq_u, nu_dev, B, s->cl_rate, g1a, g2a, w.nce_losses + 0};
    q.prob[1] = {t1, t2, U, U, 1.f, 1.f, uq_i, ni_dev, B, s->cl_rate, g1b, g2b, w.nce_losses + 1};
    q.workspace = w.nce_ws;
    q.workspace_bytes = w.nce_ws_bytes;
    SRB_TRY(srb_infonce_fwd_bwd(&q, stream));
    n_nce = 2;
  } else if (s->model == SRB_MODEL_SGL) {
    build_cat_idx_kernel<<<8, 256, 0, st>>>(s->batch, B, U, w.idx_cat, w.n_cat);
    SRB_TRY(post_launch("build_cat_idx_kernel"));
    srb_infonce_desc q = {};
    q.n_problems = 1;
    q.d = d;
    q.b_cos = 1;
    q.temperature = s->tau;
    q.prob[0] = {w.cl, w.v2, 0, 0, 1.f, 1.f, w.idx_cat, w.n_cat, 2 * B, s->cl_rate, g1a, g2a, w.nce_losses + 0};
    q.workspace = w.nce_ws;
    q.workspace_bytes = w.nce_ws_bytes;
    SRB_TRY(srb_infonce_fwd_bwd(&q, stream));
    n_nce = 1;
  }
  if (fk) SRB_TRY(check_cuda(cudaStreamWaitEvent(st, fk->join, 0), "join wait"));
  finalize_losses_kernel<<<1, 1, 0, st>>>(w.bpr_losses, w.nce_losses, n_nce, s->cl_rate, s->losses);
  SRB_TRY(post_launch("finalize_losses_kernel"));

  // ---- backward + Adam ----
  const size_t plane = (size_t)B * d;
  if (s->model == SRB_MODEL_MF) {
    const size_t bytes = (size_t)N * d * 4;
    SRB_TRY(check_cuda(cudaMemsetAsync(w.acc0, 0, bytes, st), "mf memset"));
    ScatterSegs sg;
    sg.count = 3;
    sg.s[0] = seg(w.g_emb, u_idx, b_dev, B, 0, 1.f);
    sg.s[1] = seg(w.g_emb + plane, i_idx, b_dev, B, U, 1.f);
    sg.s[2] = seg(w.g_emb + 2 * plane, j_idx, b_dev, B, U, 1.f);
    SRB_TRY(scatter_segments(w.acc0, d, sg, st));
    return srb_adam_step(s->params, s->adam_m, s->adam_v, w.acc0, (int64_t)N * d, s->scalars, s->beta1, s->beta2, s->adam_eps,
                         stream);
  }
  const float cm = 1.f / (float)((s->model == SRB_MODEL_LIGHTGCN || s->model == SRB_MODEL_SGL) ? L + 1 : L);
  bool gd_live = false;
  if (s->model == SRB_MODEL_SGL) {
    // two view encoders on their own graphs, then the main chain with Adam
    for (int v = 0; v < 2; ++v) {
      Chain c = {};
      c.adj = &s->adj_view[v];
      c.include_ego = true;
      c.final_segs.count = 1;
      c.final_segs.s[0] = seg(v == 0 ? g1a : g2a, w.idx_cat, w.n_cat, 2 * B, 0, cm);
      SRB_TRY(run_chain(s, w, c, &gd_live, false, st));
    }
  }
  Chain c = {};
  c.adj = &s->adj;
  c.include_ego = (s->model == SRB_MODEL_LIGHTGCN || s->model == SRB_MODEL_SGL);
  ScatterSegs& f = c.final_segs;
  f.count = 3;
  f.s[0] = seg(w.g_emb, u_idx, b_dev, B, 0, cm);
  f.s[1] = seg(w.g_emb + plane, i_idx, b_dev, B, U, cm);
  f.s[2] = seg(w.g_emb + 2 * plane, j_idx, b_dev, B, U, cm);
  if (s->model == SRB_MODEL_LIGHTGCN) {
    c.ego_segs.count = 3;
    c.ego_segs.s[0] = seg(w.g_l2, u_idx, b_dev, B, 0, 1.f);
    c.ego_segs.s[1] = seg(w.g_l2 + plane, i_idx, b_dev, B, U, 1.f);
    c.ego_segs.s[2] = seg(w.g_l2 + 2 * plane, j_idx, b_dev, B, U, 1.f);
  } else if (s->model == SRB_MODEL_XSIMGCL) {
    // view 1 = final (mean) rows, view 2 = layer l* output (XSimGCL.py:45-50)
    f.count = 5;
    f.s[3] = seg(g1a, uq_u, nu_dev, B, 0, cm);
    f.s[4] = seg(g1b, uq_i, ni_dev, B, U, cm);
    ScatterSegs& cs = (s->layer_cl >= 1 && s->layer_cl <= L) ? c.cl_segs : c.ego_segs;
    cs.count = 2;
    cs.s[0] = seg(g2a, uq_u, nu_dev, B, 0, 1.f);
    cs.s[1] = seg(g2b, uq_i, ni_dev, B, U, 1.f);
    c.layer_cl = (s->layer_cl >= 1 && s->layer_cl <= L) ? s->layer_cl : 0;
  } else if (s->model == SRB_MODEL_SIMGCL) {
    // all three encoders are the same linear map of E0: one merged chain
    f.count = 7;
    f.s[3] = seg(g1a, uq_u, nu_dev, B, 0, cm);
    f.s[4] = seg(g1b, uq_i, ni_dev, B, U, cm);
    f.s[5] = seg(g2a, uq_u, nu_dev, B, 0, cm);
    f.s[6] = seg(g2b, uq_i, ni_dev, B, U, cm);
  }
  return run_chain(s, w, c, &gd_live, true, st);
}
