// (i) CSR SpMM  Y = A * X  with fused epilogue (noise, layer sum, Adam, peer all-gather).
//
// Replaces torch.sparse.mm(self.sparse_norm_adj, ego_embeddings) -- LightGCN.py:72,
// SimGCL.py:85, XSimGCL.py:88, SGL.py:104-108 -- plus the elementwise tail of the encoders
// (XSimGCL.py:90-96) and, for the last backward product, torch.optim.Adam.step.
//
// Mapping: one warp per output row.  A row vector of D floats is held as one float4 per
// lane by LPR = D/4 lanes; the 32/LPR lane groups of the warp walk alternating non-zeros
// and are combined with xor-shuffles at the end.  (col, val) pairs are fetched 32 at a
// time with coalesced loads and broadcast by shuffle; the X-row gathers are 128-bit
// loads, UNROLL x (32/LPR) rows of X in flight per warp.  HBM/L2-bound integer+fp32 work:
// no tensor cores here by design.
#include "common.cuh"

namespace srb {

struct SpmmArgs {
  const int32_t* rowptr;
  const int32_t* colidx;
  const float* vals;
  const int32_t* row_order;
  int32_t n_rows;
  const float* X;
  float* Y;
  const float* extra;
  float extra_scale;
  int32_t noise_mode;
  const float* noise;
  float eps;
  uint2 pkey;
  uint2 poff;
  const int32_t* pstep;
  const float* sum_in;
  float* sum_out;
  float sum_scale;
  float* ap;
  float* am;
  float* av;
  const float* ascal;
  float b1, b2, aeps;
  int32_t world;
  int32_t row_begin;
  float* peer[8];
};

template <int D>
__global__ void __launch_bounds__(256) spmm_csr_kernel(const SpmmArgs a) {
  constexpr int LPR = D / 4;    // lanes holding one row vector
  constexpr int NZP = 32 / LPR; // non-zeros processed concurrently by the warp
  constexpr int UNROLL = 4;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int cl = lane % LPR;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;

  for (int w = warp0; w < a.n_rows; w += nwarps) {
    const int row = a.row_order ? __ldg(a.row_order + w) : w;
    const int beg = __ldg(a.rowptr + row);
    const int end = __ldg(a.rowptr + row + 1);
    float4 acc = f4_zero();
    for (int base = beg; base < end; base += 32) {
      const int idx = base + lane;
      int c = 0;
      float v = 0.f;
      if (idx < end) {
        c = __ldg(a.colidx + idx);
        v = __ldg(a.vals + idx);
      }
      const int cnt = min(32, end - base);
      for (int j = 0; j < cnt; j += NZP * UNROLL) {
        int cc[UNROLL];
        float vv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int jj = j + u * NZP + sub;
          cc[u] = __shfl_sync(SRB_FULL_MASK, c, jj & 31);
          vv[u] = __shfl_sync(SRB_FULL_MASK, v, jj & 31);
          if (jj >= cnt) vv[u] = 0.f, cc[u] = -1;
        }
        float4 x[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
          x[u] = (cc[u] >= 0) ? ldg4(a.X + (size_t)cc[u] * D + cl * 4) : f4_zero();
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc = f4_fma(vv[u], x[u], acc);
      }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
      acc.x += __shfl_xor_sync(SRB_FULL_MASK, acc.x, o);
      acc.y += __shfl_xor_sync(SRB_FULL_MASK, acc.y, o);
      acc.z += __shfl_xor_sync(SRB_FULL_MASK, acc.z, o);
      acc.w += __shfl_xor_sync(SRB_FULL_MASK, acc.w, o);
    }
    // ---- epilogue: every lane group holds the full row; group 0 stores ----
    const size_t off = (size_t)row * D + cl * 4;
    float4 y = acc;
    if (a.extra) {
      const float4 e = *reinterpret_cast<const float4*>(a.extra + off);
      y = f4_fma(a.extra_scale, e, y);
    }
    if (a.noise_mode) {
      float4 nz;
      if (a.noise_mode == 1) {
        nz = ldg4(a.noise + off);
      } else {
        const uint32_t stp = a.pstep ? (uint32_t)*a.pstep : 0u;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)row, (uint32_t)cl, a.poff.x, a.poff.y ^ stp), a.pkey);
        nz = make_float4(u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w));
      }
      float ss = f4_dot(nz, nz);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(SRB_FULL_MASK, ss, o);
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
      y.x += sgnf(y.x) * (nz.x / nrm) * a.eps;
      y.y += sgnf(y.y) * (nz.y / nrm) * a.eps;
      y.z += sgnf(y.z) * (nz.z / nrm) * a.eps;
      y.w += sgnf(y.w) * (nz.w / nrm) * a.eps;
    }
    if (sub == 0) {
      if (a.Y) st4(a.Y + off, y);
      if (a.world > 0) {
        const size_t goff = (size_t)(a.row_begin + row) * D + cl * 4;
#pragma unroll 1
        for (int g = 0; g < a.world; ++g) st4(a.peer[g] + goff, y);
      }
      if (a.sum_out) {
        float4 s = y;
        if (a.sum_in) s = f4_add(s, *reinterpret_cast<const float4*>(a.sum_in + off));
        st4(a.sum_out + off, f4_scale(a.sum_scale, s));
      }
      if (a.ap) {
        const float step_size = a.ascal[0];
        const float bc2_sqrt = a.ascal[1];
        float4 p = *reinterpret_cast<const float4*>(a.ap + off);
        float4 m = *reinterpret_cast<const float4*>(a.am + off);
        float4 v = *reinterpret_cast<const float4*>(a.av + off);
        const float w1 = 1.f - a.b1, w2 = 1.f - a.b2;
#define SRB_ADAM1(F)                                         \
  m.F = m.F + w1 * (y.F - m.F);                              \
  v.F = v.F * a.b2;                                          \
  v.F = v.F + (w2 * y.F) * y.F;                              \
  p.F = p.F - step_size * (m.F / (sqrtf(v.F) / bc2_sqrt + a.aeps));
        SRB_ADAM1(x) SRB_ADAM1(y) SRB_ADAM1(z) SRB_ADAM1(w)
#undef SRB_ADAM1
        st4(a.ap + off, p);
        st4(a.am + off, m);
        st4(a.av + off, v);
      }
    }
  }
}

static int launch_spmm(const SpmmArgs& a, int d, cudaStream_t st) {
  if (a.n_rows == 0) return SRB_OK;
  const int threads = 256;
  const int wpb = threads / 32;
  long long blocks = ((long long)a.n_rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 8;  // 8 x 256 threads = full residency
  if (blocks > cap) blocks = cap;
  switch (d) {
    case 32: spmm_csr_kernel<32><<<(int)blocks, threads, 0, st>>>(a); break;
    case 64: spmm_csr_kernel<64><<<(int)blocks, threads, 0, st>>>(a); break;
    case 128: spmm_csr_kernel<128><<<(int)blocks, threads, 0, st>>>(a); break;
    default: set_error("spmm: unsupported d=%d (32, 64, 128)", d); return SRB_ERR_ARG;
  }
  return post_launch("spmm_csr_kernel");
}

static int fill_args(const srb_spmm_desc* d, SpmmArgs& a) {
  SRB_REQUIRE(d != nullptr, "spmm: null desc");
  SRB_REQUIRE(d->rowptr && d->colidx && d->vals && d->X, "spmm: null CSR/X pointer");
  SRB_REQUIRE(d->n_rows >= 0 && d->n_cols >= 0, "spmm: negative shape");
  SRB_REQUIRE(d->noise_mode >= 0 && d->noise_mode <= 2, "spmm: bad noise_mode");
  SRB_REQUIRE(d->noise_mode != 1 || d->noise, "spmm: noise_mode 1 needs a noise tensor");
  SRB_REQUIRE(!d->adam_p || (d->adam_m && d->adam_v && d->adam_scalars), "spmm: incomplete adam pointers");
  SRB_REQUIRE(d->X != d->Y, "spmm: Y must not alias X");
  a.rowptr = d->rowptr;
  a.colidx = d->colidx;
  a.vals = d->vals;
  a.row_order = d->row_order;
  a.n_rows = d->n_rows;
  a.X = d->X;
  a.Y = d->Y;
  a.extra = d->extra;
  a.extra_scale = d->extra_scale;
  a.noise_mode = d->noise_mode;
  a.noise = d->noise;
  a.eps = d->eps;
  a.pkey = make_uint2((uint32_t)d->philox_seed, (uint32_t)(d->philox_seed >> 32));
  a.poff = make_uint2((uint32_t)d->philox_offset, (uint32_t)(d->philox_offset >> 32));
  a.pstep = d->philox_step_dev;
  a.sum_in = d->sum_in;
  a.sum_out = d->sum_out;
  a.sum_scale = d->sum_scale;
  a.ap = d->adam_p;
  a.am = d->adam_m;
  a.av = d->adam_v;
  a.ascal = d->adam_scalars;
  a.b1 = d->beta1;
  a.b2 = d->beta2;
  a.aeps = d->adam_eps;
  a.world = 0;
  a.row_begin = 0;
  for (int g = 0; g < 8; ++g) a.peer[g] = nullptr;
  return SRB_OK;
}

}  // namespace srb

extern "C" int srb_spmm_csr(const srb_spmm_desc* desc, void* stream) {
  srb::SpmmArgs a;
  SRB_TRY(srb::fill_args(desc, a));
  return srb::launch_spmm(a, desc->d, (cudaStream_t)stream);
}

extern "C" int srb_spmm_csr_allgather(const srb_spmm_sharded_desc* desc, void* stream) {
  SRB_REQUIRE(desc != nullptr, "spmm_allgather: null desc");
  SRB_REQUIRE(desc->world >= 1 && desc->world <= 8, "spmm_allgather: world must be 1..8");
  srb::SpmmArgs a;
  SRB_TRY(srb::fill_args(&desc->local, a));
  a.Y = nullptr;
  a.world = desc->world;
  a.row_begin = desc->row_begin;
  for (int g = 0; g < desc->world; ++g) {
    SRB_REQUIRE(desc->peer_Y[g] != nullptr, "spmm_allgather: null peer buffer %d", g);
    SRB_REQUIRE(desc->peer_Y[g] != desc->local.X, "spmm_allgather: peer buffer aliases X");
    a.peer[g] = desc->peer_Y[g];
  }
  return srb::launch_spmm(a, desc->local.d, (cudaStream_t)stream);
}

extern "C" int srb_encoder_forward(const srb_encoder_desc* e, void* stream) {
  SRB_REQUIRE(e != nullptr, "encoder: null desc");
  SRB_REQUIRE(e->E0 && e->final_out, "encoder: null E0/final_out");
  SRB_REQUIRE(e->n_layers >= 0, "encoder: negative n_layers");
  SRB_REQUIRE(e->n_layers == 0 || (e->work0 && e->work1), "encoder: work buffers required");
  SRB_REQUIRE(e->include_ego || e->n_layers > 0, "encoder: mean over zero layers");
  const size_t nd = (size_t)e->n * e->d;
  cudaStream_t st = (cudaStream_t)stream;
  const int L = e->n_layers;
  if (L == 0) {  // MF: the encoder is the identity
    SRB_TRY(srb::check_cuda(cudaMemcpyAsync(e->final_out, e->E0, nd * 4, cudaMemcpyDeviceToDevice, st), "encoder copy"));
    if (e->cl_out)
      SRB_TRY(srb::check_cuda(cudaMemcpyAsync(e->cl_out, e->E0, nd * 4, cudaMemcpyDeviceToDevice, st), "encoder copy"));
    return SRB_OK;
  }
  const float inv = 1.0f / (float)(e->include_ego ? L + 1 : L);
  const bool want_cl = e->cl_out != nullptr;
  const bool cl_hit = want_cl && e->layer_cl >= 1 && e->layer_cl <= L;
  if (want_cl && !cl_hit)  // XSimGCL.py:86: default CL view is the ego embedding
    SRB_TRY(srb::check_cuda(cudaMemcpyAsync(e->cl_out, e->E0, nd * 4, cudaMemcpyDeviceToDevice, st), "encoder copy"));
  const float* x = e->E0;
  for (int k = 0; k < L; ++k) {
    srb_spmm_desc s = {};
    s.rowptr = e->rowptr;
    s.colidx = e->colidx;
    s.vals = e->vals;
    s.row_order = e->row_order;
    s.n_rows = e->n;
    s.n_cols = e->n;
    s.d = e->d;
    s.X = x;
    const bool last = (k == L - 1);
    float* y = nullptr;
    if (cl_hit && k == e->layer_cl - 1) y = e->cl_out;  // this layer's output is the CL view
    else if (!last) y = (x == e->work0) ? e->work1 : e->work0;
    s.Y = y;
    s.noise_mode = e->noise_mode;
    if (e->noise_mode == 1) s.noise = e->noise + (size_t)k * nd;
    s.eps = e->eps;
    s.philox_seed = e->philox_seed;
    s.philox_offset = e->philox_offset + (uint64_t)k;
    s.philox_step_dev = e->philox_step_dev;
    // running sum lives in final_out; layer 1 seeds it (with E0 when the ego layer counts)
    s.sum_in = (k == 0) ? (e->include_ego ? e->E0 : nullptr) : e->final_out;
    s.sum_out = e->final_out;
    s.sum_scale = last ? inv : 1.0f;
    SRB_TRY(srb_spmm_csr(&s, stream));
    x = y;
  }
  return SRB_OK;
}
