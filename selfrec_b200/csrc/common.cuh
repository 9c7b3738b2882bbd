// Shared helpers for the selfrec_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "selfrec_b200.h"

#define SRB_FULL_MASK 0xffffffffu

namespace srb {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SRB_ERR_CUDA;
  }
  return SRB_OK;
}

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
inline int post_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaPeekAtLastError(), what);
}

#define SRB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      srb::set_error(__VA_ARGS__);        \
      return SRB_ERR_ARG;                 \
    }                                     \
  } while (0)

#define SRB_TRY(expr)                     \
  do {                                    \
    int _rc = (expr);                     \
    if (_rc != SRB_OK) return _rc;        \
  } while (0)

int sm_count();

// sparse-row scatter with several (src, rows) segments in one launch (bpr.cu)
struct ScatterSeg {
  const float* src;     // [n, d] compact rows
  const int32_t* rows;  // [n] destination row ids
  const int32_t* n_dev; // optional device count
  int32_t n;            // capacity / host count
  int32_t row_off;
  float scale;
  int32_t row_lo, row_hi;  // optional filter (row_hi > row_lo): only rows[r] + row_off in [row_lo, row_hi), stored at - row_lo
  int32_t mod, rem;        // optional filter (mod > 0; cyclic row ownership): only rows with row % mod == rem, stored at row / mod
};
struct ScatterSegs {
  int count;
  ScatterSeg s[8];
};
int scatter_segments(float* dst, int d, const ScatterSegs& segs, cudaStream_t st);

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SRB_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void st4(float* p, const float4& v) {
  *reinterpret_cast<float4*>(p) = v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

__device__ __forceinline__ float4 f4_fma(float a, const float4& x, const float4& acc) {
  return make_float4(fmaf(a, x.x, acc.x), fmaf(a, x.y, acc.y), fmaf(a, x.z, acc.z),
                     fmaf(a, x.w, acc.w));
}

__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

__device__ __forceinline__ float4 f4_scale(float s, const float4& a) {
  return make_float4(s * a.x, s * a.y, s * a.z, s * a.w);
}

__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

__device__ __forceinline__ float sgnf(float x) {
  return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
}

// Philox4x32-10 (Salmon et al. 2011): counter-based generator for the perf-mode noise.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float u32_to_unit(uint32_t x) {
  return (float)(x >> 8) * (1.0f / 16777216.0f);  // [0, 1) with 24 random bits
}

}  // namespace srb
