// tcgen05 / TMA probe: (A) validates the descriptors used by the scoring kernel against a CPU
// GEMM, (B) measures cycles per tf32 MMA, (C) cycles per tcgen05.ld, (D) both concurrently.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../selfrec_b200/csrc -o tc_probe tc_probe.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "tc_common.cuh"

using namespace srb::tc;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

// One CTA: D[128 x N] = A[128 x 64] * B[N x 64]^T, tf32, operands via TMA (2 k-chunks of 32 floats).
template <int N>
__global__ void __launch_bounds__(128) gemm_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                 float* D, int ldd) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  float* sA = (float*)base;                       // 2 chunks x [128][32]
  float* sB = (float*)(base + 2 * 128 * 128);     // 2 chunks x [N][32]
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base, N < 32 ? 32 : N);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_full, 2 * 128 * 128 + 2 * N * 128);
    for (int c = 0; c < 2; ++c) {
      tma_load_2d(sA + c * 128 * 32, &tmA, &bar_full, c * 32, 0);
      tma_load_2d(sB + c * N * 32, &tmB, &bar_full, c * 32, 0);
    }
    mbar_wait(&bar_full, 0);
    fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(128, N);
    for (int c = 0; c < 2; ++c)
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_smem_desc_k_sw128(smem_u32(sA + c * 128 * 32) + k * 32);
        const uint64_t db = make_smem_desc_k_sw128(smem_u32(sB + c * N * 32) + k * 32);
        umma_tf32_ss(tm, da, db, idesc, (c | k) ? 1u : 0u);
      }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  // epilogue: warp w reads lanes 32w..32w+31
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 32; ++j) D[(size_t)row * ldd + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, N < 32 ? 32 : N);
}

// throughput probes; smem content is irrelevant (zeros)
template <int N>
__global__ void __launch_bounds__(288) rate_probe(int n_mma, int n_ld, int ld_warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (128 * 128 + N * 128) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 8) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  long long t0 = clock64();
  if (warp == 8) {
    if (n_mma > 0 && elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, N);
      const uint64_t da = make_smem_desc_k_sw128(smem_u32(base));
      const uint64_t db = make_smem_desc_k_sw128(smem_u32(base + 128 * 128));
      for (int i = 0; i < n_mma; ++i) umma_tf32_ss(tm + (i & 1) * 256, da, db, idesc, 1u);
      umma_commit(&bar_mma);
      mbar_wait(&bar_mma, 0);
      out[0] = clock64() - t0;
    }
  } else if (warp < ld_warps) {
    uint32_t acc = 0;
    for (int i = 0; i < n_ld; ++i) {
      uint32_t r[32];
      tmem_ld_32x32(tm + ((uint32_t)((warp & 3) * 32) << 16) + ((i * 32) & 255) + (warp >> 2) * 256, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    if (n_ld > 0 && (threadIdx.x & 31) == 0) out[1 + warp] = (clock64() - t0) + (acc == 0x12345u ? 1 : 0);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

template <int N>
static void run_gemm() {
  const int M = 128, K = 64;
  std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  CUtensorMap tA, tB;
  if (srb::make_tmap_f32_rows(&tA, dA, M, K, 128) || srb::make_tmap_f32_rows(&tB, dB, N, K, N)) {
    printf("tensor map encode failed\n");
    exit(1);
  }
  const size_t smem = 2 * 128 * 128 + 2 * N * 128 + 1024;
  CK(cudaFuncSetAttribute(gemm_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm_probe<N><<<1, 128, smem>>>(tA, tB, dD, N);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_err_trunc = 0, max_err_exact = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, se = 0;
      for (int k = 0; k < K; ++k) {
        st += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
        se += (double)A[m * K + k] * B[n * K + k];
      }
      max_err_trunc = fmax(max_err_trunc, fabs(D[m * N + n] - st));
      max_err_exact = fmax(max_err_exact, fabs(D[m * N + n] - se));
    }
  printf("gemm_probe N=%d: max|D - tf32trunc ref| = %.3e   max|D - exact| = %.3e   D[0][0]=%f D[127][%d]=%f  -> %s\n", N,
         max_err_trunc, max_err_exact, D[0], N - 1, D[127 * N + N - 1], max_err_exact < 0.05 ? "OK" : "MISMATCH");
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

template <int N>
static void run_rate(int n_mma, int n_ld, int ld_warps, const char* tag) {
  long long* d;
  CK(cudaMalloc(&d, 16 * 8));
  CK(cudaMemset(d, 0, 16 * 8));
  const size_t smem = 128 * 128 + N * 128 + 1024;
  CK(cudaFuncSetAttribute(rate_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_probe<N><<<1, 288, smem>>>(n_mma, n_ld, ld_warps, d);
  CK(cudaDeviceSynchronize());
  long long h[16];
  CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
  printf("%-34s N=%3d: ", tag, N);
  if (n_mma) printf("%lld cyc / %d mma = %.1f cyc per MMA (128xNx8 tf32)   ", h[0], n_mma, (double)h[0] / n_mma);
  if (n_ld) {
    long long mx = 0;
    for (int w = 0; w < ld_warps; ++w) mx = h[1 + w] > mx ? h[1 + w] : mx;
    printf("%lld cyc / %d ld.x32 per warp (%d warps) = %.1f cyc per ld -> %.1f B/cyc/SM", mx, n_ld, ld_warps, (double)mx / n_ld,
           ld_warps * 32.0 * 32 * 4 * n_ld / mx);
  }
  printf("\n");
  cudaFree(d);
}

int main() {
  run_gemm<128>();
  run_gemm<256>();
  run_gemm<64>();
  run_rate<128>(2000, 0, 0, "mma only");
  run_rate<256>(2000, 0, 0, "mma only");
  run_rate<128>(0, 2000, 4, "ld only, 4 warps");
  run_rate<128>(0, 2000, 8, "ld only, 8 warps");
  run_rate<128>(0, 2000, 1, "ld only, 1 warp");
  run_rate<128>(2000, 2000, 4, "mma + ld 4 warps");
  run_rate<256>(2000, 4000, 8, "mma + ld 8 warps");
  return 0;
}
