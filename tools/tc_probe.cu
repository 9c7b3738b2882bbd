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


// ---- (E) MN-major B operand: D[128 x 64] = A[128 x 64] * B[64(k) x 64(n)], B row-major (n contiguous) ----
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// layout 0: two TMA boxes [64 rows][32 floats] (n-chunk c at c*8192)          -> lbo 8192, sbo 1024, k-step 1024
// layout 1: 16 TMA boxes [8 rows][32 floats], k-group g / n-chunk c at g*2048 + c*1024 -> lbo 1024, sbo 2048, k-step 2048
__global__ void __launch_bounds__(128) mn_probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB64,
                                               const __grid_constant__ CUtensorMap tmB8, float* D, int layout, uint32_t lbo, uint32_t sbo,
                                               uint32_t kstep, int a_tmem_mode, int set_bit, const float* Bglob) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  float* sA = (float*)base;                    // 2 chunks x [128][32]
  uint8_t* sB = base + 2 * 128 * 128;          // 16 KB
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base, 128);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_full, 2 * 128 * 128 + (layout == 3 ? 0 : 64 * 64 * 4));
    for (int c = 0; c < 2; ++c) tma_load_2d(sA + c * 128 * 32, &tmA, &bar_full, c * 32, 0);
    if (layout == 0 || layout == 2) {  // layout 2: tmB64 describes B^T [n][k] (K-major operand)
      for (int c = 0; c < 2; ++c) tma_load_2d(sB + c * 8192, &tmB64, &bar_full, c * 32, 0);
    } else if (layout == 3) {
    } else {
      for (int g = 0; g < 8; ++g)
        for (int c = 0; c < 2; ++c) tma_load_2d(sB + g * 2048 + c * 1024, &tmB8, &bar_full, c * 32, g * 8);
    }
    mbar_wait(&bar_full, 0);
  }
  if (layout == 3) {  // MN-major, no swizzle: core matrix = 8 k-rows x 16 B; (n/4) groups sbo apart, (k/8) groups lbo apart
    for (int e = threadIdx.x; e < 64 * 64; e += blockDim.x) {
      const int k = e / 64, n = e % 64;
      *reinterpret_cast<float*>(sB + (n / 4) * 128 + (k / 8) * 2048 + (k % 8) * 16 + (n % 4) * 4) = Bglob[e];
    }
    fence_proxy_async();
  }
  __syncthreads();
  if (a_tmem_mode) {  // A operand from tensor memory: thread = row, columns [64, 128) of the allocation
    const int row = threadIdx.x;
    // read my A row back from the swizzled shared tile
    uint32_t r[32];
    for (int c = 0; c < 2; ++c) {
      for (int u = 0; u < 8; ++u) {
        const float4 v = *reinterpret_cast<const float4*>((const uint8_t*)sA + c * 16384 + row * 128 + ((u ^ (row & 7)) << 4));
        r[4 * u] = __float_as_uint(v.x), r[4 * u + 1] = __float_as_uint(v.y), r[4 * u + 2] = __float_as_uint(v.z), r[4 * u + 3] = __float_as_uint(v.w);
      }
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::
              "r"(tm + ((uint32_t)(warp * 32) << 16) + 64 + c * 32),
          "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
          "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
          "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    fence_before_sync();
    __syncthreads();
  }
  if (warp == 0 && elect_one()) {
    fence_after_sync();
    const uint32_t idesc = make_idesc_tf32(128, 64) | (set_bit ? (1u << 16) : 0u);  // B MN-major
    for (int kk = 0; kk < 8; ++kk) {
      uint64_t db = desc_mn_sw128(smem_u32(sB) + kk * kstep, lbo, sbo);
      if (layout == 2) db = make_smem_desc_k_sw128(smem_u32(sB) + (kk >> 2) * 8192 + (kk & 3) * 32);
      if (layout == 3) db &= ~((uint64_t)7 << 61);  // SWIZZLE_NONE
      if (a_tmem_mode) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm),
            "r"(tm + 64 + kk * 8), "l"(db), "r"(idesc), "r"(kk ? 1u : 0u)
            : "memory");
      } else {
        const uint64_t da = make_smem_desc_k_sw128(smem_u32(sA + (kk >> 2) * 128 * 32) + (kk & 3) * 32);
        umma_tf32_ss(tm, da, db, idesc, kk ? 1u : 0u);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    const int row = warp * 32 + (threadIdx.x & 31);
    for (int j = 0; j < 32; ++j) D[(size_t)row * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 128);
}

// rate of tcgen05.mma with the A operand in tensor memory
template <int N>
__global__ void __launch_bounds__(64) rate_ts_probe(int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_mma;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (N * 128) / 4; i += blockDim.x) ((float*)base)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  if (warp == 1) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tmem_base;
  long long t0 = clock64();
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_tf32(128, N);
    const uint64_t db = make_smem_desc_k_sw128(smem_u32(base));
    for (int i = 0; i < n_mma; ++i)
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm + (i & 1) * N),
          "r"(tm + 2 * N + (i & 3) * 8), "l"(db), "r"(idesc), "r"(1u)
          : "memory");
    umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    out[0] = clock64() - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
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


static void run_mn(int layout, uint32_t lbo, uint32_t sbo, uint32_t kstep, int a_tmem, int set_bit = 1) {
  const int M = 128, K = 64, N = 64;
  std::vector<float> A(M * K), B(K * N), D(M * N, -1.f);
  srand(2);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  std::vector<float> Bt(N * K);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) Bt[n * K + k] = B[k * N + n];
  float* dBt;
  CK(cudaMalloc(&dBt, Bt.size() * 4));
  CK(cudaMemcpy(dBt, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, D.size() * 4));
  CUtensorMap tA, tB64, tB8;
  if (srb::make_tmap_f32_rows(&tA, dA, M, K, 128) || srb::make_tmap_f32_rows(&tB64, layout == 2 ? dBt : dB, K, N, 64) ||
      srb::make_tmap_f32_rows(&tB8, dB, K, N, 8)) {
    printf("tensor map encode failed\n");
    exit(1);
  }
  const size_t smem = 2 * 128 * 128 + 16384 + 1024;
  CK(cudaFuncSetAttribute(mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mn_probe<<<1, 128, smem>>>(tA, tB64, tB8, dD, layout, lbo, sbo, kstep, a_tmem, set_bit, dB);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("mn_probe layout=%d lbo=%u sbo=%u kstep=%u a_tmem=%d: CUDA error %s\n", layout, lbo, sbo, kstep, a_tmem, cudaGetErrorString(e));
    exit(1);
  }
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, ref00[4] = {0, 0, 0, 0};
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0;
      for (int k = 0; k < K; ++k) st += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[k * N + n]);
      if (m == 0 && n < 4) ref00[n] = st;
      max_err = fmax(max_err, fabs(D[m * N + n] - st));
    }
  printf("mn_probe layout=%d lbo=%u sbo=%u kstep=%u a_tmem=%d bit=%d: max err = %.3e -> %s   D[0][0..3] = %f %f %f %f   ref %f %f %f %f\n", layout,
         lbo, sbo, kstep, a_tmem, set_bit, max_err, max_err < 1e-3 ? "OK" : "MISMATCH", D[0], D[1], D[2], D[3], ref00[0], ref00[1], ref00[2],
         ref00[3]);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
}

template <int N>
static void run_rate_ts(int n_mma) {
  long long* d;
  CK(cudaMalloc(&d, 16 * 8));
  CK(cudaMemset(d, 0, 16 * 8));
  const size_t smem = N * 128 + 1024;
  CK(cudaFuncSetAttribute(rate_ts_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rate_ts_probe<N><<<1, 64, smem>>>(n_mma, d);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost));
  printf("mma, A in tensor memory            N=%3d: %.1f cyc per MMA (128xNx8 tf32)\n", N, (double)h[0] / n_mma);
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1) {  // one variant per process (a bad descriptor can fault)
    const int v = atoi(argv[1]);
    switch (v) {
      case 0: run_mn(2, 0, 0, 0, 0, 0); break;            // SS, B K-major (sanity)
      case 1: run_mn(2, 0, 0, 0, 1, 0); break;            // TS, B K-major
      case 2: run_mn(0, 8192, 1024, 1024, 0, 1); break;   // SS, B MN-major SW128
      case 3: run_mn(0, 8192, 1024, 1024, 0, 0); break;   // same descriptor, major bit clear
      case 4: run_mn(3, 2048, 128, 2048, 0, 1); break;    // SS, B MN-major, no swizzle
      case 5: run_mn(3, 128, 2048, 2048, 0, 1); break;    // lbo / sbo swapped
      case 6: run_mn(1, 1024, 2048, 2048, 0, 1); break;   // SW128, atoms tiled MN-first
      case 7: run_mn(0, 8192, 1024, 1024, 1, 1); break;   // TS + MN-major SW128
    }
    return 0;
  }

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
