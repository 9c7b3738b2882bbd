// L2 / HBM gather-bandwidth microbenchmark (evidence for DESIGN.md: the SpMM's X-row gathers
// are bound by L2->SM throughput).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2mb l2_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__global__ void stream_read(const float4* __restrict__ p, size_t n4, int reps, float* sink) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      float4 v = __ldg(p + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;
}

// each group of LPR lanes gathers one row of ROWB bytes per index; U independent rows in flight
template <int ROWB, int U>
__global__ void gather_rows(const float* __restrict__ tab, const int* __restrict__ idx, size_t n_idx, float* sink) {
  constexpr int LPR = ROWB / 16;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, cl = lane % LPR;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (size_t base = warp * GPW * U; base + GPW * U <= n_idx; base += nwarps * GPW * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = __ldg(idx + base + u * GPW + sub);
      v[u] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * (ROWB / 4)) + cl);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) *sink = acc;
}

template <typename F>
float time_ms(F f, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink;
  cudaMalloc(&sink, 4);
  printf("SMs %d\n", sms);
  for (size_t mb : {8, 16, 32, 64, 96, 256, 1024}) {
    size_t bytes = mb << 20;
    float4* p;
    cudaMalloc(&p, bytes);
    cudaMemset(p, 0, bytes);
    int reps = mb <= 96 ? 8 : 1;
    float ms = time_ms([&] { stream_read<<<sms * 8, 256>>>(p, bytes / 16, reps, sink); }, 10);
    printf("stream_read %5zu MB x%d: %8.1f GB/s\n", mb, reps, bytes * (double)reps / ms / 1e6);
    cudaFree(p);
  }
  const size_t n_idx = 1 << 22;
  for (size_t rows : {38048, 69716, 262144, 1048576, 8388608}) {
    std::vector<int> h(n_idx);
    unsigned s = 12345;
    for (size_t i = 0; i < n_idx; ++i) {
      s = s * 1664525u + 1013904223u;
      h[i] = (int)((s >> 8) % rows);
    }
    int* idx;
    cudaMalloc(&idx, n_idx * 4);
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    float* tab;
    cudaMalloc(&tab, rows * 512);
    cudaMemset(tab, 0, rows * 512);
#define RUN(ROWB, U, BLK)                                                                                   \
  {                                                                                                         \
    float ms = time_ms([&] { gather_rows<ROWB, U><<<sms * BLK, 256>>>(tab, idx, n_idx, sink); }, 10);       \
    printf("gather rows=%8zu rowB=%3d U=%d blk/SM=%d table=%7.1f MB: %8.1f GB/s\n", rows, ROWB, U, BLK,     \
           rows * (double)ROWB / 1e6, n_idx * (double)ROWB / ms / 1e6);                                     \
  }
    RUN(256, 4, 8)
    RUN(256, 8, 8)
    RUN(256, 8, 4)
    RUN(128, 8, 8)
    RUN(512, 4, 8)
    RUN(512, 8, 8)
    cudaFree(idx);
    cudaFree(tab);
  }
  return 0;
}
