// microbench_bulk.cu -- per-SM throughput of cp.async.bulk (UBLKCP) global->shared vs copy size and
// number of stages in flight.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I. tools/microbench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../drtk_b200/csrc/tma.cuh"
using namespace drtk;

// Each CTA streams `iters` chunks of CHUNK bytes (as CHUNK/COPY copies of COPY bytes) through a ring of
// STAGES stages; consumers only wait (no compute) -> pure load throughput.
// Same, but the copies of a chunk are issued by NW different warps (lane 0 of warps 0..NW-1).
template <int STAGES, int NW>
__global__ void __launch_bounds__(256, 1) bulk_kernel_mw(const char* src, size_t bytes_per_cta, int chunk, int copy, int iters, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long full[STAGES];
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init((uint64_t*)&full[s], NW); mbar_fence_init(); }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  const int ncopies = chunk / copy;
  auto issue = [&](int it) {  // called by lane 0 of warps < NW
    const int s = it % STAGES;
    const int mine = (ncopies - wid + NW - 1) / NW;
    mbar_arrive_expect_tx((uint64_t*)&full[s], mine * copy);
    for (int k = wid; k < ncopies; k += NW) bulk_g2s(smem + (size_t)s * chunk + (size_t)k * copy, base + ((size_t)it * chunk + (size_t)k * copy) % bytes_per_cta, copy, (uint64_t*)&full[s]);
  };
  if (lane == 0 && wid < NW) for (int it = 0; it < STAGES - 1 && it < iters; ++it) issue(it);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (lane == 0 && wid < NW && it + STAGES - 1 < iters) issue(it + STAGES - 1);
    const int s = it % STAGES;
    mbar_wait((uint64_t*)&full[s], (it / STAGES) & 1);
    acc += reinterpret_cast<float*>(smem + (size_t)s * chunk)[tid];
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int STAGES>
__global__ void __launch_bounds__(128, 1) bulk_kernel(const char* src, size_t bytes_per_cta, int chunk, int copy, int iters, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long full[STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init((uint64_t*)&full[s], 1); mbar_fence_init(); }
  __syncthreads();
  const char* base = src + (size_t)blockIdx.x * bytes_per_cta;
  auto issue = [&](int it) {
    const int s = it % STAGES;
    mbar_arrive_expect_tx((uint64_t*)&full[s], chunk);
    for (int o = 0; o < chunk; o += copy) bulk_g2s(smem + (size_t)s * chunk + o, base + ((size_t)it * chunk + o) % bytes_per_cta, copy, (uint64_t*)&full[s]);
  };
  if (tid == 0) for (int it = 0; it < STAGES - 1 && it < iters; ++it) issue(it);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    if (tid == 0 && it + STAGES - 1 < iters) issue(it + STAGES - 1);
    const int s = it % STAGES;
    mbar_wait((uint64_t*)&full[s], (it / STAGES) & 1);
    acc += reinterpret_cast<float*>(smem + (size_t)s * chunk)[tid];
    __syncthreads();
  }
  if (acc == 123.456f) sink[0] = acc;
}

template <int STAGES>
void run(const char* src, size_t per_cta, int chunk, int copy, float* sink) {
  const int iters = 400;
  const size_t smem = (size_t)STAGES * chunk;
  cudaFuncSetAttribute(bulk_kernel<STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  bulk_kernel<STAGES><<<148, 128, smem>>>(src, per_cta, chunk, copy, iters, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  bulk_kernel<STAGES><<<148, 128, smem>>>(src, per_cta, chunk, copy, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double gb = 148.0 * iters * chunk / 1e9;
  printf("stages %d chunk %6d B copy %5d B (%2d copies/chunk): %7.1f GB/s  (%.2f us per chunk per SM)  %s\n", STAGES, chunk, copy, chunk / copy,
         gb / (ms * 1e-3), ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
}

template <int STAGES, int NW>
void run_mw(const char* src, size_t per_cta, int chunk, int copy, float* sink) {
  const int iters = 400;
  const size_t smem = (size_t)STAGES * chunk;
  cudaFuncSetAttribute(bulk_kernel_mw<STAGES, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  bulk_kernel_mw<STAGES, NW><<<148, 256, smem>>>(src, per_cta, chunk, copy, iters, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  bulk_kernel_mw<STAGES, NW><<<148, 256, smem>>>(src, per_cta, chunk, copy, iters, sink);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double gb = 148.0 * iters * chunk / 1e9;
  printf("stages %d chunk %6d B copy %5d B issued by %d warps: %7.1f GB/s  (%.2f us per chunk per SM)  %s\n", STAGES, chunk, copy, NW,
         gb / (ms * 1e-3), ms * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t per_cta = 64ull << 20;  // 64 MiB per CTA -> 9.5 GB total, no L2 reuse
  char* src; cudaMalloc(&src, per_cta * 148); cudaMemset(src, 1, per_cta * 148);
  float* sink; cudaMalloc(&sink, 4);
  for (int copy : {512, 2048, 4096, 8192, 16384, 32768}) run<2>(src, per_cta, 65536, copy, sink);
  for (int copy : {4096, 16384}) run<3>(src, per_cta, 65536, copy, sink);
  for (int copy : {4096, 8192, 32768}) run<2>(src, per_cta, 32768, copy, sink);
  for (int copy : {4096, 8192}) run<4>(src, per_cta, 32768, copy, sink);
  for (int copy : {4096, 8192}) run<6>(src, per_cta, 32768, copy, sink);
  for (int copy : {4096}) run<8>(src, per_cta, 16384, copy, sink);
  run_mw<2, 2>(src, per_cta, 65536, 4096, sink);
  run_mw<2, 4>(src, per_cta, 65536, 4096, sink);
  run_mw<2, 8>(src, per_cta, 65536, 4096, sink);
  run_mw<2, 4>(src, per_cta, 65536, 2048, sink);
  run_mw<2, 8>(src, per_cta, 65536, 512, sink);
  return 0;
}
