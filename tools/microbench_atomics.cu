// microbench_atomics.cu -- how expensive are the gradient-scatter building blocks on B200?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_atomics.cu -o tools/microbench_atomics
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// mode 0: each half-warp adds 16 contiguous floats to a pseudo-random row (scalar RED per lane)
// mode 1: 4 lanes per row, red.v4 (16 floats per row)
// mode 2: every lane a pseudo-random scalar address
// mode 3: all lanes of a warp the same row, same 16 floats (two lanes per address)
template <int MODE>
__global__ void red_kernel(float* table, int V, int iters) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      const uint32_t row = hash((tid >> 4) * 9781u + it * 6151u) % V;
      asm volatile("red.global.add.f32 [%0], %1;" :: "l"(table + (size_t)row * 16 + (lane & 15)), "f"(1.0f) : "memory");
    } else if (MODE == 1) {
      const uint32_t row = hash((tid >> 2) * 9781u + it * 6151u) % V;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(table + (size_t)row * 16 + (lane & 3) * 4), "f"(1.0f) : "memory");
    } else if (MODE == 2) {
      const uint32_t idx = hash(tid * 9781u + it * 6151u) % (V * 16);
      asm volatile("red.global.add.f32 [%0], %1;" :: "l"(table + idx), "f"(1.0f) : "memory");
    } else {
      const uint32_t row = hash((tid >> 5) * 9781u + it * 6151u) % V;
      asm volatile("red.global.add.f32 [%0], %1;" :: "l"(table + (size_t)row * 16 + (lane & 15)), "f"(1.0f) : "memory");
    }
  }
}

// shared-memory atomics: MODE 0 float add distinct banks, 1 int add distinct banks, 2 float add 16 rows x 16 (2 lanes/addr),
// 3 u64 atomicMin distinct addresses
template <int MODE>
__global__ void smem_kernel(float* out, int iters) {
  __shared__ float sf[4096];
  __shared__ unsigned long long s64[1024];
  int* si = reinterpret_cast<int*>(sf);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sf[i] = 0.f;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s64[i] = ~0ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0; it < iters; ++it) {
    const uint32_t r = hash(warp * 131u + it * 7919u);
    if (MODE == 0) atomicAdd(&sf[(r % 127) * 32 + lane], 1.0f);
    else if (MODE == 1) atomicAdd(&si[(r % 127) * 32 + lane], 1);
    else if (MODE == 2) atomicAdd(&sf[(r % 255) * 16 + (lane & 15)], 1.0f);
    else atomicMin(&s64[(r % 31) * 32 + lane], (unsigned long long)(r + lane) << 32);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = sf[5] + (float)s64[7];
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  const int V = 50625; float* table; cudaMalloc(&table, sizeof(float) * V * 16 * 8); cudaMemset(table, 0, sizeof(float) * V * 16 * 8);
  float* out; cudaMalloc(&out, 4096 * 4);
  const int blocks = 148 * 8, threads = 256, iters = 256;
  const double lane_ops = (double)blocks * threads * iters;
  float ms;
  ms = time_ms([&] { red_kernel<0><<<blocks, threads>>>(table, V, iters); });
  printf("REDG f32, 16 contiguous lanes per random row : %.3f ms  %.1f G lane-ops/s  (%.1f G rows/s)\n", ms, lane_ops / ms / 1e6, lane_ops / 16 / ms / 1e6);
  ms = time_ms([&] { red_kernel<1><<<blocks, threads>>>(table, V, iters); });
  printf("REDG v4.f32, 4 lanes per random row         : %.3f ms  %.1f G lane-ops/s  (%.1f G rows/s, %.1f G floats/s)\n", ms, lane_ops / ms / 1e6, lane_ops / 4 / ms / 1e6, lane_ops * 4 / ms / 1e6);
  ms = time_ms([&] { red_kernel<2><<<blocks, threads>>>(table, V, iters); });
  printf("REDG f32, every lane a random address       : %.3f ms  %.1f G lane-ops/s\n", ms, lane_ops / ms / 1e6);
  ms = time_ms([&] { red_kernel<3><<<blocks, threads>>>(table, V, iters); });
  printf("REDG f32, warp on one row (2 lanes/address)  : %.3f ms  %.1f G lane-ops/s\n", ms, lane_ops / ms / 1e6);
  const int sblocks = 148 * 4, siters = 2048;
  const double sops = (double)sblocks * threads * siters;
  ms = time_ms([&] { smem_kernel<0><<<sblocks, threads>>>(out, siters); });
  printf("ATOMS float add, conflict-free               : %.3f ms  %.2f lane-ops/clk/SM @1.9GHz\n", ms, sops / (ms * 1e-3) / 148 / 1.9e9);
  ms = time_ms([&] { smem_kernel<1><<<sblocks, threads>>>(out, siters); });
  printf("ATOMS int add, conflict-free                 : %.3f ms  %.2f lane-ops/clk/SM\n", ms, sops / (ms * 1e-3) / 148 / 1.9e9);
  ms = time_ms([&] { smem_kernel<2><<<sblocks, threads>>>(out, siters); });
  printf("ATOMS float add, 2 lanes per address         : %.3f ms  %.2f lane-ops/clk/SM\n", ms, sops / (ms * 1e-3) / 148 / 1.9e9);
  ms = time_ms([&] { smem_kernel<3><<<sblocks, threads>>>(out, siters); });
  printf("ATOMS u64 min (CAS loop), conflict-free      : %.3f ms  %.2f lane-ops/clk/SM\n", ms, sops / (ms * 1e-3) / 148 / 1.9e9);
  return 0;
}
