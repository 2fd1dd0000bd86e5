// microbench_lsu.cu -- what does a partially-active memory instruction cost the LSU/L1 data pipe on B200?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/microbench_lsu.cu -o tools/microbench_lsu
// Each test: every warp of a full grid issues ITERS instructions of one kind with a fixed set of active lanes;
// reported: cycles per warp instruction per SM (lower bound of the pipe cost when many warps are resident).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// KIND 0: red.v4.f32, 1: red.f32, 2: ld.global.v4 (L1-resident table), 3: lds.128, 4: lds.32, 5: ld.global.f32 (L1-resident)
template <int KIND>
__global__ void __launch_bounds__(256) k(float* table, int rows, int iters, uint32_t lane_mask, int group, float* sink) {
  __shared__ float4 sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const bool on = (lane_mask >> lane) & 1u;
  float acc = 0.f;
  if (on) {
    for (int it = 0; it < iters; ++it) {
      // `group` consecutive lanes share one 64-B row (group 4 with v4: one full row; group 16 scalar: one full row)
      const uint32_t row = hash((warp * 32 + lane / group) * 9781u + it * 6151u) % rows;
      if (KIND == 0) {
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(table + (size_t)row * 16 + (lane % group) * 4 % 16), "f"(1.0f) : "memory");
      } else if (KIND == 1) {
        asm volatile("red.global.add.f32 [%0], %1;" :: "l"(table + (size_t)row * 16 + (lane % group) % 16), "f"(1.0f) : "memory");
      } else if (KIND == 2) {
        float4 v; const float* p = table + (size_t)(row % 256) * 16 + (lane % group) * 4 % 16;  // 16 KB hot set: L1 hits
        asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
        acc += v.x + v.w;
      } else if (KIND == 3) {
        const float4 v = sm[(row + (lane % group)) & 1023]; acc += v.x + v.w;
      } else if (KIND == 4) {
        acc += reinterpret_cast<float*>(sm)[(row * 4 + (lane % group)) & 4095];
      } else {
        float v; const float* p = table + (size_t)(row % 256) * 16 + (lane % group) % 16;
        asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
        acc += v;
      }
    }
  }
  if (acc == 12345.678f) sink[0] = acc;
}

template <int KIND>
void run(const char* name, float* table, int rows, uint32_t mask, int group, float* sink) {
  const int blocks = 148 * 8, threads = 256, iters = 512;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<KIND><<<blocks, threads>>>(table, rows, iters, mask, group, sink); cudaDeviceSynchronize();
  cudaEventRecord(a); k<KIND><<<blocks, threads>>>(table, rows, iters, mask, group, sink); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double winstr = (double)blocks * threads / 32 * iters;
  printf("%-34s lanes %2d group %2d : %.3f ms  %.2f cycles/warp-instr/SM @1.965GHz  (%.1f G instr/s)\n", name, __builtin_popcount(mask), group,
         ms, ms * 1e-3 * 1.965e9 * 148 / winstr, winstr / ms / 1e6);
}

int main() {
  const int rows = 50625 * 8; float* table; cudaMalloc(&table, sizeof(float) * 16 * rows); cudaMemset(table, 0, sizeof(float) * 16 * rows);
  float* sink; cudaMalloc(&sink, 64);
  const uint32_t M4 = 0xFu, M8 = 0xFFu, M16 = 0xFFFFu, M32 = 0xFFFFFFFFu, Mq = 0x0F0F0F0Fu /* 4 lanes in each quarter */, M4x2 = 0x000F000Fu;
  run<0>("red.v4  full warp (8 rows)", table, rows, M32, 4, sink);
  run<0>("red.v4  16 lanes (4 rows)", table, rows, M16, 4, sink);
  run<0>("red.v4  8 lanes (2 rows)", table, rows, M8, 4, sink);
  run<0>("red.v4  4 lanes (1 row)", table, rows, M4, 4, sink);
  run<0>("red.v4  4 lanes in each quarter", table, rows, Mq, 4, sink);
  run<0>("red.v4  2x4 lanes (2 quarters)", table, rows, M4x2, 4, sink);
  run<1>("red.f32 full warp (2 rows)", table, rows, M32, 16, sink);
  run<1>("red.f32 16 lanes (1 row)", table, rows, M16, 16, sink);
  run<2>("ldg.128 L1 hit full warp", table, rows, M32, 4, sink);
  run<2>("ldg.128 L1 hit 8 lanes", table, rows, M8, 4, sink);
  run<2>("ldg.128 L1 hit 4 lanes", table, rows, M4, 4, sink);
  run<2>("ldg.128 L1 hit 4 lanes same addr", table, rows, M4, 1, sink);
  run<2>("ldg.128 L1 hit full warp bcast/4", table, rows, M32, 1, sink);
  run<5>("ldg.32  L1 hit full warp", table, rows, M32, 16, sink);
  run<3>("lds.128 full warp distinct", table, rows, M32, 32, sink);
  run<3>("lds.128 full warp bcast per 4 lanes", table, rows, M32, 1, sink);
  run<3>("lds.128 8 lanes", table, rows, M8, 8, sink);
  run<4>("lds.32  full warp", table, rows, M32, 32, sink);
  return 0;
}
