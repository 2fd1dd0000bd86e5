// reduce.cu -- batch reduction of per-item gradients of parameters SHARED by all batch items.
//
// The reference has no distributed layer; a pipeline that renders N views of ONE mesh / attribute
// table gets its shared-parameter gradient as  sum_n grad[n, ...]  (autograd's expand backward).  On the multi-GPU
// path (batch sharded, one process per GPU) that local sum is what the ranks exchange, so it is written straight
// into the (registered / symmetric) communication bucket:
//
//   drtk_b200_batch_sum           out[m] = sum_n x[n * batch_stride + m]                 (one launch, 128-bit accesses)
//   drtk_b200_batch_sum_allreduce the same, then summed over all ranks IN THE SAME KERNEL through NVSwitch multicast
//                                 memory: every rank adds its local sums into every rank's bucket with
//                                 multimem.red.add.f32 (the switch fans the reduction out), followed by one flag
//                                 barrier over peer memory (double-buffered bucket).  No NCCL call on this path.
#include "common.cuh"

namespace drtk {
namespace {

template <bool VEC>
__global__ void __launch_bounds__(256) batch_sum_kernel(const float* __restrict__ x, int N, int64_t M, int64_t stride,
                                                        float* __restrict__ out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < M / 4; q += step) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int n = 0; n < N; ++n) {
        const float4 t = ldg_stream_f4(x + (int64_t)n * stride + q * 4);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      *reinterpret_cast<float4*>(out + q * 4) = acc;
    }
  } else {
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += step) {
      float acc = 0.f;
      for (int n = 0; n < N; ++n) acc += x[(int64_t)n * stride + m];
      out[m] = acc;
    }
  }
}

// ---- cross-GPU pieces (peer / multicast addresses come from the caller: torch symmetric memory) ----
__device__ __forceinline__ void multimem_red_add_v4(float* mc, float a, float b, float c, float d) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(mc), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void multimem_red_add(float* mc, float a) {
  asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" :: "l"(mc), "f"(a) : "memory");
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Barrier over all ranks, entered by every CTA of the (co-resident) grid: CTA b of rank r raises flag
// [r][b] on every peer (peer_flags[p] = base of rank p's flag array, world * gridDim.x words) to the epoch value and
// waits until its own array shows the epoch for (every rank, b).  Epochs grow monotonically (one per call and
// barrier), so flags are never reset.  flag_stride = words per rank in a flag array (the full-wave grid): a call may
// run on fewer CTAs than that (every rank the same number) and then simply leaves the upper slots of a row alone.
// A rank that never arrives (crashed peer) must not hang the GPU: after ~4e9 SM cycles (about two seconds) the
// wait gives up and raises *timeout_flag; the caller treats the bucket as invalid.
__device__ __forceinline__ void rank_barrier(uint32_t* const* peer_flags, int flag_stride, int rank, int world, uint32_t epoch,
                                             int* timeout_flag) {
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peer_flags[threadIdx.x] + (size_t)rank * flag_stride + blockIdx.x, epoch);
    const uint32_t* mine = peer_flags[rank] + (size_t)threadIdx.x * flag_stride + blockIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      __nanosleep(40);
      if (clock64() - t0 > 4000000000LL) {
        if (timeout_flag) atomicExch(timeout_flag, 1);
        break;
      }
    }
  }
  __syncthreads();
}

struct PeerFlags { uint32_t* p[16]; };

// One co-resident wave.  The bucket has TWO halves used alternately (`half` = 0 / 1, flipped by the caller once per
// backward pass): this call accumulates into half `half` and zero-fills this rank's copy of the OTHER half for the next
// pass, so ONE barrier per call suffices:
//   * every rank's copy of half `half` was zeroed by that rank's previous call, which ended with a barrier;
//   * local batch sum, multimem.red.add into ALL copies of half `half`; zero the local copy of the other half;
//   * barrier: every copy of half `half` now holds the sum over ranks, and nobody will add into the half just zeroed
//     before the next call (whose reductions start after the peers passed this barrier).
// The caller consumes half `half` (stream order) before its next call on the SAME half, i.e. two passes later.
template <bool VEC>
__global__ void __launch_bounds__(256) batch_sum_allreduce_kernel(const float* __restrict__ x, int N, int64_t M, int64_t stride,
                                                                  float* __restrict__ zero_local, float* acc_mc,
                                                                  PeerFlags flags, int flag_stride, int rank, int world,
                                                                  uint32_t epoch, int* timeout_flag) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (VEC) {
    for (int64_t q = t0; q < M / 4; q += step) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int n = 0; n < N; ++n) {
        const float4 t = ldg_stream_f4(x + (int64_t)n * stride + q * 4);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      multimem_red_add_v4(acc_mc + q * 4, acc.x, acc.y, acc.z, acc.w);
      *reinterpret_cast<float4*>(zero_local + q * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int64_t m = t0; m < M; m += step) {
      float acc = 0.f;
      for (int n = 0; n < N; ++n) acc += x[(int64_t)n * stride + m];
      multimem_red_add(acc_mc + m, acc);
      zero_local[m] = 0.f;
    }
  }
  rank_barrier(flags.p, flag_stride, rank, world, epoch + 1, timeout_flag);
}

inline bool vec_ok(const float* x, int64_t M, int64_t stride, const float* out) {
  return (M % 4 == 0) && (stride % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
         (reinterpret_cast<uintptr_t>(out) % 16 == 0);
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" int drtk_b200_batch_sum(const float* x, int64_t N, int64_t M, int64_t batch_stride, float* out, void* stream_) {
  if (N < 0 || M < 0) return DRTK_B200_EINVAL;
  if (M == 0) return 0;
  if (!out || (N > 0 && !x)) return DRTK_B200_EINVAL;
  if (N > (1 << 30)) return DRTK_B200_EUNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = vec_ok(x, M, batch_stride, out);
  const int64_t items = vec ? M / 4 : M;
  const int64_t need = (items + 255) / 256, cap = (int64_t)num_sms() * 8;
  const unsigned grid = (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
  if (vec) batch_sum_kernel<true><<<grid, 256, 0, stream>>>(x, (int)N, M, batch_stride, out);
  else batch_sum_kernel<false><<<grid, 256, 0, stream>>>(x, (int)N, M, batch_stride, out);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_batch_sum_allreduce_grid(void) {
  // CTAs of the fused kernel: ONE co-resident wave (the in-kernel barriers need every CTA running), fixed so that
  // the caller can size the flag arrays: world * grid words per rank
  return num_sms();
}

extern "C" int drtk_b200_batch_sum_allreduce(const float* x, int64_t N, int64_t M, int64_t batch_stride,
                                             float* zero_local, float* acc_multicast, void* const* peer_flags,
                                             int rank, int world, uint32_t epoch, int* timeout_flag, int max_ctas,
                                             void* stream_) {
  if (N < 0 || M < 0 || world < 1 || world > 16 || rank < 0 || rank >= world || max_ctas < 0) return DRTK_B200_EINVAL;
  if (M == 0) return 0;
  if (!zero_local || !acc_multicast || !peer_flags || (N > 0 && !x)) return DRTK_B200_EINVAL;
  float* bucket_local = zero_local;
  float* bucket_multicast = acc_multicast;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PeerFlags fl;
  for (int i = 0; i < 16; ++i) fl.p[i] = i < world ? static_cast<uint32_t*>(peer_flags[i]) : nullptr;
  const bool vec = vec_ok(x, M, batch_stride, bucket_local) && (reinterpret_cast<uintptr_t>(bucket_multicast) % 16 == 0);
  // max_ctas: 0 = the full wave (an exchange on the critical path); a smaller grid for an exchange that overlaps
  // other kernels -- its CTAs wait at the rank barrier for the slowest rank and should not hold every SM meanwhile
  const int wave = num_sms();
  const unsigned grid = (unsigned)((max_ctas > 0 && max_ctas < wave) ? max_ctas : wave);
  if (vec) batch_sum_allreduce_kernel<true><<<grid, 256, 0, stream>>>(x, (int)N, M, batch_stride, bucket_local, bucket_multicast, fl, wave, rank, world, epoch, timeout_flag);
  else batch_sum_allreduce_kernel<false><<<grid, 256, 0, stream>>>(x, (int)N, M, batch_stride, bucket_local, bucket_multicast, fl, wave, rank, world, epoch, timeout_flag);
  DRTK_CHECK_LAUNCH();
  return 0;
}
