// interpolate.cu -- barycentric interpolation of vertex attributes, forward and backward.
//
// Semantics: src/interpolate/interpolate_kernel.cu:38-111 (forward) and :113-299 (backward) of
// the reference; fp32 results within 1e-5 relative.
//
// Forward  (16 + 4C B/px, write dominated): one thread owns four adjacent pixels and walks the
// channels four at a time: per 4-channel group 3 x LDG.128 per distinct triangle (attribute rows
// are 16-B aligned when C % 4 == 0) and 4 x STG.128 into the planar output, so every warp store
// is a full 512-B span of one channel plane.
//
// Backward (16 + 4C B/px read, +12 B/px written when bary_img needs grad):
//   bary grad  : direct per-pixel dot products.
//   vertex grad: lanes of a warp are consecutive pixels; runs of equal (image, triangle) are
//                reduced with a segmented shuffle scan, 4 channels x 3 vertices at a time, and the
//                head lane of each run issues one 128-bit vector reduction per vertex
//                (red.global.add.v4.f32) -- no shared memory, no block barriers (the reference
//                needs one barrier and up to three scalar atomics per channel per run).
#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace drtk {
namespace {

struct InterpArgs {
  const float* attr;
  Strides3 as;
  const int32_t* vi;
  Strides3 vis;
  const int32_t* index_img;
  Strides3 is;
  const float* bary;
  Strides4 bs;
  int N, V, F, C, H, W;
};

__device__ __forceinline__ void load_vi(const InterpArgs& a, int n, int t, int& i0, int& i1, int& i2) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)t * a.vis.s1;
  i0 = vip[0]; i1 = vip[a.vis.s2]; i2 = vip[2 * a.vis.s2];
}

__device__ __forceinline__ float sweep_x(int w, int W) { return ((float)w * 2.0f + 1.0f) / (float)W - 1.0f; }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// VEC: index/bary images dense & 16-B aligned along W (4 px per thread).
// AVEC: attribute rows 16-B aligned and C % 4 == 0 (float4 gathers).
template <bool VEC, bool AVEC>
__global__ void __launch_bounds__(256) interp_fwd_kernel(InterpArgs a, float* __restrict__ out) {
  const int HW = a.H * a.W;  // blockIdx.y = image, 32-bit pixel arithmetic inside it
  constexpr int PX = VEC ? 4 : 1;
  const int n = blockIdx.y;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HW / PX; q += gridDim.x * blockDim.x) {
    const int rem = q * PX;
    const int h = rem / a.W, w = rem - h * a.W;
    int ids[PX];
    float b0[PX], b1[PX], b2[PX];
    const int32_t* ip = a.index_img + (int64_t)n * a.is.s0 + (int64_t)h * a.is.s1 + (int64_t)w * a.is.s2;
    const float* bp = a.bary + (int64_t)n * a.bs.s0 + (int64_t)h * a.bs.s2 + (int64_t)w * a.bs.s3;
    if (VEC) {
      const int4 id4 = ldg_stream_i4(ip);
      ids[0] = id4.x; ids[PX > 1 ? 1 : 0] = id4.y; ids[PX > 2 ? 2 : 0] = id4.z; ids[PX > 3 ? 3 : 0] = id4.w;
      const float4 x0 = ldg_stream_f4(bp), x1 = ldg_stream_f4(bp + a.bs.s1), x2 = ldg_stream_f4(bp + 2 * a.bs.s1);
      b0[0] = x0.x; b0[PX > 1 ? 1 : 0] = x0.y; b0[PX > 2 ? 2 : 0] = x0.z; b0[PX > 3 ? 3 : 0] = x0.w;
      b1[0] = x1.x; b1[PX > 1 ? 1 : 0] = x1.y; b1[PX > 2 ? 2 : 0] = x1.z; b1[PX > 3 ? 3 : 0] = x1.w;
      b2[0] = x2.x; b2[PX > 1 ? 1 : 0] = x2.y; b2[PX > 2 ? 2 : 0] = x2.z; b2[PX > 3 ? 3 : 0] = x2.w;
    } else {
      ids[0] = ip[0];
      b0[0] = bp[0]; b1[0] = bp[a.bs.s1]; b2[0] = bp[2 * a.bs.s1];
    }
    int i0[PX], i1[PX], i2[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      i0[j] = i1[j] = i2[j] = 0;
      if (ids[j] != -1) {
        if (j > 0 && ids[j] == ids[j - 1]) { i0[j] = i0[j - 1]; i1[j] = i1[j - 1]; i2[j] = i2[j - 1]; }
        else load_vi(a, n, ids[j], i0[j], i1[j], i2[j]);
      }
    }
    const float sy = sweep_x(h, a.H);
    const float* an = a.attr + (int64_t)n * a.as.s0;
    float* op = out + (int64_t)n * a.C * HW + rem;  // 64-bit: n*C*HW may exceed 2^31

    if (AVEC) {
      // row offsets in 32 bits (host guarantees V * row_stride < 2^31 on this path)
      const unsigned rs = (unsigned)a.as.s1;
#pragma unroll
      for (int j = 0; j < PX; ++j) { i0[j] *= rs; i1[j] *= rs; i2[j] *= rs; }
      for (int c = 0; c < a.C; c += 4) {
        float r[4][PX];  // [channel][pixel]
        float4 A0, A1, A2;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          if (ids[j] != -1) {
            if (!(j > 0 && ids[j] == ids[j - 1])) {
              A0 = *reinterpret_cast<const float4*>(an + (unsigned)(i0[j] + c));
              A1 = *reinterpret_cast<const float4*>(an + (unsigned)(i1[j] + c));
              A2 = *reinterpret_cast<const float4*>(an + (unsigned)(i2[j] + c));
            }
            r[0][j] = A0.x * b0[j] + A1.x * b1[j] + A2.x * b2[j];  // (:102)
            r[1][j] = A0.y * b0[j] + A1.y * b1[j] + A2.y * b2[j];
            r[2][j] = A0.z * b0[j] + A1.z * b1[j] + A2.z * b2[j];
            r[3][j] = A0.w * b0[j] + A1.w * b1[j] + A2.w * b2[j];
          } else {  // coordinate sweep for empty pixels (:104-109): even channel -> x, odd -> y
            const float sx = sweep_x(w + j, a.W);
            r[0][j] = sx; r[1][j] = sy; r[2][j] = sx; r[3][j] = sy;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* o = op + (int64_t)(c + k) * HW;
          if (VEC) stg_stream_f4(o, make_float4(r[k][0], r[k][PX > 1 ? 1 : 0], r[k][PX > 2 ? 2 : 0], r[k][PX > 3 ? 3 : 0]));
          else o[0] = r[k][0];
        }
      }
    } else {
      for (int c = 0; c < a.C; ++c) {
        float r[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          if (ids[j] != -1) {
            const float v0 = an[(int64_t)i0[j] * a.as.s1 + (int64_t)c * a.as.s2];
            const float v1 = an[(int64_t)i1[j] * a.as.s1 + (int64_t)c * a.as.s2];
            const float v2 = an[(int64_t)i2[j] * a.as.s1 + (int64_t)c * a.as.s2];
            r[j] = v0 * b0[j] + v1 * b1[j] + v2 * b2[j];
          } else {
            r[j] = (c & 1) ? sy : sweep_x(w + j, a.W);
          }
        }
        float* o = op + (int64_t)c * HW;
        if (VEC) stg_stream_f4(o, make_float4(r[0], r[PX > 1 ? 1 : 0], r[PX > 2 ? 2 : 0], r[PX > 3 ? 3 : 0]));
        else o[0] = r[0];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct InterpBwdArgs {
  InterpArgs f;
  const float* grad_out;
  Strides4 gs;
};

// One thread per pixel.  NEED_VERT: accumulate vertex-attribute gradients; NEED_BARY: write
// the barycentric gradient image.  RV4: vert_grad rows are 16-B aligned (C % 4 == 0) so run
// heads can use 128-bit vector reductions.
template <bool NEED_VERT, bool NEED_BARY, bool RV4>
__global__ void __launch_bounds__(256) interp_bwd_kernel(InterpBwdArgs b, float* __restrict__ vert_grad,
                                                         float* __restrict__ bary_grad) {
  const InterpArgs& a = b.f;
  const int64_t HW = (int64_t)a.H * a.W;  // blockIdx.y = image
  const int lane = threadIdx.x & 31;
  const int rem = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = rem < (int)HW;
  const int n = blockIdx.y;
  int h = 0, w = 0, id = -1;
  if (in_range) {
    h = rem / a.W; w = rem - h * a.W;
    id = a.index_img[(int64_t)n * a.is.s0 + (int64_t)h * a.is.s1 + (int64_t)w * a.is.s2];
  }
  const bool used = id != -1;
  const bool warp_used = __any_sync(0xffffffffu, used);
  if (!warp_used) {
    if (NEED_BARY && in_range) {  // every pixel of bary_grad is written (:282-297)
      float* g = bary_grad + (int64_t)n * 3 * HW + rem;
      g[0] = 0.f; g[HW] = 0.f; g[2 * HW] = 0.f;
    }
    return;
  }
  int i0 = 0, i1 = 0, i2 = 0;
  float b0 = 0.f, b1 = 0.f, b2 = 0.f;
  if (used) {
    load_vi(a, n, id, i0, i1, i2);
    if (NEED_VERT) {
      const float* bp = a.bary + (int64_t)n * a.bs.s0 + (int64_t)h * a.bs.s2 + (int64_t)w * a.bs.s3;
      b0 = ldg_stream_f(bp); b1 = ldg_stream_f(bp + a.bs.s1); b2 = ldg_stream_f(bp + 2 * a.bs.s1);
    }
  }
  // runs of equal (image, triangle) along the warp
  const int key = used ? id : (-2 - lane);  // a block never spans two images
  const int key_up = __shfl_up_sync(0xffffffffu, key, 1);
  const int key_dn = __shfl_down_sync(0xffffffffu, key, 1);
  const bool head = (lane == 0) || (key_up != key);
  const bool tail = (lane == 31) || (key_dn != key);
  const unsigned tail_mask = __ballot_sync(0xffffffffu, tail);

  const float* gp = b.grad_out + (int64_t)n * b.gs.s0 + (int64_t)h * b.gs.s2 + (int64_t)w * b.gs.s3;
  const float* an = a.attr + (int64_t)n * a.as.s0;
  const float* a0 = an + (int64_t)i0 * a.as.s1;
  const float* a1 = an + (int64_t)i1 * a.as.s1;
  const float* a2 = an + (int64_t)i2 * a.as.s1;
  float* vg = NEED_VERT ? vert_grad + (int64_t)n * a.V * a.C : nullptr;
  float gb0 = 0.f, gb1 = 0.f, gb2 = 0.f;

  for (int c = 0; c < a.C; c += 4) {
    float g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      g[k] = (in_range && c + k < a.C) ? ldg_stream_f(gp + (int64_t)(c + k) * b.gs.s1) : 0.f;
    if (NEED_BARY && used) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (c + k < a.C) {
          gb0 += g[k] * a0[(int64_t)(c + k) * a.as.s2];  // (:256-258)
          gb1 += g[k] * a1[(int64_t)(c + k) * a.as.s2];
          gb2 += g[k] * a2[(int64_t)(c + k) * a.as.s2];
        }
      }
    }
    if (NEED_VERT) {
      float s[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) { s[k] = g[k] * b0; s[4 + k] = g[k] * b1; s[8 + k] = g[k] * b2; }  // (:262-267)
      seg_reduce_to_head<12>(s, tail_mask, lane);
      if (head && used) {
        float* r0 = vg + (int64_t)i0 * a.C + c;
        float* r1 = vg + (int64_t)i1 * a.C + c;
        float* r2 = vg + (int64_t)i2 * a.C + c;
        if (RV4) {
          red_add_v4(r0, s[0], s[1], s[2], s[3]);
          red_add_v4(r1, s[4], s[5], s[6], s[7]);
          red_add_v4(r2, s[8], s[9], s[10], s[11]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < a.C) { red_add(r0 + k, s[k]); red_add(r1 + k, s[4 + k]); red_add(r2 + k, s[8 + k]); }
        }
      }
    }
  }
  if (NEED_BARY && in_range) {
    float* g = bary_grad + (int64_t)n * 3 * HW + rem;
    g[0] = gb0; g[HW] = gb1; g[2 * HW] = gb2;  // zeros for empty pixels
  }
}


// ------------------------------------------------------------------------------------------
// backward, tiled (the fast path): bulk-async staged tiles + register run-reduction
// ------------------------------------------------------------------------------------------
// One persistent CTA of 512 threads per SM walks "tiles" of TP consecutive pixels of the flattened
// H*W plane of one image (planes are dense, so a tile is ONE contiguous 4*TP-byte span per plane).
// A two-stage shared-memory ring is fed by cp.async.bulk copies (the TMA engine, SASS UBLKCP:
// one copy per plane, completing on an mbarrier); the copies of tile i+1 are in flight while tile i
// is reduced, and no LSU instruction or register is spent on the 16+4C B/px input stream.
//   phase A (vertex grads): "walkers" of LPW lanes -- lane = channel -- walk a pixel segment in
//     4-pixel steps (one conflict-free LDS.128 of their own channel plane, broadcast LDS.128 of
//     index/bary), accumulate g*bary_k of the current triangle run in three registers and, when the
//     triangle id changes, flush the run with three reductions whose LPW lanes hit LPW consecutive
//     floats of one vertex row (one coalesced 64-B RED per vertex at C=16; measured 69 G rows/s).
//     No shuffles, no shared-memory atomics, one reduction per (run, vertex) instead of per pixel.
//     A run is "consecutive pixels showing the same triangle", so crossing an image-row boundary
//     inside a tile is harmless: sums are per triangle.
//   phase B (bary grads): thread = 2 (or 4) consecutive pixels; dot products of the staged gradients
//     with the three attribute rows (LDG.128 row gathers, shared by neighbouring pixels of a triangle).
constexpr int kBwdThreads = 512;               // consumer threads (16 warps)
constexpr int kBwdProducers = 4;                // producer warps: a bulk copy costs its issuing thread ~0.145 us
                                                // whatever its size (profiles/r01_microbench_bulk_copy.txt), so the
                                                // ~20 copies of a tile are spread over 4 issuing warps
constexpr int kBwdBlock = kBwdThreads + 32 * kBwdProducers;
constexpr int kBwdStages = 2;

template <int LPW> struct BwdTileCfg { static constexpr int TP = (LPW == 16) ? 1024 : 2048; };

template <int LPW>
struct BwdStage {
  static constexpr int TP = BwdTileCfg<LPW>::TP;
  static constexpr int PITCH = TP + 4;  // plane pitch == 4 (mod 32) words: LDS.128 of LPW consecutive
                                        // planes by LPW lanes is bank-conflict free
  float g[LPW * PITCH];
  float bary[3 * TP];
  int idx[TP];
};

template <int LPW>
struct BwdTileSmem {
  BwdStage<LPW> st[kBwdStages];
  unsigned long long full[kBwdStages];   // producer -> consumers: tile landed (transaction barrier)
  unsigned long long empty[kBwdStages];  // consumers -> producer: every warp is done reading the stage
};

// Phase A of the tiled backward for one walker lane (see the kernel comment).  Kept out of line so
// that its loop gets its own register allocation: the run-boundary block must stay short (it
// executes once per ~5 pixels), which needs the table pointers and strides resident in registers.
//   gp: this lane's channel plane; ip/bp: index and bary planes of the stage (bary plane pitch TP)
//   vg: vertex-gradient table of this image, already offset by this lane's channel
//   vib: vi rows of this image; vs1/vs2 element strides of vi (row, corner)
// NOTE: it must be entered from CONVERGENT control flow with a warp-uniform trip count: uniform-datapath
// instructions cannot be issued from divergent code, and a divergent entry makes the compiler keep the
// global-memory descriptor in vector registers and R2UR it before every LDG/REDG (measured: 14 % of all
// executed instructions).  Tail tiles are therefore padded with index -1 instead of shortening the loop.
template <int TP, int SEG>
__device__ __noinline__ void walk_runs(const float* __restrict__ gp, const int* __restrict__ ip,
                                       const float* __restrict__ bp, int xs, float* vg,
                                       const int32_t* __restrict__ vib, int vs1, int vs2, unsigned Cs,
                                       bool c_on) {
  int cur = -1;
  unsigned v0 = 0, v1 = 0, v2 = 0;  // vertex ids of the current run
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 1
  for (int x = xs; x < xs + SEG; x += 4) {
    const float4 gq = *reinterpret_cast<const float4*>(gp + x);
    const int4 iq = *reinterpret_cast<const int4*>(ip + x);
    const float4 p0q = *reinterpret_cast<const float4*>(bp + x);
    const float4 p1q = *reinterpret_cast<const float4*>(bp + TP + x);
    const float4 p2q = *reinterpret_cast<const float4*>(bp + 2 * TP + x);
    const int ids[4] = {iq.x, iq.y, iq.z, iq.w};
    const float gs[4] = {gq.x, gq.y, gq.z, gq.w};
    const float q0[4] = {p0q.x, p0q.y, p0q.z, p0q.w};
    const float q1[4] = {p1q.x, p1q.y, p1q.z, p1q.w};
    const float q2[4] = {p2q.x, p2q.y, p2q.z, p2q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int id = ids[j];
      if (id != cur) {  // run boundary (uniform across the walker's lanes)
        if (cur >= 0 && c_on) {
          red_add(vg + (size_t)(v0 * Cs), a0);
          red_add(vg + (size_t)(v1 * Cs), a1);
          red_add(vg + (size_t)(v2 * Cs), a2);
        }
        // vertex ids of the new run: fetched now, consumed at its flush.  Empty pixels (id < 0)
        // read triangle 0 harmlessly; their run is never flushed.
        const int32_t* vip = vib + (size_t)((unsigned)max(id, 0) * (unsigned)vs1);
        v0 = (unsigned)vip[0];
        v1 = (unsigned)vip[vs2];
        v2 = (unsigned)vip[2 * vs2];
        a0 = a1 = a2 = 0.f;
        cur = id;
      }
      // unconditional: while cur < 0 the sums are garbage that is reset before use
      a0 = fmaf(gs[j], q0[j], a0);
      a1 = fmaf(gs[j], q1[j], a1);
      a2 = fmaf(gs[j], q2[j], a2);
    }
  }
  if (cur >= 0 && c_on) {
    red_add(vg + (size_t)(v0 * Cs), a0);
    red_add(vg + (size_t)(v1 * Cs), a1);
    red_add(vg + (size_t)(v2 * Cs), a2);
  }
}

template <int LPW, bool NEED_VERT, bool NEED_BARY, bool AVEC>
__global__ void __launch_bounds__(kBwdBlock, 1)
interp_bwd_tile_kernel(InterpBwdArgs b, float* __restrict__ vert_grad, float* __restrict__ bary_grad,
                       int tiles_per_img, int64_t num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdTileSmem<LPW>& S = *reinterpret_cast<BwdTileSmem<LPW>*>(smem_raw);
  constexpr int TP = BwdTileCfg<LPW>::TP;
  constexpr int PITCH = BwdStage<LPW>::PITCH;
  const InterpArgs& a = b.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int HW = a.H * a.W;  // < 2^31 guaranteed by the host
  if (tid == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), kBwdProducers);
      mbar_init(reinterpret_cast<uint64_t*>(&S.empty[s]), kBwdThreads / 32);  // one arrival per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int nchunks = (a.C + LPW - 1) / LPW;
  // pipeline items: (tile, channel pass); this CTA owns tiles blockIdx.x, +gridDim.x, ...
  const int64_t my_tiles = (num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_items = my_tiles * nchunks;

  // Executed by one lane of a dedicated producer warp: bulk copies take warp-uniform operands (UBLKCP).
  // The warp role is made warp-uniform with a shuffle broadcast (the CUTLASS canonical_warp_idx idiom)
  // so that the role branch is uniform and the consumers' code stays in CONVERGENT control flow: behind
  // a thread-dependent branch, or with UBLKCP code inside the consumers' loop, the compiler keeps the
  // global-memory descriptor in vector registers and R2URs it before every LDG/REDG (measured: 14 %
  // of all executed instructions).
  auto issue = [&](int64_t item, int pw) {  // producer warp pw issues copies pw, pw + kBwdProducers, ...
    const int s = (int)(item & 1);
    const int64_t tile = blockIdx.x + (item / nchunks) * gridDim.x;
    const int chunk = (int)(item % nchunks);
    const int n = (int)(tile / tiles_per_img);
    const int p0 = (int)(tile - (int64_t)n * tiles_per_img) * TP;
    const int npx = min(TP, HW - p0);
    const int c0 = chunk * LPW, nc = min(LPW, a.C - c0);
    const int ncopies = nc + (NEED_VERT ? 3 : 0) + 1;
    uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[s]);
    BwdStage<LPW>& st = S.st[s];
    const uint32_t bytes = (uint32_t)npx * 4u;
    const int mine = (ncopies - pw + kBwdProducers - 1) / kBwdProducers;
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(bar, (uint32_t)mine * bytes);  // one arrival per producer warp, also when mine == 0
    for (int k = pw; k < ncopies; k += kBwdProducers) {
      const void* src;
      void* dst;
      if (k < nc) {
        src = b.grad_out + (int64_t)n * b.gs.s0 + (int64_t)(c0 + k) * b.gs.s1 + p0;
        dst = st.g + k * PITCH;
      } else if (NEED_VERT && k < nc + 3) {
        src = a.bary + (int64_t)n * a.bs.s0 + (int64_t)(k - nc) * a.bs.s1 + p0;
        dst = st.bary + (k - nc) * TP;
      } else {
        src = a.index_img + (int64_t)n * a.is.s0 + p0;
        dst = st.idx;
      }
      bulk_g2s(dst, src, bytes, bar);
    }
  };

  const int warp_role = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  if (warp_role >= kBwdThreads / 32) {  // ---- producer warps ----
    if (lane == 0) {
      const int pw = warp_role - kBwdThreads / 32;
      for (int64_t item = 0; item < n_items; ++item) {
        if (item >= kBwdStages)  // wait until the consumers released this stage (its previous use)
          mbar_wait_backoff(reinterpret_cast<uint64_t*>(&S.empty[item & 1]), (uint32_t)((item / kBwdStages - 1) & 1), 400);
        issue(item, pw);
      }
    }
    return;
  }
  uint32_t phase_bits = 0;  // bit s = parity to wait for on full[s]

  // phase-B state of the current tile (a thread owns the same PPT pixels across channel passes)
  constexpr int PPT = TP / kBwdThreads;  // 2 (TP 1024) or 4 (TP 2048) consecutive pixels per thread
  float gb[PPT][3];

  for (int64_t item = 0; item < n_items; ++item) {
    const int s = (int)(item & 1);
    const int64_t tile = blockIdx.x + (item / nchunks) * gridDim.x;
    const int chunk = (int)(item % nchunks);
    const int n = (int)(tile / tiles_per_img);
    const int p0 = (int)(tile - (int64_t)n * tiles_per_img) * TP;
    const int npx = min(TP, HW - p0);
    const int c0 = chunk * LPW, nc = min(LPW, a.C - c0);
    BwdStage<LPW>& st = S.st[s];
    mbar_wait_backoff(reinterpret_cast<uint64_t*>(&S.full[s]), (phase_bits >> s) & 1u, 40);
    phase_bits ^= (1u << s);
    if (npx < TP) {  // last tile of an image (warp-uniform): pad with "no triangle" so loops keep full length
      for (int i = npx + tid; i < TP; i += kBwdThreads) st.idx[i] = -1;
      asm volatile("bar.sync 1, %0;" :: "n"(kBwdThreads) : "memory");  // consumers only (producer warp excluded)
    }

    // ---- phase A: vertex-attribute gradients ----
    if (NEED_VERT) {
      constexpr int WALKERS = kBwdThreads / LPW;
      constexpr int SEG = TP / WALKERS;  // pixels per walker (32 / 32 / 16 for LPW 16 / 8 / 4)
      const int walker = tid / LPW, c = tid - walker * LPW;
      const bool c_on = c < nc;
      walk_runs<TP, SEG>(st.g + (c_on ? c : 0) * PITCH, st.idx, st.bary, walker * SEG,
                         vert_grad + (int64_t)n * a.V * a.C + c0 + (c_on ? c : 0),
                         a.vi + (int64_t)n * a.vis.s0, (int)a.vis.s1, (int)a.vis.s2, (unsigned)a.C, c_on);
    }

    // ---- phase B: barycentric gradients: thread = PPT consecutive pixels, all channels ----
    if (NEED_BARY) {
      const int x = tid * PPT;
      if (chunk == 0) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) gb[j][0] = gb[j][1] = gb[j][2] = 0.f;
      }
      {
        int qid[PPT], qv[PPT][3];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          qid[j] = st.idx[x + j];
          qv[j][0] = qv[j][1] = qv[j][2] = 0;
          if (qid[j] >= 0) {
            if (j > 0 && qid[j] == qid[j - 1]) {
              qv[j][0] = qv[j - 1][0]; qv[j][1] = qv[j - 1][1]; qv[j][2] = qv[j - 1][2];
            } else {
              load_vi(a, n, qid[j], qv[j][0], qv[j][1], qv[j][2]);
            }
          }
        }
        const float* an = a.attr + (int64_t)n * a.as.s0 + (int64_t)c0 * a.as.s2;
        const float* r0[PPT], *r1[PPT], *r2[PPT];  // attribute rows of each pixel's three vertices
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          r0[j] = an + (int64_t)qv[j][0] * a.as.s1;
          r1[j] = an + (int64_t)qv[j][1] * a.as.s1;
          r2[j] = an + (int64_t)qv[j][2] * a.as.s1;
        }
        const float* gq_base = st.g + x;
        if (AVEC) {
#pragma unroll(PPT == 2 ? 2 : 1)  // two channel groups in flight: their LDG.128 (L2 hits) overlap
          for (int cc = 0; cc < nc; cc += 4) {
            float gq[4][PPT];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int j = 0; j < PPT; ++j) gq[k][j] = gq_base[(cc + k) * PITCH + j];
            }
            float4 A0, A1, A2;
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              // rows are (re)loaded only when the triangle changes; empty pixels (qid < 0) use the rows of
              // vertex 0 and accumulate into a slot that is zeroed at the store
              if (j == 0 || qid[j] != qid[j - 1]) {
                A0 = *reinterpret_cast<const float4*>(r0[j] + cc);
                A1 = *reinterpret_cast<const float4*>(r1[j] + cc);
                A2 = *reinterpret_cast<const float4*>(r2[j] + cc);
              }
              gb[j][0] += gq[0][j] * A0.x + gq[1][j] * A0.y + gq[2][j] * A0.z + gq[3][j] * A0.w;
              gb[j][1] += gq[0][j] * A1.x + gq[1][j] * A1.y + gq[2][j] * A1.z + gq[3][j] * A1.w;
              gb[j][2] += gq[0][j] * A2.x + gq[1][j] * A2.y + gq[2][j] * A2.z + gq[3][j] * A2.w;
            }
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < nc; ++cc) {
            float A0, A1, A2;
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              if (j == 0 || qid[j] != qid[j - 1]) {
                A0 = r0[j][(int64_t)cc * a.as.s2];
                A1 = r1[j][(int64_t)cc * a.as.s2];
                A2 = r2[j][(int64_t)cc * a.as.s2];
              }
              const float g = gq_base[cc * PITCH + j];
              gb[j][0] += g * A0; gb[j][1] += g * A1; gb[j][2] += g * A2;
            }
          }
        }
        if (chunk == nchunks - 1 && x < npx) {
          float* gp = bary_grad + (int64_t)n * 3 * HW + p0 + x;
#pragma unroll
          for (int j = 0; j < PPT; ++j)
            if (qid[j] < 0) gb[j][0] = gb[j][1] = gb[j][2] = 0.f;  // (:282-297) zeros where empty
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            if (PPT == 2) *reinterpret_cast<float2*>(gp + (int64_t)k * HW) = make_float2(gb[0][k], gb[PPT - 1][k]);
            else *reinterpret_cast<float4*>(gp + (int64_t)k * HW) = make_float4(gb[0][k], gb[1 % PPT][k], gb[2 % PPT][k], gb[3 % PPT][k]);
          }
        }
      }
    }
    __syncwarp();  // this warp is done reading stage s: let the producer refill it
    if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.empty[s]));
  }
}


// ------------------------------------------------------------------------------------------
// backward, tiled, v5 (C % 4 == 0): quad-lane walkers + packed triangle table + A/B warp teams
// ------------------------------------------------------------------------------------------
// What the ncu capture of the kernel above showed (profiles/r01_ncu_interp_bwd_v4_vs_v5.md): 20 warp
// instructions per pixel -- 10.4 in the walkers (each lane carries ONE channel, so the per-pixel run test
// and the 27-instruction run-boundary block are paid per channel, and the two walkers of a warp diverge at
// their boundaries), 5.8 in phase B, 2.6 in per-tile bookkeeping (64-bit divisions) -- and the LSU data pipe
// at 67 % (shared-memory wavefronts of the broadcast index/bary loads, three scalar vi loads per run).  Here:
//   * a walker is FOUR lanes, each carrying four channels (12 accumulators): the run test, the boundary
//     block and the broadcast index/bary loads are amortised over 4x the channels; a run is flushed with
//     three red.global.add.v4.f32 per lane (the four lanes of a walker cover one 64-B vertex row);
//   * triangle -> vertex ids come from a packed int4 table (one LDG.128 instead of three strided LDG.32
//     plus their address arithmetic), built per call by a tiny kernel into a stream-ordered allocation;
//   * a warp is 8 walkers x 16-pixel segments = 128 pixels.  The consumer warps of a CTA form two teams
//     that swap roles every tile: one team walks (phase A), the other computes the barycentric gradients
//     of the same tile (phase B, 2 x 64 pixels per warp);
//   * tiles are 512 pixels and TWO CTAs of 8 consumer + 2 producer warps share an SM: two independent
//     two-stage pipelines, so that one CTA computes while the other sits at its tile hand-over;
//   * channel planes are staged with a 16-B skew per group of four planes, which makes the walkers'
//     LDS.128 (lane = plane group, two walkers 64 B apart per quarter warp) bank-conflict free;
//   * tile bookkeeping is 32-bit and incremental.
constexpr int kQSeg = 16;       // pixels per walker
constexpr int kQCh = 16;        // channels per pass over a tile
constexpr int kQUnit = 128;     // pixels per warp: 8 walkers x kQSeg (phase A) = 2 x 32 lanes x 2 px (phase B)
constexpr int kQProducers = 2;  // 20 bulk copies per tile at ~0.145 us each per issuing thread
constexpr int kQChunk = 8;      // tiles per scheduling chunk (see the tile order comment in the kernel)
// tile pixels / pipeline stages per CTA / co-resident CTAs per SM of the quad kernel (A/B knobs, tools/build_variants.sh)
#ifndef DRTK_INTERP_BWD_TP
#define DRTK_INTERP_BWD_TP 512
#endif
#ifndef DRTK_INTERP_BWD_STAGES
#define DRTK_INTERP_BWD_STAGES 2
#endif
#ifndef DRTK_INTERP_BWD_CTAS
#define DRTK_INTERP_BWD_CTAS 2
#endif
constexpr int kQStages = DRTK_INTERP_BWD_STAGES;
constexpr int kQCtas = DRTK_INTERP_BWD_CTAS;

template <int TP>
struct QStage {
  float g[kQCh * TP + 16];  // plane c starts at c * TP + (c >> 2) * 4
  float bary[3 * TP];
  int idx[TP];
};
template <int TP>
struct QSmem {
  QStage<TP> st[kQStages];
  unsigned long long full[kQStages];
  unsigned long long empty[kQStages];
};

// Also zero-fills the vertex-gradient table the walkers reduce into (saves a memset launch per step).
__global__ void __launch_bounds__(256) vi_table_kernel(const int32_t* __restrict__ vi, Strides3 vis, int F,
                                                       int64_t total, int4* __restrict__ tab,
                                                       float4* __restrict__ zero, int64_t zero_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t j = i; j < zero_count; j += (int64_t)gridDim.x * blockDim.x) zero[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i >= total) return;
  const int64_t n = i / F, f = i - n * F;
  const int32_t* vip = vi + n * vis.s0 + f * vis.s1;
  tab[i] = make_int4(vip[0], vip[vis.s2], vip[2 * vis.s2], 0);
}

// Phase A of one walker lane: 16 pixels, four channels.  gp: plane 4q of the stage at the segment start
// (the next three planes follow at TP); ip/bp: index / bary planes at the segment start; vg: vertex
// gradient table of the image offset by this lane's first channel; tab: packed triangle table of the image;
// off_mask: 0 for a lane that carries channels, -1 for an idle lane (never flushes).
// Must be entered from convergent control flow (see walk_runs).
template <int TP>
__device__ __forceinline__ void walk_runs_quad(const float* __restrict__ gp, const int* __restrict__ ip,
                                               const float* __restrict__ bp, float* vg,
                                               const int4* __restrict__ tab, unsigned Cs, int off_mask) {
  int cur = -1;
  int4 t = make_int4(0, 0, 0, 0);  // vertex ids of the current run: fetched at its first pixel, consumed
                                   // (multiplied, added to the table base) only at its flush, so the load's
                                   // latency hides behind the run's FMAs
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int x = 0; x < kQSeg; x += 4) {
    const float4 G0 = *reinterpret_cast<const float4*>(gp + x);
    const float4 G1 = *reinterpret_cast<const float4*>(gp + TP + x);
    const float4 G2 = *reinterpret_cast<const float4*>(gp + 2 * TP + x);
    const float4 G3 = *reinterpret_cast<const float4*>(gp + 3 * TP + x);
    const int4 iq = *reinterpret_cast<const int4*>(ip + x);
    const float4 p0q = *reinterpret_cast<const float4*>(bp + x);
    const float4 p1q = *reinterpret_cast<const float4*>(bp + TP + x);
    const float4 p2q = *reinterpret_cast<const float4*>(bp + 2 * TP + x);
    const int ids[4] = {iq.x, iq.y, iq.z, iq.w};
    const float g[4][4] = {{G0.x, G0.y, G0.z, G0.w}, {G1.x, G1.y, G1.z, G1.w},
                           {G2.x, G2.y, G2.z, G2.w}, {G3.x, G3.y, G3.z, G3.w}};  // [channel][pixel]
    const float q0[4] = {p0q.x, p0q.y, p0q.z, p0q.w};
    const float q1[4] = {p1q.x, p1q.y, p1q.z, p1q.w};
    const float q2[4] = {p2q.x, p2q.y, p2q.z, p2q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int id = ids[j];
      if (id != cur) {  // run boundary (uniform across the four lanes of the walker)
        if ((cur | off_mask) >= 0) {
          red_add_v4(vg + (unsigned)t.x * Cs, a0[0], a0[1], a0[2], a0[3]);
          red_add_v4(vg + (unsigned)t.y * Cs, a1[0], a1[1], a1[2], a1[3]);
          red_add_v4(vg + (unsigned)t.z * Cs, a2[0], a2[1], a2[2], a2[3]);
        }
        t = tab[max(id, 0)];  // empty pixels read triangle 0 harmlessly; never flushed
#pragma unroll
        for (int k = 0; k < 4; ++k) a0[k] = a1[k] = a2[k] = 0.f;
        cur = id;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a0[k] = fmaf(g[k][j], q0[j], a0[k]);
        a1[k] = fmaf(g[k][j], q1[j], a1[k]);
        a2[k] = fmaf(g[k][j], q2[j], a2[k]);
      }
    }
  }
  if ((cur | off_mask) >= 0) {
    red_add_v4(vg + (unsigned)t.x * Cs, a0[0], a0[1], a0[2], a0[3]);
    red_add_v4(vg + (unsigned)t.y * Cs, a1[0], a1[1], a1[2], a1[3]);
    red_add_v4(vg + (unsigned)t.z * Cs, a2[0], a2[1], a2[2], a2[3]);
  }
}

// Phase B of one thread: two adjacent pixels, NC channels of the staged gradients against the three attribute
// rows of each pixel's triangle (rows are shared when both pixels show the same triangle).  NC > 0: compile-
// time channel count (full unroll, immediate offsets); NC == 0: nc channels at run time.
template <int TP, int NC, bool V8>
__device__ __forceinline__ void bary_grad_pair(const float* __restrict__ gq_base, const float* __restrict__ an,
                                               unsigned rs, int4 t0, int4 t1, bool same, int nc,
                                               float (&g0)[3], float (&g1)[3]) {
  const float* r00 = an + (unsigned)t0.x * rs; const float* r01 = an + (unsigned)t0.y * rs;
  const float* r02 = an + (unsigned)t0.z * rs;
  const float* r10 = an + (unsigned)t1.x * rs; const float* r11 = an + (unsigned)t1.y * rs;
  const float* r12 = an + (unsigned)t1.z * rs;
  auto group = [&](int cc) {
    float2 gq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) gq[k] = *reinterpret_cast<const float2*>(gq_base + (cc + k) * TP + (cc >> 2) * 4);
    float4 A0 = *reinterpret_cast<const float4*>(r00 + cc);
    float4 A1 = *reinterpret_cast<const float4*>(r01 + cc);
    float4 A2 = *reinterpret_cast<const float4*>(r02 + cc);
    g0[0] = fmaf(gq[3].x, A0.w, fmaf(gq[2].x, A0.z, fmaf(gq[1].x, A0.y, fmaf(gq[0].x, A0.x, g0[0]))));
    g0[1] = fmaf(gq[3].x, A1.w, fmaf(gq[2].x, A1.z, fmaf(gq[1].x, A1.y, fmaf(gq[0].x, A1.x, g0[1]))));
    g0[2] = fmaf(gq[3].x, A2.w, fmaf(gq[2].x, A2.z, fmaf(gq[1].x, A2.y, fmaf(gq[0].x, A2.x, g0[2]))));
    if (!same) {
      A0 = *reinterpret_cast<const float4*>(r10 + cc);
      A1 = *reinterpret_cast<const float4*>(r11 + cc);
      A2 = *reinterpret_cast<const float4*>(r12 + cc);
    }
    g1[0] = fmaf(gq[3].y, A0.w, fmaf(gq[2].y, A0.z, fmaf(gq[1].y, A0.y, fmaf(gq[0].y, A0.x, g1[0]))));
    g1[1] = fmaf(gq[3].y, A1.w, fmaf(gq[2].y, A1.z, fmaf(gq[1].y, A1.y, fmaf(gq[0].y, A1.x, g1[1]))));
    g1[2] = fmaf(gq[3].y, A2.w, fmaf(gq[2].y, A2.z, fmaf(gq[1].y, A2.y, fmaf(gq[0].y, A2.x, g1[2]))));
  };
  // 256-bit row loads (attribute rows 32-B aligned; SASS LDG.E.ENL2.256): half the load instructions and L1
  // tag requests of the gathers (0.779 -> 0.750 ms)
  auto group8 = [&](int cc) {
    float2 gq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) gq[k] = *reinterpret_cast<const float2*>(gq_base + (cc + k) * TP + ((cc + k) >> 2) * 4);
    float8 A0 = ldg_f8(r00 + cc), A1 = ldg_f8(r01 + cc), A2 = ldg_f8(r02 + cc);
    auto dot8 = [&](const float8& A, float acc, bool second) {
      const float x0 = second ? gq[0].y : gq[0].x, x1 = second ? gq[1].y : gq[1].x, x2 = second ? gq[2].y : gq[2].x,
                  x3 = second ? gq[3].y : gq[3].x, x4 = second ? gq[4].y : gq[4].x, x5 = second ? gq[5].y : gq[5].x,
                  x6 = second ? gq[6].y : gq[6].x, x7 = second ? gq[7].y : gq[7].x;
      acc = fmaf(x3, A.lo.w, fmaf(x2, A.lo.z, fmaf(x1, A.lo.y, fmaf(x0, A.lo.x, acc))));
      return fmaf(x7, A.hi.w, fmaf(x6, A.hi.z, fmaf(x5, A.hi.y, fmaf(x4, A.hi.x, acc))));
    };
    g0[0] = dot8(A0, g0[0], false); g0[1] = dot8(A1, g0[1], false); g0[2] = dot8(A2, g0[2], false);
    if (!same) { A0 = ldg_f8(r10 + cc); A1 = ldg_f8(r11 + cc); A2 = ldg_f8(r12 + cc); }
    g1[0] = dot8(A0, g1[0], true); g1[1] = dot8(A1, g1[1], true); g1[2] = dot8(A2, g1[2], true);
  };
  if (NC > 0 && V8) {
#pragma unroll
    for (int cc = 0; cc < NC; cc += 8) group8(cc);
  } else if (NC > 0) {
#pragma unroll
    for (int cc = 0; cc < NC; cc += 4) group(cc);
  } else {
#pragma unroll 1
    for (int cc = 0; cc < nc; cc += 4) group(cc);
  }
}

// MULTI: C > 16, i.e. several channel passes per tile (phase-B sums are then carried between passes)
template <int TP, bool NEED_VERT, bool NEED_BARY, bool AVEC, bool MULTI>
__global__ void __launch_bounds__(32 * (TP / kQUnit * 2 + kQProducers), kQCtas)
interp_bwd_quad_kernel(InterpBwdArgs b, float* __restrict__ vert_grad, float* __restrict__ bary_grad,
                       const int4* __restrict__ tab, int tab_img_stride, int tiles_per_img, int num_tiles, bool v8) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  QSmem<TP>& S = *reinterpret_cast<QSmem<TP>*>(smem_raw);
  constexpr int TEAM = TP / kQUnit;        // warps per team
  constexpr int CONSUMERS = 2 * TEAM;      // consumer warps
  const InterpArgs& a = b.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int HW = a.H * a.W;
  if (tid == 0) {
    for (int s = 0; s < kQStages; ++s) {
      mbar_init(reinterpret_cast<uint64_t*>(&S.full[s]), kQProducers);
      mbar_init(reinterpret_cast<uint64_t*>(&S.empty[s]), CONSUMERS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int nchunks = MULTI ? (a.C + kQCh - 1) / kQCh : 1;
  const int warp_role = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
  // Tile order: chunks of kQChunk consecutive tiles are dealt round-robin to the CTAs (consecutive tiles share
  // triangle-table lines and attribute rows in this SM's L1; a vertical-neighbour order was measured and made
  // no difference: the kernel is bound by L1 data-pipe throughput, not by L2 latency).
  auto decode = [&](int t, int& n, int& tl) {
    n = t / tiles_per_img;
    tl = t - n * tiles_per_img;
  };

  if (warp_role >= CONSUMERS) {  // ---- producer warps: one bulk copy per plane ----
    if (lane == 0) {
      const int pw = warp_role - CONSUMERS;
      int item = 0;
      for (int t0 = blockIdx.x * kQChunk; t0 < num_tiles; t0 += gridDim.x * kQChunk)
      for (int tile = t0; tile < min(t0 + kQChunk, num_tiles); ++tile) {
        int n, tl;
        decode(tile, n, tl);
        const int p0 = tl * TP;
        const uint32_t bytes = (uint32_t)min(TP, HW - p0) * 4u;
        for (int chunk = 0; chunk < nchunks; ++chunk, ++item) {
          const int s = item % kQStages;
          if (item >= kQStages)
            mbar_wait_backoff(reinterpret_cast<uint64_t*>(&S.empty[s]), (uint32_t)((item / kQStages - 1) & 1), 100);
          const int c0 = chunk * kQCh, nc = min(kQCh, a.C - c0);
          const int ncopies = nc + (NEED_VERT ? 3 : 0) + 1;
          uint64_t* bar = reinterpret_cast<uint64_t*>(&S.full[s]);
          QStage<TP>& st = S.st[s];
          const int mine = (ncopies - pw + kQProducers - 1) / kQProducers;
          fence_proxy_async_smem();
          mbar_arrive_expect_tx(bar, (uint32_t)mine * bytes);
          for (int k = pw; k < ncopies; k += kQProducers) {
            const void* src;
            void* dst;
            if (k < nc) {
              src = b.grad_out + (int64_t)n * b.gs.s0 + (int64_t)(c0 + k) * b.gs.s1 + p0;
              dst = st.g + k * TP + (k >> 2) * 4;
            } else if (NEED_VERT && k < nc + 3) {
              src = a.bary + (int64_t)n * a.bs.s0 + (int64_t)(k - nc) * a.bs.s1 + p0;
              dst = st.bary + (k - nc) * TP;
            } else {
              src = a.index_img + (int64_t)n * a.is.s0 + p0;
              dst = st.idx;
            }
            bulk_g2s(dst, src, bytes, bar);
          }
        }
      }
    }
    return;
  }

  // ---- consumer warps ----
  const int team = warp_role / TEAM, unit = warp_role - team * TEAM;  // a unit is 128 pixels of the tile
  uint32_t phase_bits = 0;
  float gb[MULTI ? 2 : 1][2][3];  // phase-B accumulators: [pass][pixel][vertex], carried across channel passes (C > 16)
  int item = 0, seq = 0;
  for (int t0 = blockIdx.x * kQChunk; t0 < num_tiles; t0 += gridDim.x * kQChunk)
  for (int tile = t0; tile < min(t0 + kQChunk, num_tiles); ++tile, ++seq) {
    int n, tl;
    decode(tile, n, tl);
    const int p0 = tl * TP;
    const int npx = min(TP, HW - p0);
    const bool a_team = ((team ^ seq) & 1) == 0;  // the teams swap roles every tile
    const int4* tabn = tab + (size_t)((unsigned)n * (unsigned)tab_img_stride);
    for (int chunk = 0; chunk < nchunks; ++chunk, ++item) {
      const int s = item % kQStages;
      const int c0 = chunk * kQCh, nc = min(kQCh, a.C - c0);
      QStage<TP>& st = S.st[s];
      mbar_wait_backoff(reinterpret_cast<uint64_t*>(&S.full[s]), (phase_bits >> s) & 1u, 32);
      phase_bits ^= (1u << s);
      if (npx < TP) {  // last tile of an image (uniform over the CTA): pad with "no triangle"
        for (int i = npx + tid; i < TP; i += 32 * CONSUMERS) st.idx[i] = -1;
        asm volatile("bar.sync 1, %0;" :: "n"(32 * CONSUMERS) : "memory");
      }

      if (NEED_VERT && a_team) {  // ---- phase A: 8 walkers x 16 pixels, lane = 4 channels ----
        const int q = lane & 3, seg = unit * kQUnit + (lane >> 2) * kQSeg;
        const bool c_on = 4 * q < nc;
        const int qq = c_on ? q : 0;
        walk_runs_quad<TP>(st.g + (4 * qq) * TP + qq * 4 + seg, st.idx + seg, st.bary + seg,
                           vert_grad + ((size_t)n * (size_t)a.V * (size_t)a.C + (size_t)(c0 + 4 * qq)), tabn,
                           (unsigned)a.C, c_on ? 0 : -1);
      }

      if (NEED_BARY && !a_team) {  // ---- phase B: 2 passes x (thread = 2 pixels) ----
        // triangle rows of both passes are fetched up front (two dependent L2 round trips per pass otherwise)
        int2 qids[2];
        int4 t0s[2], t1s[2];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          qids[pass] = *reinterpret_cast<const int2*>(st.idx + unit * kQUnit + pass * 64 + lane * 2);
          t0s[pass] = tabn[max(qids[pass].x, 0)];
          t1s[pass] = tabn[max(qids[pass].y, 0)];
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {  // unrolled: gb[pass] must stay in registers
          const int x = unit * kQUnit + pass * 64 + lane * 2;
          float g0[3], g1[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            g0[k] = (MULTI && chunk) ? gb[MULTI ? pass : 0][0][k] : 0.f;
            g1[k] = (MULTI && chunk) ? gb[MULTI ? pass : 0][1][k] : 0.f;
          }
          const int2 qid = qids[pass];
          const int4 t0 = t0s[pass], t1 = t1s[pass];
          const bool same = qid.y == qid.x;
          const float* gq_base = st.g + x;
          if (AVEC) {
            const float* an = a.attr + ((size_t)n * (size_t)a.as.s0 + (size_t)c0);  // s2 == 1 on this path
            if (nc == kQCh && v8) bary_grad_pair<TP, kQCh, true>(gq_base, an, (unsigned)a.as.s1, t0, t1, same, nc, g0, g1);
            else if (nc == kQCh) bary_grad_pair<TP, kQCh, false>(gq_base, an, (unsigned)a.as.s1, t0, t1, same, nc, g0, g1);
            else bary_grad_pair<TP, 0, false>(gq_base, an, (unsigned)a.as.s1, t0, t1, same, nc, g0, g1);
          } else {
            const float* an = a.attr + (int64_t)n * a.as.s0 + (int64_t)c0 * a.as.s2;
            const float* r00 = an + (int64_t)t0.x * a.as.s1; const float* r01 = an + (int64_t)t0.y * a.as.s1;
            const float* r02 = an + (int64_t)t0.z * a.as.s1;
            const float* r10 = an + (int64_t)t1.x * a.as.s1; const float* r11 = an + (int64_t)t1.y * a.as.s1;
            const float* r12 = an + (int64_t)t1.z * a.as.s1;
#pragma unroll 1
            for (int cc = 0; cc < nc; ++cc) {
              const float2 gq = *reinterpret_cast<const float2*>(gq_base + cc * TP + (cc >> 2) * 4);
              const int64_t co = (int64_t)cc * a.as.s2;
              g0[0] = fmaf(gq.x, r00[co], g0[0]); g0[1] = fmaf(gq.x, r01[co], g0[1]); g0[2] = fmaf(gq.x, r02[co], g0[2]);
              g1[0] = fmaf(gq.y, r10[co], g1[0]); g1[1] = fmaf(gq.y, r11[co], g1[1]); g1[2] = fmaf(gq.y, r12[co], g1[2]);
            }
          }
          if (!MULTI || chunk == nchunks - 1) {
            if (x < npx) {
              float* gp = bary_grad + ((size_t)n * 3 * (size_t)HW + (size_t)(p0 + x));
              if (qid.x < 0) g0[0] = g0[1] = g0[2] = 0.f;  // (:282-297) zeros where empty
              if (qid.y < 0) g1[0] = g1[1] = g1[2] = 0.f;
#pragma unroll
              for (int k = 0; k < 3; ++k) *reinterpret_cast<float2*>(gp + (size_t)k * HW) = make_float2(g0[k], g1[k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) { gb[MULTI ? pass : 0][0][k] = g0[k]; gb[MULTI ? pass : 0][1][k] = g1[k]; }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(&S.empty[s]));
    }
  }
}

// Developer switches, read ONCE per process (never on the launch path):
//   DRTK_B200_BWD_V4=1   take the round-1 v4 tile kernel instead of the quad-walker kernel (A/B runs)
struct BwdSwitches {
  bool v4;
  BwdSwitches() : v4(getenv("DRTK_B200_BWD_V4") != nullptr) {}
};
inline const BwdSwitches& bwd_switches() {
  static const BwdSwitches s;
  return s;
}

inline size_t bwd_table_bytes(int64_t tab_imgs, int64_t F) { return sizeof(int4) * (size_t)(tab_imgs * F) + 32; }

}  // namespace
}  // namespace drtk

using namespace drtk;

static int fill_args(InterpArgs& a, const float* attr, const int64_t* attr_strides, const int32_t* vi,
                     const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                     const float* bary_img, const int64_t* bary_strides, int64_t N, int64_t V, int64_t F,
                     int64_t C, int64_t H, int64_t W) {
  // an empty face list (F == 0) may come with a null vi pointer: no pixel can reference a triangle then
  if (!attr || (!vi && F > 0) || !index_img || !bary_img || !attr_strides || !vi_strides || !index_strides ||
      !bary_strides)
    return DRTK_B200_EINVAL;
  if (H > (1 << 30) || W > (1 << 30) || N > (1 << 30) || C > (1 << 20)) return DRTK_B200_EUNSUPPORTED;
  a.attr = attr; a.as = make3(attr_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.index_img = index_img; a.is = make3(index_strides); a.bary = bary_img; a.bs = make4(bary_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.C = (int)C; a.H = (int)H; a.W = (int)W;
  return 0;
}

extern "C" int drtk_b200_interpolate_forward(const float* vert_attributes, const int64_t* attr_strides,
                                             const int32_t* vi, const int64_t* vi_strides,
                                             const int32_t* index_img, const int64_t* index_strides,
                                             const float* bary_img, const int64_t* bary_strides,
                                             int64_t N, int64_t V, int64_t F, int64_t C, int64_t H,
                                             int64_t W, float* out, void* stream_) {
  if (N < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  if (N * C * H * W == 0) return 0;
  if (!out) return DRTK_B200_EINVAL;
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.y: slices of the batch
    if (!vert_attributes || !index_img || !bary_img || !attr_strides || !vi_strides || !index_strides || !bary_strides)
      return DRTK_B200_EINVAL;
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = drtk_b200_interpolate_forward(
          vert_attributes + n0 * attr_strides[0], attr_strides, vi ? vi + n0 * vi_strides[0] : nullptr, vi_strides,
          index_img + n0 * index_strides[0], index_strides, bary_img + n0 * bary_strides[0], bary_strides, nn, V, F, C,
          H, W, out + n0 * C * H * W, stream_);
      if (rc) return rc;
    }
    return 0;
  }
  InterpArgs a;
  const int rc = fill_args(a, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = VecOk::image(index_img, W, a.is.s2, a.is.s1, a.is.s0) &&
                   VecOk::image(bary_img, W, a.bs.s3, a.bs.s2, a.bs.s1, a.bs.s0);
  const bool avec = (C % 4 == 0) && a.as.s2 == 1 && (a.as.s1 % 4 == 0) && (a.as.s0 % 4 == 0) && a.as.s1 > 0 &&
                    (V * a.as.s1 + C < (int64_t)0x7FFFFFF0) && (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
  if (H * W >= (int64_t)0x7FFFFFF0) return DRTK_B200_EUNSUPPORTED;
  // Grid: one wave of co-resident CTAs (grid-stride loops inside).  The CTAs of all images together must not
  // exceed what is resident at once -- 148 SMs x occupancy -- or the last, partly filled wave costs up to a
  // third of the kernel (measured at config 4: 1184 CTAs on 444 slots = 2.67 waves).
  auto launch = [&](auto kern) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ < 1) occ = 1;
    const int64_t slots = (int64_t)num_sms() * occ;
    const int64_t need = ((vec ? H * W / 4 : H * W) + 255) / 256;
    int64_t gx = slots / N;
    if (gx < 1) gx = 1;
    if (gx > need) gx = need;
    kern<<<dim3((unsigned)gx, (unsigned)N), 256, 0, stream>>>(a, out);
  };
  if (vec && avec) launch(interp_fwd_kernel<true, true>);
  else if (vec) launch(interp_fwd_kernel<true, false>);
  else if (avec) launch(interp_fwd_kernel<false, true>);
  else launch(interp_fwd_kernel<false, false>);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t drtk_b200_interpolate_backward_workspace_bytes(int64_t N, int64_t F, int64_t vi_batch_stride) {
  if (N <= 0 || F <= 0) return 0;
  return bwd_table_bytes(vi_batch_stride == 0 ? 1 : N, F);  // the packed int4 triangle table of the quad-walker path
}

extern "C" int drtk_b200_interpolate_backward(
    const float* grad_out, const int64_t* grad_out_strides, const float* vert_attributes,
    const int64_t* attr_strides, const int32_t* vi, const int64_t* vi_strides, const int32_t* index_img,
    const int64_t* index_strides, const float* bary_img, const int64_t* bary_strides, int64_t N,
    int64_t V, int64_t F, int64_t C, int64_t H, int64_t W, float* vert_attributes_grad,
    float* bary_img_grad, void* workspace, size_t workspace_bytes, void* stream_) {
  if (N < 0 || V < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.y in the generic kernels: slices of the batch
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = drtk_b200_interpolate_backward(
          grad_out ? grad_out + n0 * grad_out_strides[0] : nullptr, grad_out_strides,
          vert_attributes ? vert_attributes + n0 * attr_strides[0] : nullptr, attr_strides,
          vi ? vi + n0 * vi_strides[0] : nullptr, vi_strides, index_img ? index_img + n0 * index_strides[0] : nullptr,
          index_strides, bary_img ? bary_img + n0 * bary_strides[0] : nullptr, bary_strides, nn, V, F, C, H, W,
          vert_attributes_grad ? vert_attributes_grad + n0 * V * C : nullptr,
          bary_img_grad ? bary_img_grad + n0 * 3 * H * W : nullptr, workspace, workspace_bytes, stream_);
      if (rc) return rc;
    }
    return 0;
  }
  // (:661) the vertex-gradient table starts at zero: vi_table_kernel does it on the quad-walker path, a memset elsewhere
  auto zero_vert_grad = [&]() -> int {
    if (vert_attributes_grad && N * V * C > 0)
      DRTK_CUDA(cudaMemsetAsync(vert_attributes_grad, 0, sizeof(float) * (size_t)(N * V * C), stream));
    return 0;
  };
  const int64_t npix = N * H * W;
  if (npix == 0) return zero_vert_grad();
  if (!vert_attributes_grad && !bary_img_grad) return 0;
  if (C == 0) {
    if (bary_img_grad) DRTK_CUDA(cudaMemsetAsync(bary_img_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return 0;
  }
  if (!grad_out || !grad_out_strides) return DRTK_B200_EINVAL;
  if (V == 0 || F == 0) {  // nothing can be covered: all gradients are zero (index_img must be all -1)
    if (bary_img_grad) DRTK_CUDA(cudaMemsetAsync(bary_img_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return zero_vert_grad();
  }
  InterpBwdArgs b;
  const int rc = fill_args(b.f, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  b.grad_out = grad_out; b.gs = make4(grad_out_strides);
  const bool nv = vert_attributes_grad != nullptr, nb = bary_img_grad != nullptr;

  // fast path: tiles staged through shared memory by bulk-async copies; needs dense, 16-B aligned rows
  // (dense H*W planes, 16-B aligned plane starts)
  const bool rows_ok =
      ((H * W) % 4 == 0) && (H * W < (int64_t)0x7FFFFFF0) && (V * C < (int64_t)0x7FFFFFF0) &&
      (F == 0 || (F * (b.f.vis.s1 > 0 ? b.f.vis.s1 : 1) < (int64_t)0x7FFFFFF0 && b.f.vis.s2 >= 0 && b.f.vis.s2 < (1 << 28) && b.f.vis.s1 >= 0)) &&
      VecOk::image(grad_out, 4, b.gs.s3, 4, b.gs.s1, b.gs.s0) && b.gs.s2 == W &&
      VecOk::image(index_img, 4, b.f.is.s2, 4, b.f.is.s0) && b.f.is.s1 == W &&
      (!nv || (VecOk::image(bary_img, 4, b.f.bs.s3, 4, b.f.bs.s1, b.f.bs.s0) && b.f.bs.s2 == W)) &&
      (!nb || reinterpret_cast<uintptr_t>(bary_img_grad) % 16 == 0);
  if (rows_ok) {
    const bool avec = nb && (C % 4 == 0) && b.f.as.s2 == 1 && (b.f.as.s1 % 4 == 0) && (b.f.as.s0 % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0) &&
                      (V * (b.f.as.s1 > 0 ? b.f.as.s1 : 1) + C < (int64_t)0x7FFFFFF0) && b.f.as.s1 >= 0;
    int rc2 = 0;
    // v5: quad-lane walkers + packed triangle table (needs whole groups of four channels)
    constexpr int QTP = DRTK_INTERP_BWD_TP;
    const int64_t tiles_q = N * ((H * W + QTP - 1) / QTP);
    const int64_t tab_imgs = (b.f.vis.s0 == 0) ? 1 : N;
    if ((C % 4 == 0) && tiles_q < (int64_t)0x7FFFFFF0 && tab_imgs * F < (int64_t)0x0FFFFFFF &&
        (!nv || reinterpret_cast<uintptr_t>(vert_attributes_grad) % 16 == 0) && !bwd_switches().v4) {
      // the packed triangle table lives in the caller's workspace: the library never allocates
      if (!workspace || workspace_bytes < bwd_table_bytes(tab_imgs, F)) return DRTK_B200_EWORKSPACE;
      int4* tab = reinterpret_cast<int4*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~uintptr_t(15));
      const int64_t total = tab_imgs * F;
      vi_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
          vi, b.f.vis, (int)F, total, tab, reinterpret_cast<float4*>(vert_attributes_grad), nv ? N * V * C / 4 : 0);
      // 256-bit attribute-row loads need 32-B aligned rows
      const bool v8 = avec && (b.f.as.s1 % 8 == 0) && (b.f.as.s0 % 8 == 0) && (C % 8 == 0) &&
                      (reinterpret_cast<uintptr_t>(vert_attributes) % 32 == 0);
      auto launch5 = [&](auto kern) {
        const size_t smem = sizeof(QSmem<QTP>) + 128;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { rc2 = (int)e; return; }
        const int tiles_per_img = (int)((H * W + QTP - 1) / QTP);
        const int64_t ctas = (int64_t)kQCtas * num_sms();  // co-resident CTAs per SM
        const int64_t chunks = (tiles_q + kQChunk - 1) / kQChunk;
        const unsigned grid = (unsigned)(chunks < ctas ? chunks : ctas);
        kern<<<grid, 32 * (QTP / kQUnit * 2 + kQProducers), smem, stream>>>(
            b, vert_attributes_grad, bary_img_grad, tab, tab_imgs == 1 ? 0 : (int)F, tiles_per_img, (int)tiles_q, v8);
      };
#define DRTK_Q5(MULTI)                                                                                     \
      do {                                                                                                   \
        if (nv && nb) { if (avec) launch5(interp_bwd_quad_kernel<QTP, true, true, true, MULTI>);             \
                        else launch5(interp_bwd_quad_kernel<QTP, true, true, false, MULTI>); }               \
        else if (nv) launch5(interp_bwd_quad_kernel<QTP, true, false, false, MULTI>);                        \
        else { if (avec) launch5(interp_bwd_quad_kernel<QTP, false, true, true, MULTI>);                     \
               else launch5(interp_bwd_quad_kernel<QTP, false, true, false, MULTI>); }                       \
      } while (0)
      if (C > kQCh) DRTK_Q5(true); else DRTK_Q5(false);
#undef DRTK_Q5
      if (rc2) return rc2;
      DRTK_CHECK_LAUNCH();
      return 0;
    }
    if (const int rcz = zero_vert_grad()) return rcz;
    auto launch = [&](auto kern, size_t smem, int TP) {
      const int tiles_per_img = (int)((H * W + TP - 1) / TP);
      const int64_t num_tiles = N * (int64_t)tiles_per_img;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { rc2 = (int)e; return; }
      const unsigned grid = (unsigned)(num_tiles < num_sms() ? num_tiles : (int64_t)num_sms());  // one persistent CTA per SM
      kern<<<grid, kBwdBlock, smem, stream>>>(b, vert_attributes_grad, bary_img_grad, tiles_per_img, num_tiles);
    };
#define DRTK_BWD_TILE(LPW)                                                                                  \
    do {                                                                                                    \
      const size_t smem = sizeof(BwdTileSmem<LPW>) + 128;                                                   \
      const int TP = BwdTileCfg<LPW>::TP;                                                                   \
      if (nv && nb) { if (avec) launch(interp_bwd_tile_kernel<LPW, true, true, true>, smem, TP);            \
                      else launch(interp_bwd_tile_kernel<LPW, true, true, false>, smem, TP); }              \
      else if (nv) launch(interp_bwd_tile_kernel<LPW, true, false, false>, smem, TP);                       \
      else { if (avec) launch(interp_bwd_tile_kernel<LPW, false, true, true>, smem, TP);                    \
             else launch(interp_bwd_tile_kernel<LPW, false, true, false>, smem, TP); }                      \
    } while (0)
    if (C <= 4) DRTK_BWD_TILE(4);
    else if (C <= 8) DRTK_BWD_TILE(8);
    else DRTK_BWD_TILE(16);
#undef DRTK_BWD_TILE
    if (rc2) return rc2;
    DRTK_CHECK_LAUNCH();
    return 0;
  }

  // generic path (arbitrary strides / odd widths): one thread per pixel, segmented shuffle reduction
  if (H * W >= (int64_t)0x7FFFFFF0) return DRTK_B200_EUNSUPPORTED;
  if (const int rcz = zero_vert_grad()) return rcz;
  const dim3 blocks((unsigned)((H * W + 255) / 256), (unsigned)N);
  const bool rv4 = (C % 4 == 0) && (reinterpret_cast<uintptr_t>(vert_attributes_grad) % 16 == 0);
  if (nv && nb) {
    if (rv4) interp_bwd_kernel<true, true, true><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, bary_img_grad);
    else interp_bwd_kernel<true, true, false><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, bary_img_grad);
  } else if (nv) {
    if (rv4) interp_bwd_kernel<true, false, true><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, nullptr);
    else interp_bwd_kernel<true, false, false><<<blocks, 256, 0, stream>>>(b, vert_attributes_grad, nullptr);
  } else {
    interp_bwd_kernel<false, true, false><<<blocks, 256, 0, stream>>>(b, nullptr, bary_img_grad);
  }
  DRTK_CHECK_LAUNCH();
  return 0;
}
