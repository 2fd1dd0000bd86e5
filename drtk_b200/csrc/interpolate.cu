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

inline unsigned grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  const int64_t need = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace
}  // namespace drtk

using namespace drtk;

static int fill_args(InterpArgs& a, const float* attr, const int64_t* attr_strides, const int32_t* vi,
                     const int64_t* vi_strides, const int32_t* index_img, const int64_t* index_strides,
                     const float* bary_img, const int64_t* bary_strides, int64_t N, int64_t V, int64_t F,
                     int64_t C, int64_t H, int64_t W) {
  if (!attr || !vi || !index_img || !bary_img || !attr_strides || !vi_strides || !index_strides ||
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
  InterpArgs a;
  const int rc = fill_args(a, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = VecOk::image(index_img, W, a.is.s2, a.is.s1, a.is.s0) &&
                   VecOk::image(bary_img, W, a.bs.s3, a.bs.s2, a.bs.s1, a.bs.s0);
  const bool avec = (C % 4 == 0) && a.as.s2 == 1 && (a.as.s1 % 4 == 0) && (a.as.s0 % 4 == 0) && a.as.s1 > 0 &&
                    (V * a.as.s1 + C < (int64_t)0x7FFFFFF0) && (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
  if (H * W >= (int64_t)0x7FFFFFF0 || N > 65535) return DRTK_B200_EUNSUPPORTED;
  const int per_img = (int)((8 + N - 1) / N);
  const unsigned gx = grid_for(vec ? H * W / 4 : H * W, 256, per_img > 0 ? per_img : 1);
  const dim3 grid(gx, (unsigned)N);
  if (vec && avec) interp_fwd_kernel<true, true><<<grid, 256, 0, stream>>>(a, out);
  else if (vec) interp_fwd_kernel<true, false><<<grid, 256, 0, stream>>>(a, out);
  else if (avec) interp_fwd_kernel<false, true><<<grid, 256, 0, stream>>>(a, out);
  else interp_fwd_kernel<false, false><<<grid, 256, 0, stream>>>(a, out);
  DRTK_CHECK_LAUNCH();
  return 0;
}

extern "C" int drtk_b200_interpolate_backward(
    const float* grad_out, const int64_t* grad_out_strides, const float* vert_attributes,
    const int64_t* attr_strides, const int32_t* vi, const int64_t* vi_strides, const int32_t* index_img,
    const int64_t* index_strides, const float* bary_img, const int64_t* bary_strides, int64_t N,
    int64_t V, int64_t F, int64_t C, int64_t H, int64_t W, float* vert_attributes_grad,
    float* bary_img_grad, void* stream_) {
  if (N < 0 || V < 0 || C < 0 || H < 0 || W < 0) return DRTK_B200_EINVAL;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (vert_attributes_grad && N * V * C > 0)
    DRTK_CUDA(cudaMemsetAsync(vert_attributes_grad, 0, sizeof(float) * (size_t)(N * V * C), stream));  // (:661)
  const int64_t npix = N * H * W;
  if (npix == 0) return 0;
  if (!vert_attributes_grad && !bary_img_grad) return 0;
  if (C == 0) {
    if (bary_img_grad) DRTK_CUDA(cudaMemsetAsync(bary_img_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return 0;
  }
  if (!grad_out || !grad_out_strides) return DRTK_B200_EINVAL;
  if (V == 0 || F == 0) {  // nothing can be covered: all gradients are zero (index_img must be all -1)
    if (bary_img_grad) DRTK_CUDA(cudaMemsetAsync(bary_img_grad, 0, sizeof(float) * (size_t)(npix * 3), stream));
    return 0;
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
                      (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
    int rc2 = 0;
    auto launch = [&](auto kern, size_t smem, int TP) {
      const int tiles_per_img = (int)((H * W + TP - 1) / TP);
      const int64_t num_tiles = N * (int64_t)tiles_per_img;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { rc2 = (int)e; return; }
      const unsigned grid = (unsigned)(num_tiles < kNumSMs ? num_tiles : kNumSMs);  // one persistent CTA per SM
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
  if (H * W >= (int64_t)0x7FFFFFF0 || N > 65535) return DRTK_B200_EUNSUPPORTED;
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
