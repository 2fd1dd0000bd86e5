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
  const int64_t HW = (int64_t)a.H * a.W;
  constexpr int PX = VEC ? 4 : 1;
  const int64_t ngroups = (int64_t)a.N * HW / PX;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < ngroups;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = q * PX;
    const int n = (int)(pix / HW);
    const int64_t rem = pix - (int64_t)n * HW;
    const int h = (int)(rem / a.W), w = (int)(rem - (int64_t)h * a.W);
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
    float* op = out + (int64_t)n * a.C * HW + rem;

    if (AVEC) {
      for (int c = 0; c < a.C; c += 4) {
        float r[4][PX];  // [channel][pixel]
        float4 A0, A1, A2;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          if (ids[j] != -1) {
            if (!(j > 0 && ids[j] == ids[j - 1])) {
              A0 = *reinterpret_cast<const float4*>(an + (int64_t)i0[j] * a.as.s1 + c);
              A1 = *reinterpret_cast<const float4*>(an + (int64_t)i1[j] * a.as.s1 + c);
              A2 = *reinterpret_cast<const float4*>(an + (int64_t)i2[j] * a.as.s1 + c);
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
  const int64_t HW = (int64_t)a.H * a.W;
  const int64_t npix = (int64_t)a.N * HW;
  const int lane = threadIdx.x & 31;
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = pix < npix;
  int n = 0, h = 0, w = 0, id = -1;
  int64_t rem = 0;
  if (in_range) {
    n = (int)(pix / HW);
    rem = pix - (int64_t)n * HW;
    h = (int)(rem / a.W); w = (int)(rem - (int64_t)h * a.W);
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
  const int64_t key = used ? (((int64_t)n << 32) | (uint32_t)id) : ((int64_t)-1 - lane);
  const int64_t key_up = __shfl_up_sync(0xffffffffu, key, 1);
  const int64_t key_dn = __shfl_down_sync(0xffffffffu, key, 1);
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
// A CTA owns a 128 x 8 pixel tile.  One warp issues cp.async.bulk row copies (TMA engine, SASS
// UBLKCP) that land the tile's grad_out planes (up to LPW channels per pass), bary planes and
// index row segments in shared memory and complete on an mbarrier; nobody spends LSU
// instructions or registers on the 16+4C B/px stream.  Then
//   phase A (vertex grads): "walkers" of LPW lanes -- lane = channel -- walk a row segment in
//     4-pixel steps (one conflict-free LDS.128 of their own channel plane, broadcast LDS.128 of
//     index/bary), accumulate g*bary_k for the current triangle run in three registers and, when
//     the triangle id changes, flush the run with three reductions whose LPW lanes hit LPW
//     consecutive floats of one vertex row (a single coalesced 64-B RED per vertex at C=16).
//     No shuffles, no shared-memory atomics, one reduction per (run, vertex) instead of per pixel.
//   phase B (bary grads): lane = pixel quad; dot products of the staged gradients with the three
//     attribute rows, accumulated across channel passes in registers, 128-bit streaming stores.
constexpr int kTW = 128, kTH = 8, kTP = kTW * kTH;
constexpr int kBwdThreads = 256;

template <int LPW>
struct BwdTileSmem {
  static constexpr int PITCH = kTP + 4;  // plane pitch == 4 (mod 32) words: LDS.128 by LPW lanes of
                                         // consecutive planes is bank-conflict free
  float g[LPW * PITCH];
  float bary[3 * kTP];
  int idx[kTP];
  unsigned long long bar;
};

template <int LPW, bool NEED_VERT, bool NEED_BARY, bool AVEC>
__global__ void __launch_bounds__(kBwdThreads, 2)
interp_bwd_tile_kernel(InterpBwdArgs b, float* __restrict__ vert_grad, float* __restrict__ bary_grad,
                       int tilesX, int tilesY, int64_t num_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  BwdTileSmem<LPW>& S = *reinterpret_cast<BwdTileSmem<LPW>*>(smem_raw);
  constexpr int PITCH = BwdTileSmem<LPW>::PITCH;
  const InterpArgs& a = b.f;
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t HW = (int64_t)a.H * a.W;
  uint64_t* bar = reinterpret_cast<uint64_t*>(&S.bar);
  if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  uint32_t phase = 0;
  const int nchunks = (a.C + LPW - 1) / LPW;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tpi = tilesX * tilesY;
    const int n = (int)(tile / tpi);
    const int tl = (int)(tile - (int64_t)n * tpi);
    const int ty = tl / tilesX, tx = tl - ty * tilesX;
    const int x0 = tx * kTW, y0 = ty * kTH;
    const int tw = min(kTW, a.W - x0), th = min(kTH, a.H - y0);
    const uint32_t row_bytes = (uint32_t)tw * 4u;

    // phase-B ownership: thread -> pixel quad (row qr, columns qx..qx+3)
    const int qr = tid >> 5, qx = (tid & 31) << 2;
    const bool q_in = NEED_BARY && qr < th && qx < tw;
    float gb[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) gb[j][0] = gb[j][1] = gb[j][2] = 0.f;
    int qid[4] = {-1, -1, -1, -1};
    int qv[4][3];

    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int c0 = chunk * LPW;
      const int nc = min(LPW, a.C - c0);
      // ---- stage the tile: warp 0 issues the bulk copies ----
      if (tid < 32) {
        const int extra_planes = (chunk == 0) ? ((NEED_VERT ? 3 : 0) + 1) : 0;
        const int items = (nc + extra_planes) * th;
        if (lane == 0) {
          fence_proxy_async_smem();
          mbar_arrive_expect_tx(bar, (uint32_t)items * row_bytes);
        }
        __syncwarp();
        for (int i = lane; i < items; i += 32) {
          const int pl = i / th, r = i - pl * th;
          const void* src;
          void* dst;
          if (pl < nc) {
            src = b.grad_out + (int64_t)n * b.gs.s0 + (int64_t)(c0 + pl) * b.gs.s1 + (int64_t)(y0 + r) * b.gs.s2 + x0;
            dst = S.g + pl * PITCH + r * kTW;
          } else if (NEED_VERT && pl < nc + 3) {
            const int k = pl - nc;
            src = a.bary + (int64_t)n * a.bs.s0 + (int64_t)k * a.bs.s1 + (int64_t)(y0 + r) * a.bs.s2 + x0;
            dst = S.bary + k * kTP + r * kTW;
          } else {
            src = a.index_img + (int64_t)n * a.is.s0 + (int64_t)(y0 + r) * a.is.s1 + x0;
            dst = S.idx + r * kTW;
          }
          bulk_g2s(dst, src, row_bytes, bar);
        }
      }
      mbar_wait(bar, phase);
      phase ^= 1u;

      // ---- phase A: vertex-attribute gradients ----
      if (NEED_VERT) {
        constexpr int WALKERS = kBwdThreads / LPW;
        constexpr int SEGS = WALKERS / kTH;  // walkers per tile row
        constexpr int SEG = kTW / SEGS;      // pixels per walker
        const int walker = tid / LPW, c = tid - walker * LPW;
        const int row = walker / SEGS, seg = walker - row * SEGS;
        if (row < th) {
          const bool c_on = c < nc;
          const float* gp = S.g + (c_on ? c : 0) * PITCH + row * kTW;
          const int* ip = S.idx + row * kTW;
          const float* b0p = S.bary + row * kTW;
          const float* b1p = b0p + kTP;
          const float* b2p = b1p + kTP;
          float* vg = vert_grad + (int64_t)n * a.V * a.C + c0 + c;
          const int32_t* vib = a.vi + (int64_t)n * a.vis.s0;
          const int xe = min(seg * SEG + SEG, tw);
          int cur = -1, v0 = 0, v1 = 0, v2 = 0;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f;
          for (int x = seg * SEG; x < xe; x += 4) {
            const float4 gq = *reinterpret_cast<const float4*>(gp + x);
            const int4 iq = *reinterpret_cast<const int4*>(ip + x);
            const float4 p0 = *reinterpret_cast<const float4*>(b0p + x);
            const float4 p1 = *reinterpret_cast<const float4*>(b1p + x);
            const float4 p2 = *reinterpret_cast<const float4*>(b2p + x);
            const int ids[4] = {iq.x, iq.y, iq.z, iq.w};
            const float gs[4] = {gq.x, gq.y, gq.z, gq.w};
            const float q0[4] = {p0.x, p0.y, p0.z, p0.w};
            const float q1[4] = {p1.x, p1.y, p1.z, p1.w};
            const float q2[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int id = ids[j];
              if (id != cur) {  // uniform across the walker's lanes
                if (cur >= 0 && c_on) {
                  red_add(vg + (int64_t)v0 * a.C, a0);
                  red_add(vg + (int64_t)v1 * a.C, a1);
                  red_add(vg + (int64_t)v2 * a.C, a2);
                }
                a0 = a1 = a2 = 0.f;
                cur = id;
                if (id >= 0) {
                  const int32_t* vip = vib + (int64_t)id * a.vis.s1;
                  v0 = vip[0]; v1 = vip[a.vis.s2]; v2 = vip[2 * a.vis.s2];
                }
              }
              if (id >= 0) {
                a0 = fmaf(gs[j], q0[j], a0);
                a1 = fmaf(gs[j], q1[j], a1);
                a2 = fmaf(gs[j], q2[j], a2);
              }
            }
          }
          if (cur >= 0 && c_on) {
            red_add(vg + (int64_t)v0 * a.C, a0);
            red_add(vg + (int64_t)v1 * a.C, a1);
            red_add(vg + (int64_t)v2 * a.C, a2);
          }
        }
      }

      // ---- phase B: barycentric gradients (partial over this channel pass) ----
      if (q_in) {
        if (chunk == 0) {
          const int4 iq = *reinterpret_cast<const int4*>(S.idx + qr * kTW + qx);
          qid[0] = iq.x; qid[1] = iq.y; qid[2] = iq.z; qid[3] = iq.w;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            qv[j][0] = qv[j][1] = qv[j][2] = 0;
            if (qid[j] >= 0) {
              if (j > 0 && qid[j] == qid[j - 1]) {
                qv[j][0] = qv[j - 1][0]; qv[j][1] = qv[j - 1][1]; qv[j][2] = qv[j - 1][2];
              } else {
                load_vi(a, n, qid[j], qv[j][0], qv[j][1], qv[j][2]);
              }
            }
          }
        }
        const float* an = a.attr + (int64_t)n * a.as.s0;
        const float* gq_base = S.g + qr * kTW + qx;
        if (AVEC) {
          for (int cc = 0; cc < nc; cc += 4) {
            float4 gq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) gq[k] = *reinterpret_cast<const float4*>(gq_base + (cc + k) * PITCH);
            float4 A0, A1, A2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (qid[j] < 0) continue;
              if (!(j > 0 && qid[j] == qid[j - 1])) {
                A0 = *reinterpret_cast<const float4*>(an + (int64_t)qv[j][0] * a.as.s1 + c0 + cc);
                A1 = *reinterpret_cast<const float4*>(an + (int64_t)qv[j][1] * a.as.s1 + c0 + cc);
                A2 = *reinterpret_cast<const float4*>(an + (int64_t)qv[j][2] * a.as.s1 + c0 + cc);
              }
              const float g0 = j == 0 ? gq[0].x : j == 1 ? gq[0].y : j == 2 ? gq[0].z : gq[0].w;
              const float g1 = j == 0 ? gq[1].x : j == 1 ? gq[1].y : j == 2 ? gq[1].z : gq[1].w;
              const float g2 = j == 0 ? gq[2].x : j == 1 ? gq[2].y : j == 2 ? gq[2].z : gq[2].w;
              const float g3 = j == 0 ? gq[3].x : j == 1 ? gq[3].y : j == 2 ? gq[3].z : gq[3].w;
              gb[j][0] += g0 * A0.x + g1 * A0.y + g2 * A0.z + g3 * A0.w;
              gb[j][1] += g0 * A1.x + g1 * A1.y + g2 * A1.z + g3 * A1.w;
              gb[j][2] += g0 * A2.x + g1 * A2.y + g2 * A2.z + g3 * A2.w;
            }
          }
        } else {
          for (int cc = 0; cc < nc; ++cc) {
            const float4 gq = *reinterpret_cast<const float4*>(gq_base + cc * PITCH);
            const float gs[4] = {gq.x, gq.y, gq.z, gq.w};
            float A0 = 0.f, A1 = 0.f, A2 = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (qid[j] < 0) continue;
              if (!(j > 0 && qid[j] == qid[j - 1])) {
                A0 = an[(int64_t)qv[j][0] * a.as.s1 + (int64_t)(c0 + cc) * a.as.s2];
                A1 = an[(int64_t)qv[j][1] * a.as.s1 + (int64_t)(c0 + cc) * a.as.s2];
                A2 = an[(int64_t)qv[j][2] * a.as.s1 + (int64_t)(c0 + cc) * a.as.s2];
              }
              gb[j][0] += gs[j] * A0; gb[j][1] += gs[j] * A1; gb[j][2] += gs[j] * A2;
            }
          }
        }
      }
      __syncthreads();  // all reads of this pass done before the next bulk copies overwrite the tile
    }
    if (q_in) {
      float* gp = bary_grad + (int64_t)n * 3 * HW + (int64_t)(y0 + qr) * a.W + x0 + qx;
      stg_stream_f4(gp, make_float4(gb[0][0], gb[1][0], gb[2][0], gb[3][0]));
      stg_stream_f4(gp + HW, make_float4(gb[0][1], gb[1][1], gb[2][1], gb[3][1]));
      stg_stream_f4(gp + 2 * HW, make_float4(gb[0][2], gb[1][2], gb[2][2], gb[3][2]));
    }
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
  const bool avec = (C % 4 == 0) && a.as.s2 == 1 && (a.as.s1 % 4 == 0) && (a.as.s0 % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
  const int64_t npix = N * H * W;
  if (vec && avec)
    interp_fwd_kernel<true, true><<<grid_for(npix / 4, 256, 8), 256, 0, stream>>>(a, out);
  else if (vec)
    interp_fwd_kernel<true, false><<<grid_for(npix / 4, 256, 8), 256, 0, stream>>>(a, out);
  else if (avec)
    interp_fwd_kernel<false, true><<<grid_for(npix, 256, 8), 256, 0, stream>>>(a, out);
  else
    interp_fwd_kernel<false, false><<<grid_for(npix, 256, 8), 256, 0, stream>>>(a, out);
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
  InterpBwdArgs b;
  const int rc = fill_args(b.f, vert_attributes, attr_strides, vi, vi_strides, index_img, index_strides,
                           bary_img, bary_strides, N, V, F, C, H, W);
  if (rc) return rc;
  b.grad_out = grad_out; b.gs = make4(grad_out_strides);
  const bool nv = vert_attributes_grad != nullptr, nb = bary_img_grad != nullptr;

  // fast path: tiles staged through shared memory by bulk-async copies; needs dense, 16-B aligned rows
  const bool rows_ok =
      (W % 4 == 0) && VecOk::image(grad_out, W, b.gs.s3, b.gs.s2, b.gs.s1, b.gs.s0) &&
      VecOk::image(index_img, W, b.f.is.s2, b.f.is.s1, b.f.is.s0) &&
      (!nv || VecOk::image(bary_img, W, b.f.bs.s3, b.f.bs.s2, b.f.bs.s1, b.f.bs.s0)) &&
      (!nb || reinterpret_cast<uintptr_t>(bary_img_grad) % 16 == 0);
  if (rows_ok) {
    const bool avec = nb && (C % 4 == 0) && b.f.as.s2 == 1 && (b.f.as.s1 % 4 == 0) && (b.f.as.s0 % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(vert_attributes) % 16 == 0);
    const int tilesX = (int)((W + kTW - 1) / kTW), tilesY = (int)((H + kTH - 1) / kTH);
    const int64_t num_tiles = N * tilesX * tilesY;
    int rc2 = 0;
    auto launch = [&](auto kern, size_t smem) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { rc2 = (int)e; return; }
      int occ = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBwdThreads, smem);
      if (e != cudaSuccess || occ < 1) { rc2 = e != cudaSuccess ? (int)e : DRTK_B200_EUNSUPPORTED; return; }
      const int64_t cap = (int64_t)kNumSMs * occ;
      const unsigned grid = (unsigned)(num_tiles < cap ? num_tiles : cap);
      kern<<<grid, kBwdThreads, smem, stream>>>(b, vert_attributes_grad, bary_img_grad, tilesX, tilesY, num_tiles);
    };
#define DRTK_BWD_TILE(LPW)                                                                                  \
    do {                                                                                                    \
      const size_t smem = sizeof(BwdTileSmem<LPW>) + 128;                                                   \
      if (nv && nb) { if (avec) launch(interp_bwd_tile_kernel<LPW, true, true, true>, smem);                \
                      else launch(interp_bwd_tile_kernel<LPW, true, true, false>, smem); }                  \
      else if (nv) launch(interp_bwd_tile_kernel<LPW, true, false, false>, smem);                           \
      else { if (avec) launch(interp_bwd_tile_kernel<LPW, false, true, true>, smem);                        \
             else launch(interp_bwd_tile_kernel<LPW, false, true, false>, smem); }                          \
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
  const unsigned blocks = (unsigned)((npix + 255) / 256);
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
