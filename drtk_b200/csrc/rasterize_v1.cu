// rasterize_v1.cu -- the ROUND-1 tile rasteriser (per-sample tests + 64-bit shared-memory CAS z-buffer), kept for
// one round as the A/B baseline of the record/span/owner design in rasterize.cu: DRTK_B200_RASTER_V1=1 (read once
// per process) routes drtk_b200_rasterize here.  Same contract, same results, bit for bit.
//
// Contract (bit-exact with the reference CUDA build): src/rasterize/rasterize_kernel.cu:42-168
// (+ unpack :402-415, memset :484-488) of facebookresearch/DRTK.  Per (triangle, pixel)
// sample the arithmetic below reproduces the reference's compiled sm_100 SASS
// (--use_fast_math: FTZ, MUFU.RCP, one specific FMA contraction per expression) with
// explicit intrinsics, so that the packed (depth_bits << 32 | triangle_id) minimum -- an
// order-independent quantity -- comes out identical however the work is organised.
//
// Organisation (NOT the reference's thread-per-triangle + 64-bit global atomics + memset +
// unpack): triangles are binned to 32x32-pixel screen tiles; one CTA per tile keeps the packed
// z-buffer of its tile in shared memory, resolves it there and writes index_img / depth_img
// once with coalesced 128-bit stores.  Global traffic: the 8 B/px outputs plus the bin lists.
//
//   bin_count  -> scan_offsets -> bin_fill -> raster_tiles
//
// A triangle whose clamped bounding box spans at most 2x2 tiles ("small") is appended to
// those tiles' lists (<= 4 entries, so the list storage is bounded by 4*N*F and the call needs
// no device->host sync to size anything).  Everything else ("large") goes to one per-image
// list that every tile of that image walks cooperatively (all threads of the CTA split the
// pixels of the clipped bounding box).
#include "common.cuh"

namespace drtk {
namespace {

constexpr int kTileLog = 5;
constexpr int kTile = 1 << kTileLog;        // 32 x 32 pixels
constexpr int kTilePix = kTile * kTile;     // 1024
constexpr int kRasterThreads = 128;

struct RasterArgs {
  const float* v;
  Strides3 vs;
  const int32_t* vi;
  Strides3 vis;
  int N, V, F, H, W;
  int tilesX, tilesY;
};

// Everything a sample test needs, derived once per triangle.
struct TriSetup {
  // canonical edges k = 0,1,2  <->  (v1,v2), (v2,v0), (v0,v1); origin o = endpoint with the lower
  // vertex index (src/rasterize/rasterize_kernel.cu:29-40).  (ax, ay) is the edge direction times
  // s = sign(den) * (swapped ? -1 : 1): b_k = s * fma(-ab.y, p.x-o.x, rn((p.y-o.y)*ab.x)) equals
  // fma(-ay, p.x-o.x, rn((p.y-o.y)*ax)) bit for bit (round-to-nearest is sign symmetric), which
  // saves the three multiplications by +-1 per sample.
  float ox[3], oy[3], ax[3], ay[3];
  float d0, d1, d2;  // MUFU.RCP(epsclamp(z_k))
  float rden;        // MUFU.RCP(|den|)
  bool tl[3];
  int bx0, by0, bx1, by1;  // clamped pixel bounding box (inclusive); may be empty
};

__device__ __forceinline__ void canon_edge_setup(int ia, int ib, float pax, float pay, float pbx,
                                                 float pby, float sgn, float& ox, float& oy,
                                                 float& ax, float& ay) {
  if (ia <= ib) {
    ox = pax; oy = pay;
    ax = mul_rn(sub_rn(pbx, pax), sgn); ay = mul_rn(sub_rn(pby, pay), sgn);
  } else {
    ox = pbx; oy = pby;
    ax = mul_rn(sub_rn(pax, pbx), -sgn); ay = mul_rn(sub_rn(pay, pby), -sgn);
  }
}

// Loads triangle f of image n, applies the reference's rejection rules (:81, :96-100, :107) and
// fills the setup.  Returns false when the triangle produces no samples.
__device__ __forceinline__ bool tri_setup(const RasterArgs& a, int n, int f, TriSetup& s) {
  const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)f * a.vis.s1;
  const int i0 = (int)(((uint32_t)vip[0]) & 0x0FFFFFFFu);  // top nibble reserved (:74)
  const int i1 = vip[a.vis.s2];
  const int i2 = vip[2 * a.vis.s2];
  if (i0 == i1 && i1 == i2) return false;  // padding triangles (:81)

  const float* vp = a.v + (int64_t)n * a.vs.s0;
  const float* q0 = vp + (int64_t)i0 * a.vs.s1;
  const float* q1 = vp + (int64_t)i1 * a.vs.s1;
  const float* q2 = vp + (int64_t)i2 * a.vs.s1;
  const float p0x = q0[0], p0y = q0[a.vs.s2], z0 = q0[2 * a.vs.s2];
  const float p1x = q1[0], p1y = q1[a.vs.s2], z1 = q1[2 * a.vs.s2];
  const float p2x = q2[0], p2y = q2[a.vs.s2], z2 = q2[2 * a.vs.s2];

  if (!(z0 > 1e-8f && z1 > 1e-8f && z2 > 1e-8f)) return false;  // (:96)
  const float mnx = fminf(fminf(p0x, p1x), p2x), mny = fminf(fminf(p0y, p1y), p2y);
  const float mxx = fmaxf(fmaxf(p0x, p1x), p2x), mxy = fmaxf(fmaxf(p0y, p1y), p2y);
  if (!(mnx <= (float)(a.W - 1) && mny <= (float)(a.H - 1) && mxx > 0.f && mxy > 0.f))
    return false;  // (:97-98)

  const float v01x = sub_rn(p1x, p0x), v01y = sub_rn(p1y, p0y);
  const float v02x = sub_rn(p2x, p0x), v02y = sub_rn(p2y, p0y);
  const float v12x = sub_rn(p2x, p1x), v12y = sub_rn(p2y, p1y);
  const float den = diff_of_products(v01x, v02y, v01y, v02x);  // (:105) FMUL + FFMA as compiled
  if (den == 0.f) return false;                                   // (:107)

  // bounding box with the reference's truncation and +1 border (:109-113)
  s.bx0 = max(0, __float2int_rz(mnx));
  s.by0 = max(0, __float2int_rz(mny));
  s.bx1 = min(a.W - 1, (int)((unsigned)__float2int_rz(mxx) + 1u));
  s.by1 = min(a.H - 1, (int)((unsigned)__float2int_rz(mxy) + 1u));

  const float sgn = den > 0.f ? 1.f : -1.f;  // sign(den), den != 0 (:125)
  canon_edge_setup(i1, i2, p1x, p1y, p2x, p2y, sgn, s.ox[0], s.oy[0], s.ax[0], s.ay[0]);
  canon_edge_setup(i2, i0, p2x, p2y, p0x, p0y, sgn, s.ox[1], s.oy[1], s.ax[1], s.ay[1]);
  canon_edge_setup(i0, i1, p0x, p0y, p1x, p1y, sgn, s.ox[2], s.oy[2], s.ax[2], s.ay[2]);

  if (den > 0.f) {  // top-left classification (:133-141)
    s.tl[0] = (v12y < 0.f) || (v12y == 0.f && v12x > 0.f);
    s.tl[1] = (v02y > 0.f) || (v02y == 0.f && v02x < 0.f);
    s.tl[2] = (v01y < 0.f) || (v01y == 0.f && v01x > 0.f);
  } else {
    s.tl[0] = (v12y > 0.f) || (v12y == 0.f && v12x < 0.f);
    s.tl[1] = (v02y < 0.f) || (v02y == 0.f && v02x > 0.f);
    s.tl[2] = (v01y > 0.f) || (v01y == 0.f && v01x < 0.f);
  }
  s.rden = rcp_approx(fabsf(den));  // (:148) under fast-math: bary * MUFU.RCP(|den|)
  s.d0 = rcp_approx(epsclamp(z0));  // (:151)
  s.d1 = rcp_approx(epsclamp(z1));
  s.d2 = rcp_approx(epsclamp(z2));
  return true;
}

// One (triangle, pixel) sample.  Returns true and the depth bits when the pixel centre (x, y)
// is covered under the top-left rule (:118-153).  row[k] = rn((p.y - o_k.y) * ax_k) is per row (the
// reference compiler hoists the same product; same value either way).
__device__ __forceinline__ bool sample(const float (&ox)[3], const float (&ay)[3], const bool (&tl)[3],
                                       float rden, float d0, float d1, float d2, float px,
                                       const float (&row)[3], uint32_t& depth_bits) {
  const float b0 = fma_rn(-ay[0], sub_rn(px, ox[0]), row[0]);
  const float b1 = fma_rn(-ay[1], sub_rn(px, ox[1]), row[1]);
  const float b2 = fma_rn(-ay[2], sub_rn(px, ox[2]), row[2]);
  if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) return false;
  // top-left rule: only reached by samples exactly on an edge (one min3 + compare guards the three
  // equality tests; all b are >= 0 and not NaN here, so min == 0 <=> some b == 0)
  if (fminf(fminf(b0, b1), b2) == 0.f) {
    if ((b0 == 0.f && !tl[0]) || (b1 == 0.f && !tl[1]) || (b2 == 0.f && !tl[2])) return false;
  }
  const float c0 = mul_rn(b0, rden), c1 = mul_rn(b1, rden), c2 = mul_rn(b2, rden);
  // dot(d_inv, bary) as compiled: FMUL(b1,d1) -> FFMA(b0,d0,.) -> FFMA(b2,d2,.)
  const float inv = fma_rn(c2, d2, fma_rn(c0, d0, mul_rn(c1, d1)));
  depth_bits = __float_as_uint(rcp_approx(epsclamp(inv)));
  return true;
}
__device__ __forceinline__ bool sample(const TriSetup& s, float px, const float (&row)[3],
                                       uint32_t& depth_bits) {
  return sample(s.ox, s.ay, s.tl, s.rden, s.d0, s.d1, s.d2, px, row, depth_bits);
}

__device__ __forceinline__ void row_terms(const TriSetup& s, float py, float (&row)[3]) {
  row[0] = mul_rn(sub_rn(py, s.oy[0]), s.ax[0]);
  row[1] = mul_rn(sub_rn(py, s.oy[1]), s.ax[1]);
  row[2] = mul_rn(sub_rn(py, s.oy[2]), s.ax[2]);
}

// Conservative x-range of one image row: pixels outside [xs, xe] cannot pass the edge tests.
// Edge k crosses zero at x* = o.x + row/ay; samples within one pixel of x* are always kept, which
// covers the rounding of the exact test (<= 2^-23 (|dy| |ax/ay| + |dx|) px) as long as the edge is
// not nearly horizontal and the coordinates are moderate; otherwise the edge does not prune.
__device__ __forceinline__ void row_span(const float (&ox)[3], const float (&ax)[3], const float (&ay)[3],
                                         const float (&row)[3], float dy_max, int& xs, int& xe) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float aay = fabsf(ay[k]);
    if (aay * 1048576.f >= fabsf(ax[k]) * dy_max && fabsf(ox[k]) < 1048576.f) {  // (false for ay == 0, NaN)
      const float xstar = fma_rn(row[k], rcp_approx(ay[k]), ox[k]);
      if (fabsf(xstar) < 1.0e9f) {
        if (ay[k] < 0.f) xs = max(xs, __float2int_rd(xstar) - 1);  // b grows with x: x >= x*
        else xe = min(xe, __float2int_ru(xstar) + 1);              // b falls with x: x <= x*
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) bin_kernel(RasterArgs a, int64_t total, uint32_t* tile_count,
                                                  const uint32_t* tile_offset, uint32_t* tile_list,
                                                  uint32_t* large_count, uint32_t* large_list) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx / a.F);
  const int f = (int)(idx - (int64_t)n * a.F);
  TriSetup s;
  if (!tri_setup(a, n, f, s)) return;
  if (s.bx0 > s.bx1 || s.by0 > s.by1) return;
  const int tx0 = s.bx0 >> kTileLog, tx1 = s.bx1 >> kTileLog;
  const int ty0 = s.by0 >> kTileLog, ty1 = s.by1 >> kTileLog;
  const int64_t tbase = (int64_t)n * a.tilesX * a.tilesY;
  if (tx1 - tx0 <= 1 && ty1 - ty0 <= 1) {
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int64_t t = tbase + (int64_t)ty * a.tilesX + tx;
        const uint32_t k = atomicAdd(&tile_count[t], 1u);
        if (FILL) tile_list[tile_offset[t] + k] = (uint32_t)f;
      }
  } else if (FILL) {
    const uint32_t k = atomicAdd(&large_count[n], 1u);
    large_list[(int64_t)n * a.F + k] = (uint32_t)f;
  }
}

// Per-image exclusive scan of the tile counts into list offsets, zeroing `count` so that bin_kernel<true> can
// reuse it as the per-tile cursor.  One CTA of 1024 threads per IMAGE (a small triangle adds at most four list
// entries, so image n owns the fixed list region [n * 4F, (n + 1) * 4F) and the images scan independently):
// T tiles in coalesced slabs of 4096 entries (uint4 per thread) carrying the running total -- one slab at
// config 4, four at config 5.  (A single CTA over all N * T counters took 13.6 us at config 4.)
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* count, uint32_t* offset, int64_t M,
                                                    uint32_t list_stride, bool vec) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  count += (int64_t)blockIdx.x * M;
  offset += (int64_t)blockIdx.x * M;
  if (tid == 0) carry_s = blockIdx.x * list_stride;
  __syncthreads();
  for (int64_t base = 0; base < M; base += 4096) {
    const int64_t i = base + (int64_t)tid * 4;
    uint32_t c[4] = {0u, 0u, 0u, 0u};
    if (vec && i + 3 < M) {
      const uint4 q = *reinterpret_cast<const uint4*>(count + i);
      c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w;
    } else {
      for (int k = 0; k < 4; ++k) if (i + k < M) c[k] = count[i + k];
    }
    const uint32_t sum = c[0] + c[1] + c[2] + c[3];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      uint32_t w = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_sums[lane] = w;
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t run = carry + inc - sum + (wid ? warp_sums[wid - 1] : 0u);
    uint32_t o4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { o4[k] = run; run += c[k]; }
    if (vec && i + 3 < M) {
      *reinterpret_cast<uint4*>(offset + i) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      *reinterpret_cast<uint4*>(count + i) = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int k = 0; k < 4; ++k) if (i + k < M) { offset[i + k] = o4[k]; count[i + k] = 0u; }
    }
    __syncthreads();
    if (tid == 1023) carry_s = carry + warp_sums[31];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// per-tile resolve
// ------------------------------------------------------------------------------------------
// Shared-memory records of the triangles of one pass (structure of arrays, one slot per thread).
struct TileRecs {
  float ox[3][kRasterThreads], oy[3][kRasterThreads], ax[3][kRasterThreads], ay[3][kRasterThreads];
  float d[3][kRasterThreads], rden[kRasterThreads];
  int meta[kRasterThreads];  // tl bits 0-2 | bx0 << 3 | bx1 << 8 | by0 << 13   (tile-local 0..31)
  int tri[kRasterThreads];
  int prefix[kRasterThreads + 1];  // exclusive scan of the per-triangle row counts
};

__global__ void __launch_bounds__(kRasterThreads) raster_tiles_kernel(
    RasterArgs a, const uint32_t* __restrict__ tile_count, const uint32_t* __restrict__ tile_offset,
    const uint32_t* __restrict__ tile_list, const uint32_t* __restrict__ large_count,
    const uint32_t* __restrict__ large_list, float* __restrict__ depth_img,
    int32_t* __restrict__ index_img) {
  __shared__ unsigned long long zbuf[kTilePix];
  __shared__ TileRecs R;
  __shared__ __align__(16) int warp_tot[kRasterThreads / 32];  // aligned: its vector load must not straddle R.prefix
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y, n = blockIdx.z;
  const int64_t t = ((int64_t)n * a.tilesY + tile_y) * a.tilesX + tile_x;
  const int x_lo = tile_x << kTileLog, y_lo = tile_y << kTileLog;
  const int x_hi = min(x_lo + kTile - 1, a.W - 1), y_hi = min(y_lo + kTile - 1, a.H - 1);

  for (int i = tid; i < kTilePix; i += kRasterThreads) zbuf[i] = ~0ull;  // (:484-488)

  // (1) small triangles.  Work item = one image row of one triangle's clipped bounding box, so the
  // threads of a warp do equally sized pieces of work whatever the triangle sizes are.
  const uint32_t cnt = tile_count[t];
  const uint32_t* list = tile_list + tile_offset[t];
  for (uint32_t base = 0; base < cnt; base += kRasterThreads) {
    __syncthreads();  // zbuf initialised / previous pass done with the records
    int rows = 0;
    if (base + tid < cnt) {
      const int f = (int)list[base + tid];
      TriSetup s;
      if (tri_setup(a, n, f, s)) {
        const int bx0 = max(s.bx0, x_lo), bx1 = min(s.bx1, x_hi);
        const int by0 = max(s.by0, y_lo), by1 = min(s.by1, y_hi);
        if (bx0 <= bx1 && by0 <= by1) {
          rows = by1 - by0 + 1;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            R.ox[k][tid] = s.ox[k]; R.oy[k][tid] = s.oy[k]; R.ax[k][tid] = s.ax[k]; R.ay[k][tid] = s.ay[k];
          }
          R.d[0][tid] = s.d0; R.d[1][tid] = s.d1; R.d[2][tid] = s.d2; R.rden[tid] = s.rden;
          R.meta[tid] = (s.tl[0] ? 1 : 0) | (s.tl[1] ? 2 : 0) | (s.tl[2] ? 4 : 0) | ((bx0 - x_lo) << 3) |
                        ((bx1 - x_lo) << 8) | ((by0 - y_lo) << 13);
          R.tri[tid] = f;
        }
      }
    }
    // block-wide exclusive scan of `rows`
    int inc = rows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < kRasterThreads / 32; ++w) woff += (w < wid) ? warp_tot[w] : 0;
    R.prefix[tid] = woff + inc - rows;
    if (tid == kRasterThreads - 1) R.prefix[kRasterThreads] = woff + inc;
    __syncthreads();
    const int total = R.prefix[kRasterThreads];

    for (int item = tid; item < total; item += kRasterThreads) {
      // owner = largest slot with prefix[slot] <= item (slots with zero rows are skipped naturally)
      int lo = 0, hi = kRasterThreads;
#pragma unroll
      for (int it = 0; it < 7; ++it) {  // log2(128)
        const int mid = (lo + hi) >> 1;
        if (R.prefix[mid] <= item) lo = mid; else hi = mid;
      }
      const int sl = lo;
      const int meta = R.meta[sl];
      const int ly = ((meta >> 13) & 31) + (item - R.prefix[sl]);
      int xs = (meta >> 3) & 31, xe = (meta >> 8) & 31;
      const float ox[3] = {R.ox[0][sl], R.ox[1][sl], R.ox[2][sl]};
      const float ax[3] = {R.ax[0][sl], R.ax[1][sl], R.ax[2][sl]};
      const float ay[3] = {R.ay[0][sl], R.ay[1][sl], R.ay[2][sl]};
      const bool tl[3] = {(meta & 1) != 0, (meta & 2) != 0, (meta & 4) != 0};
      const float py = (float)(y_lo + ly);
      float row[3], dy_max = 1.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float dy = sub_rn(py, R.oy[k][sl]);
        row[k] = mul_rn(dy, ax[k]);
        dy_max = fmaxf(dy_max, fabsf(dy));
      }
      int gxs = x_lo + xs, gxe = x_lo + xe;
      row_span(ox, ax, ay, row, dy_max, gxs, gxe);
      if (gxs > gxe) continue;
      const float rden = R.rden[sl], d0 = R.d[0][sl], d1 = R.d[1][sl], d2 = R.d[2][sl];
      const unsigned long long f = (unsigned long long)(uint32_t)R.tri[sl];
      unsigned long long* zrow = zbuf + (ly << kTileLog) - x_lo;
      for (int x = gxs; x <= gxe; ++x) {
        uint32_t db;
        if (sample(ox, ay, tl, rden, d0, d1, d2, (float)x, row, db)) {
          // (:155-161) packed minimum.  The first write to a pixel is by far the common case: one native
          // compare-and-swap against "empty" settles it without the load + compare + CAS loop that a 64-bit
          // shared-memory atomicMin compiles to; only a pixel that is already taken pays for the loop.
          const unsigned long long key = ((unsigned long long)db << 32) | f;
          const unsigned long long old = atomicCAS(zrow + x, ~0ull, key);
          if (old != ~0ull && key < old) atomicMin(zrow + x, key);
        }
      }
    }
  }
  __syncthreads();

  // (2) large triangles of this image: whole CTA cooperates on each one
  const uint32_t nlarge = large_count[n];
  const uint32_t* llist = large_list + (int64_t)n * a.F;
  for (uint32_t j = 0; j < nlarge; ++j) {
    const int f = (int)llist[j];
    TriSetup s;
    if (!tri_setup(a, n, f, s)) continue;  // uniform across the CTA
    const int bx0 = max(s.bx0, x_lo), bx1 = min(s.bx1, x_hi);
    const int by0 = max(s.by0, y_lo), by1 = min(s.by1, y_hi);
    if (bx0 > bx1 || by0 > by1) continue;
    const int bw = bx1 - bx0 + 1, npx = bw * (by1 - by0 + 1);
    for (int p = tid; p < npx; p += kRasterThreads) {
      const int yy = p / bw, xx = p - yy * bw;
      const int x = bx0 + xx, y = by0 + yy;
      float row[3];
      row_terms(s, (float)y, row);
      uint32_t db;
      if (sample(s, (float)x, row, db)) {
        const unsigned long long packed = ((unsigned long long)db << 32) | (uint32_t)f;
        atomicMin(&zbuf[((y - y_lo) << kTileLog) + (x - x_lo)], packed);
      }
    }
  }
  __syncthreads();

  // (3) resolve + store (:402-415): empty -> index -1 (low word all ones), depth 0
  const int64_t img_base = (int64_t)n * a.H * a.W;
  if ((a.W & 3) == 0) {
    for (int q = tid; q < kTilePix / 4; q += kRasterThreads) {
      const int ly = q >> 3, lx = (q & 7) << 2;
      const int x = x_lo + lx, y = y_lo + ly;
      if (x > x_hi || y > y_hi) continue;  // W % 4 == 0 -> the whole quad is inside or outside
      int4 id;
      float4 dp;
      const unsigned long long z0 = zbuf[(ly << kTileLog) + lx], z1 = zbuf[(ly << kTileLog) + lx + 1],
                               z2 = zbuf[(ly << kTileLog) + lx + 2], z3 = zbuf[(ly << kTileLog) + lx + 3];
      id.x = (int)(uint32_t)z0; id.y = (int)(uint32_t)z1; id.z = (int)(uint32_t)z2; id.w = (int)(uint32_t)z3;
      const uint32_t d0 = (uint32_t)(z0 >> 32), d1 = (uint32_t)(z1 >> 32), d2 = (uint32_t)(z2 >> 32),
                     d3 = (uint32_t)(z3 >> 32);
      dp.x = d0 == 0xFFFFFFFFu ? 0.f : __uint_as_float(d0);
      dp.y = d1 == 0xFFFFFFFFu ? 0.f : __uint_as_float(d1);
      dp.z = d2 == 0xFFFFFFFFu ? 0.f : __uint_as_float(d2);
      dp.w = d3 == 0xFFFFFFFFu ? 0.f : __uint_as_float(d3);
      const int64_t o = img_base + (int64_t)y * a.W + x;
      stg_stream_i4(index_img + o, id);
      stg_stream_f4(depth_img + o, dp);
    }
  } else {
    for (int q = tid; q < kTilePix; q += kRasterThreads) {
      const int ly = q >> kTileLog, lx = q & (kTile - 1);
      const int x = x_lo + lx, y = y_lo + ly;
      if (x > x_hi || y > y_hi) continue;
      const unsigned long long z = zbuf[q];
      const uint32_t d = (uint32_t)(z >> 32);
      const int64_t o = img_base + (int64_t)y * a.W + x;
      index_img[o] = (int)(uint32_t)z;
      depth_img[o] = d == 0xFFFFFFFFu ? 0.f : __uint_as_float(d);
    }
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct TiledWorkspace {
  size_t off_count, off_offset, off_large_count, off_large_list, off_tile_list, total;
  int64_t M;
};

inline TiledWorkspace tiled_layout(int64_t N, int64_t F, int64_t H, int64_t W) {
  TiledWorkspace w;
  const int64_t tilesX = (W + kTile - 1) >> kTileLog, tilesY = (H + kTile - 1) >> kTileLog;
  w.M = N * tilesX * tilesY;
  size_t o = 0;
  w.off_count = o;       o = align_up(o + sizeof(uint32_t) * (size_t)w.M, 256);
  w.off_large_count = o; o = align_up(o + sizeof(uint32_t) * (size_t)N, 256);
  const size_t zero_end = o;  // [0, zero_end) is memset to 0 per call
  (void)zero_end;
  w.off_offset = o;      o = align_up(o + sizeof(uint32_t) * (size_t)w.M, 256);
  w.off_large_list = o;  o = align_up(o + sizeof(uint32_t) * (size_t)(N * F), 256);
  w.off_tile_list = o;   o = align_up(o + sizeof(uint32_t) * (size_t)(4 * N * F), 256);
  w.total = o;
  return w;
}

}  // namespace

size_t rasterize_v1_workspace_bytes(int64_t N, int64_t F, int64_t H, int64_t W) { return tiled_layout(N, F, H, W).total + 256; }

int rasterize_v1(const float* v, const int64_t* v_strides, const int32_t* vi, const int64_t* vi_strides, int64_t N,
                 int64_t V, int64_t F, int64_t H, int64_t W, float* depth_img, int32_t* index_img, void* workspace,
                 cudaStream_t stream) {
  RasterArgs a;
  a.v = v; a.vs = make3(v_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.H = (int)H; a.W = (int)W;
  a.tilesX = (int)((W + kTile - 1) >> kTileLog);
  a.tilesY = (int)((H + kTile - 1) >> kTileLog);
  const int64_t total = N * F;
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const TiledWorkspace w = tiled_layout(N, F, H, W);
  if (w.M > 0x7FFFFFFFLL) return DRTK_B200_EUNSUPPORTED;
  uint32_t* tile_count = reinterpret_cast<uint32_t*>(ws + w.off_count);
  uint32_t* large_count = reinterpret_cast<uint32_t*>(ws + w.off_large_count);
  uint32_t* tile_offset = reinterpret_cast<uint32_t*>(ws + w.off_offset);
  uint32_t* large_list = reinterpret_cast<uint32_t*>(ws + w.off_large_list);
  uint32_t* tile_list = reinterpret_cast<uint32_t*>(ws + w.off_tile_list);
  DRTK_CUDA(cudaMemsetAsync(ws, 0, w.off_offset, stream));  // tile_count + large_count
  if (total > 0) {
    const unsigned blocks = (unsigned)((total + 255) / 256);
    bin_kernel<false><<<blocks, 256, 0, stream>>>(a, total, tile_count, nullptr, nullptr, nullptr, nullptr);
    DRTK_CHECK_LAUNCH();
    const int64_t T = w.M / N;
    if (4 * F * N > 0xFFFFFFFFLL) return DRTK_B200_EUNSUPPORTED;
    scan_kernel<<<(unsigned)N, 1024, 0, stream>>>(tile_count, tile_offset, T, (uint32_t)(4 * F), (T & 3) == 0);
    DRTK_CHECK_LAUNCH();
    bin_kernel<true><<<blocks, 256, 0, stream>>>(a, total, tile_count, tile_offset, tile_list, large_count, large_list);
    DRTK_CHECK_LAUNCH();
  }
  if (N > 65535 || a.tilesY > 65535) return DRTK_B200_EUNSUPPORTED;
  raster_tiles_kernel<<<dim3((unsigned)a.tilesX, (unsigned)a.tilesY, (unsigned)N), kRasterThreads, 0, stream>>>(
      a, tile_count, tile_offset, tile_list, large_count, large_list, depth_img, index_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}

}  // namespace drtk
