// rasterize.cu -- z-buffered triangle coverage for sm_100a.
//
// Contract (bit-exact with the reference CUDA build): src/rasterize/rasterize_kernel.cu:42-168
// (+ unpack :402-415, memset :484-488) of facebookresearch/DRTK.  Per (triangle, pixel)
// sample the arithmetic reproduces the reference's compiled sm_100 SASS (--use_fast_math: FTZ,
// MUFU.RCP, one specific FMA contraction per expression; raster_core.cuh) with explicit
// intrinsics, so that the packed (depth_bits << 32 | triangle_id) minimum -- an
// order-independent quantity -- comes out identical however the work is organised.
//
// Organisation (NOT the reference's thread-per-triangle walk + 64-bit global atomics + memset +
// unpack):
//
//   bin_count -> scan -> bin_fill -> raster_tiles
//
// * The triangle SETUP runs once per triangle, in bin_fill: vertex gathers, culling rules, canonical
//   edges, 1/|den|, 1/z_k, bounding box.  The result is an 80-byte RECORD which bin_fill appends to the
//   list of every 32x32-pixel tile the bounding box touches (at most 2x2 tiles = "small"; list storage is
//   bounded by 4*N*F records, so nothing is sized by a device->host sync).  A tile's records are
//   CONTIGUOUS, so the tile CTA stages them with ONE bulk-async copy (cp.async.bulk -> SASS UBLKCP)
//   completing on an mbarrier.
// * Triangles whose box spans more than 2x2 tiles ("large") go to one compact per-image list of
//   bounding boxes; every tile CTA scans that list with all its threads (16 B per entry) and builds the
//   records of the few entries that touch it.
// * raster_tiles, one CTA per tile, two phases per pass of <= 248 records:
//     phase 1  work item = one image row of one triangle (block-balanced).  The covered pixels of a
//              (triangle, row) are ONE interval whose ends are solved exactly (raster_core.cuh:
//              row_span_exact, ~one evaluation of the true edge function per edge instead of one test
//              per sample).  Each covered pixel records the triangle's slot in a per-pixel owner word
//              (four slot bytes, 32-bit shared-memory CAS): no depth is computed here.
//     phase 2  work item = pixel.  Every pixel shades its <= 4 owners (depth as the reference computes
//              it) and keeps the minimum key in a register: all lanes busy, no 64-bit atomics.
//              A pixel covered by more than four triangles of one pass (rare) shades the surplus in
//              phase 1 with a 64-bit shared-memory atomicMin, merged at the end.
//   Finally each thread writes index_img / depth_img for four adjacent pixels with 128-bit stores.
//   No global atomics, no memset of the images, no unpack pass.
#include <cstdlib>

#include "common.cuh"
#include "raster_core.cuh"
#include "tma.cuh"

namespace drtk {
// round-1 tile rasteriser (rasterize_v1.cu), selectable with DRTK_B200_RASTER_V1=1 for A/B runs
size_t rasterize_v1_workspace_bytes(int64_t N, int64_t F, int64_t H, int64_t W);
int rasterize_v1(const float* v, const int64_t* v_strides, const int32_t* vi, const int64_t* vi_strides, int64_t N,
                 int64_t V, int64_t F, int64_t H, int64_t W, float* depth_img, int32_t* index_img, void* workspace,
                 cudaStream_t stream);
namespace {

inline bool use_v1() {
  static const bool on = getenv("DRTK_B200_RASTER_V1") != nullptr;  // read once per process
  return on;
}

constexpr int kTileLog = 5;
constexpr int kTile = 1 << kTileLog;        // 32 x 32 pixels
constexpr int kTilePix = kTile * kTile;     // 1024
// tuning knobs (compile-time; tools/build_variants.sh builds A/B libraries with other values)
#ifndef RV_PASS
#define RV_PASS 120
#endif
#ifndef RV_MIN_CTAS
#define RV_MIN_CTAS 7
#endif
#ifndef RV_THREADS
#define RV_THREADS 128
#endif
constexpr int kRasterThreads = RV_THREADS;  // 128 or 256
constexpr int kPassRecs = RV_PASS;          // records per pass (<= 248, <= threads): slot ids are bytes, 0xFF = "no owner"
constexpr int kRecF4 = 5;                   // a record is 5 x 16 B
constexpr uint32_t kOwnEmpty = 0xFFFFFFFFu;
constexpr int kLargeBlock = 64;             // large-list entries examined per round of a tile CTA (narrow form)
constexpr int kScanPerThread = 4;           // ... and per thread in an optimistic wide round
constexpr int kScanBlock = kScanPerThread * kRasterThreads;
static_assert(kPassRecs <= kRasterThreads && kPassRecs <= 248 && kPassRecs > kLargeBlock, "pass size");

constexpr int kSuperLog = 3;                // a super-tile is 8 x 8 tiles (256 x 256 pixels): the bins of the "medium" triangles

struct RasterArgs {
  const float* v;
  Strides3 vs;
  const int32_t* vi;
  Strides3 vis;
  int N, V, F, H, W;
  int tilesX, tilesY;
  int superX, superY;  // super-tiles per image
};

// Triangle lists built by the bin kernels.  Three classes by the extent of the clamped bounding box:
//   small  (<= 2 x 2 tiles):        one 80-B RECORD per (triangle, tile) in the tile's own contiguous list;
//   medium (<= 2 x 2 super-tiles):  one (id, bbox) entry per (triangle, super-tile); a tile CTA scans the list of ITS
//                                   super-tile only (a few hundred boxes at most for any mesh that is not pathological);
//   large  (the rest):              one (id, bbox) entry in the per-image list every tile CTA scans; a triangle in it
//                                   spans more than 256 pixels, so an image holds few of them that are visible.
// (Round 2 first shipped small + large only: a 4 096^2 image of 50-px triangles then put 11 000 boxes in front of each of
// its 16 384 tile CTAs -- O(tiles x triangles).)
struct BinLists {
  uint32_t* tile_count;        // per tile: entries (count pass) / fill cursor (fill pass) / entries (tile kernel)
  const uint32_t* tile_offset; // per tile: first record of its list
  float4* recs;
  uint32_t* med_count;         // per super-tile, same protocol
  const uint32_t* med_offset;
  uint32_t* med_id;
  int4* med_bbox;
  uint32_t* large_count;       // per image
  uint32_t* large_id;
  int4* large_bbox;
};

// Everything a sample needs, derived once per triangle.
struct TriFull {
  EdgeSetup e;
  float d0, d1, d2;  // MUFU.RCP(epsclamp(z_k))
  float rden;        // MUFU.RCP(|den|)
  int bx0, by0, bx1, by1;  // clamped pixel bounding box (inclusive); may be empty
};

// Loads triangle f of image n, applies the reference's rejection rules (:81, :96-100, :107) and
// fills the setup.  Returns false when the triangle produces no samples.
__device__ __forceinline__ bool tri_full(const RasterArgs& a, int n, int f, TriFull& s) {
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

  edge_setup(i0, i1, i2, p0x, p0y, p1x, p1y, p2x, p2y, s.e);
  if (s.e.den == 0.f) return false;  // (:107)

  // bounding box with the reference's truncation and +1 border (:109-113)
  s.bx0 = max(0, __float2int_rz(mnx));
  s.by0 = max(0, __float2int_rz(mny));
  s.bx1 = min(a.W - 1, (int)((unsigned)__float2int_rz(mxx) + 1u));
  s.by1 = min(a.H - 1, (int)((unsigned)__float2int_rz(mxy) + 1u));

  s.rden = rcp_approx(fabsf(s.e.den));  // (:148) under fast-math: bary * MUFU.RCP(|den|)
  s.d0 = rcp_approx(epsclamp(z0));      // (:151)
  s.d1 = rcp_approx(epsclamp(z1));
  s.d2 = rcp_approx(epsclamp(z2));
  return true;
}

// depth bits of a COVERED sample (:148-153): c_k = b_k * RCP(|den|); inv = FFMA(c2,d2, FFMA(c0,d0, FMUL(c1,d1)));
// depth = MUFU.RCP(epsclamp(inv))
__device__ __forceinline__ uint32_t depth_bits_of(float b0, float b1, float b2, float rden, float d0, float d1,
                                                  float d2) {
  const float c0 = mul_rn(b0, rden), c1 = mul_rn(b1, rden), c2 = mul_rn(b2, rden);
  const float inv = fma_rn(c2, d2, fma_rn(c0, d0, mul_rn(c1, d1)));
  return __float_as_uint(rcp_approx(epsclamp(inv)));
}

// One (triangle, pixel) sample of the validation / wide-coordinate paths: coverage test + depth.
__device__ __forceinline__ bool sample(const TriFull& s, float px, const float (&row)[3], uint32_t& depth_bits) {
  if (!sample_covered(s.e.ox, s.e.ay, row, s.e.tl_bits, px)) return false;
  depth_bits = depth_bits_of(edge_value(s.e.ay[0], s.e.ox[0], row[0], px), edge_value(s.e.ay[1], s.e.ox[1], row[1], px),
                             edge_value(s.e.ay[2], s.e.ox[2], row[2], px), s.rden, s.d0, s.d1, s.d2);
  return true;
}

__device__ __forceinline__ void row_terms(const TriFull& s, float py, float (&row)[3]) {
#pragma unroll
  for (int k = 0; k < 3; ++k) row[k] = edge_row_term(py, s.e.oy[k], s.e.ax[k]);
}

// ------------------------------------------------------------------------------------------
// records
// ------------------------------------------------------------------------------------------
// float4 0..2: (ox_k, oy_k, ax_k, ay_k)   3: (d0, d1, d2, rden)   4: (triangle id, meta, 0, 0) as ints
// meta (raster_core.cuh): top/left bits, bounding box clipped to the tile and relative to it, WILD flag
__device__ __forceinline__ int record_meta(const TriFull& s, int x_lo, int y_lo, int x_hi, int y_hi, int W) {
  const int bx0 = max(s.bx0, x_lo) - x_lo, bx1 = min(s.bx1, x_hi) - x_lo;
  const int by0 = max(s.by0, y_lo) - y_lo, by1 = min(s.by1, y_hi) - y_lo;
  const bool tame = fabsf(s.e.ox[0]) < kSpanCoordMax && fabsf(s.e.ox[1]) < kSpanCoordMax &&
                    fabsf(s.e.ox[2]) < kSpanCoordMax && W <= (int)kSpanCoordMax;  // false for NaN
  return (int)s.e.tl_bits | (bx0 << RASTER_META_BX0) | (bx1 << RASTER_META_BX1) | (by0 << RASTER_META_BY0) |
         (by1 << RASTER_META_BY1) | (tame ? 0 : RASTER_META_WILD);
}

template <class Store>  // Store(int q, float4 value)
__device__ __forceinline__ void write_record(const TriFull& s, int f, int meta, Store&& st) {
#pragma unroll
  for (int k = 0; k < 3; ++k) st(k, make_float4(s.e.ox[k], s.e.oy[k], s.e.ax[k], s.e.ay[k]));
  st(3, make_float4(s.d0, s.d1, s.d2, s.rden));
  st(4, make_float4(__int_as_float(f), __int_as_float(meta), 0.f, 0.f));
}

// ------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------
// FILL = false: count the list entries of every tile.  FILL = true: append the records (small triangles) and the
// bounding boxes (large triangles).  Both passes take the same decisions from the same arithmetic.
template <bool FILL>
__global__ void __launch_bounds__(256) bin_kernel(RasterArgs a, int64_t total, BinLists L) {
  uint32_t* tile_count = L.tile_count;
  const uint32_t* __restrict__ tile_offset = L.tile_offset;
  float4* __restrict__ recs = L.recs;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // every lane of the warp stays in the kernel (the fill pass uses warp-wide MATCH / SHFL); `valid` carries the culling
  const bool in_range = idx < total;
  const int n = in_range ? (int)(idx / a.F) : 0;
  const int f = in_range ? (int)(idx - (int64_t)n * a.F) : 0;
  TriFull s;
  bool valid = in_range && tri_full(a, n, f, s);
  valid = valid && s.bx0 <= s.bx1 && s.by0 <= s.by1;
  const int tx0 = valid ? s.bx0 >> kTileLog : 0, tx1 = valid ? s.bx1 >> kTileLog : 0;
  const int ty0 = valid ? s.by0 >> kTileLog : 0, ty1 = valid ? s.by1 >> kTileLog : 0;
  const int64_t tbase = (int64_t)n * a.tilesX * a.tilesY;
  const bool small = valid && tx1 - tx0 <= 1 && ty1 - ty0 <= 1;
  // Neighbouring triangles of a mesh are neighbouring threads and fall into the same few tiles: in the fill pass the
  // lanes that append to the same tile are grouped (MATCH.ANY) and ONE of them takes a range of slots for all of them
  // -- a returning atomic per lane on the same counter serialises in L2 (the fill kernel ran at 27 % issue).
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int tx = tx0 + (q & 1), ty = ty0 + (q >> 1);
    const bool mine = small && tx <= tx1 && ty <= ty1;
    const int64_t t = tbase + (int64_t)ty * a.tilesX + tx;
    if (!FILL) {
      if (mine) atomicAdd(&tile_count[t], 1u);  // no return value: a fire-and-forget reduction
      continue;
    }
    // key: the tile for participating lanes (t < 2^31, host check), a private value for the others
    const unsigned peers = __match_any_sync(0xffffffffu, mine ? (unsigned)t : (0x80000000u | (unsigned)lane));
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (mine && lane == leader) base = atomicAdd(&tile_count[t], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (mine) {
      const uint32_t k = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
      const int x_lo = tx << kTileLog, y_lo = ty << kTileLog;
      const int meta = record_meta(s, x_lo, y_lo, min(x_lo + kTile - 1, a.W - 1), min(y_lo + kTile - 1, a.H - 1), a.W);
      float4* dst = recs + (size_t)(tile_offset[t] + k) * kRecF4;
      write_record(s, f, meta, [&](int q2, float4 val) { dst[q2] = val; });
    }
  }
  if (!valid || small) return;  // (after the last warp-wide operation)
  const int sx0 = tx0 >> kSuperLog, sx1 = tx1 >> kSuperLog, sy0 = ty0 >> kSuperLog, sy1 = ty1 >> kSuperLog;
  if (sx1 - sx0 <= 1 && sy1 - sy0 <= 1) {  // medium: an entry in each super-tile it touches
    const int64_t sbase = (int64_t)n * a.superX * a.superY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int sx = sx0 + (q & 1), sy = sy0 + (q >> 1);
      if (sx > sx1 || sy > sy1) continue;
      const int64_t st = sbase + (int64_t)sy * a.superX + sx;
      if (!FILL) {
        atomicAdd(&L.med_count[st], 1u);
      } else {
        const uint32_t pos = L.med_offset[st] + atomicAdd(&L.med_count[st], 1u);
        L.med_id[pos] = (uint32_t)f;
        L.med_bbox[pos] = make_int4(s.bx0, s.by0, s.bx1, s.by1);
      }
    }
  } else if (FILL) {  // large
    const uint32_t k = atomicAdd(&L.large_count[n], 1u);
    L.large_id[(int64_t)n * a.F + k] = (uint32_t)f;
    L.large_bbox[(int64_t)n * a.F + k] = make_int4(s.bx0, s.by0, s.bx1, s.by1);
  }
}

// Per-image exclusive scan of the tile counts (blockIdx.y == 0) and of the super-tile counts (blockIdx.y == 1) into
// list offsets, zeroing `count` so that bin_kernel<true> can reuse it as the fill cursor.  One CTA of 1024 threads per IMAGE (a small triangle adds at most four list
// entries, so image n owns the fixed list region [n * 4F, (n + 1) * 4F) and the images scan independently):
// T tiles in coalesced slabs of 4096 entries (uint4 per thread) carrying the running total -- one slab at
// config 4, four at config 5.  (A single CTA over all N * T counters took 13.6 us at config 4.)
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* count, uint32_t* offset, int64_t M, bool vec,
                                                    uint32_t* count2, uint32_t* offset2, int64_t M2, bool vec2,
                                                    uint32_t list_stride) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (blockIdx.y == 1) { count = count2; offset = offset2; M = M2; vec = vec2; }  // the super-tile (medium) lists
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
constexpr uint32_t kCellNone = 0xFFFFFFFFu;
constexpr int kRowsPerWarp = kTile / (kRasterThreads / 32);  // tile rows owned by one warp in phase 2 (4 or 8)

struct __align__(16) TileSmem {
  float4 rec[kPassRecs * kRecF4];          // the records of this pass (bulk-async copy / built in place)
  float4 srcp[kPassRecs];                  // per record: MUFU.RCP(ay_k), k = 0..2 (the span solver's only reciprocals)
  unsigned long long zbuf[kTilePix];       // packed minimum of the collision path (two spans starting on one pixel)
  uint32_t cell[kTilePix];                 // span starts: cell(ly, xs) = slot | xe << 8, kCellNone = no span starts here
  uint8_t rowmap[kPassRecs * kTile];       // row item -> record slot
  int prefix[kPassRecs + 1];               // exclusive scan of the per-record row counts
  int warp_tot[kRasterThreads / 32];
  uint32_t mlist[kPassRecs];               // large triangles that touch this tile (ids), current round
  int nmatch;
  int collided;                            // some span took the collision path: zbuf holds keys to merge at the end
  unsigned long long bar;                  // mbarrier of the record copy
};

// Span-start cells: row ly is rotated by ly columns.  Phase 1 has the lanes of a warp on consecutive ROWS of one triangle
// whose spans start at nearly the same column (one bank in a plain row-major layout); phase 2 reads one whole row per
// warp, for which any rotation is conflict free.
__device__ __forceinline__ int cell_slot(int ly, int lx) { return (ly << kTileLog) | ((lx + ly) & (kTile - 1)); }

// key of record `rec` at pixel (px, py): depth exactly as the reference computes it for a covered sample
__device__ __forceinline__ unsigned long long shade_key(const float4* __restrict__ rec, float px, float py) {
  const float4 e0 = rec[0], e1 = rec[1], e2 = rec[2], z = rec[3];
  const uint32_t id = __float_as_uint(rec[4].x);
  const float b0 = edge_value(e0.w, e0.x, edge_row_term(py, e0.y, e0.z), px);
  const float b1 = edge_value(e1.w, e1.x, edge_row_term(py, e1.y, e1.z), px);
  const float b2 = edge_value(e2.w, e2.x, edge_row_term(py, e2.y, e2.z), px);
  return ((unsigned long long)depth_bits_of(b0, b1, b2, z.w, z.x, z.y, z.z) << 32) | id;
}

// One pass over the m records in S.rec (m <= kPassRecs).  Entered and left by all threads of the CTA.
__device__ __forceinline__ void tile_pass(TileSmem& S, int m, int x_lo, int y_lo, float x_lo_f, float y_lo_f,
                                          unsigned long long (&best)[kRowsPerWarp]) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // ---- A. per record: reciprocals of the edge slopes; row items: exclusive block scan, row item -> slot map ----
  int rows = 0;
  if (tid < m) {
    const float4* rec = S.rec + tid * kRecF4;
    const int meta = __float_as_int(rec[4].y);
    rows = max(0, ((meta >> RASTER_META_BY1) & 31) - ((meta >> RASTER_META_BY0) & 31) + 1);
    S.srcp[tid] = make_float4(core_rcp(rec[0].w), core_rcp(rec[1].w), core_rcp(rec[2].w), 0.f);
  }
  int inc = rows;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) S.warp_tot[wid] = inc;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kRasterThreads / 32; ++w) {
    const int t = S.warp_tot[w];
    woff += (w < wid) ? t : 0;
    total += t;
  }
  const int excl = woff + inc - rows;
  if (tid < m) {
    S.prefix[tid] = excl;
    for (int r = 0; r < rows; ++r) S.rowmap[excl + r] = (uint8_t)tid;
  }
  __syncthreads();

  // ---- B. phase 1: one (triangle, row) per thread: exact span -> ONE span-start cell ----
  for (int item = tid; item < total; item += kRasterThreads) {
    const int slot = S.rowmap[item];
    const float4* rec = S.rec + slot * kRecF4;
    const float4 e0 = rec[0], e1 = rec[1], e2 = rec[2], rc = S.srcp[slot];
    const int meta = __float_as_int(rec[4].y);
    const int ly = ((meta >> RASTER_META_BY0) & 31) + (item - S.prefix[slot]);
    const float py = y_lo_f + small_i2f(ly);
    const float ox[3] = {e0.x, e1.x, e2.x}, ay[3] = {e0.w, e1.w, e2.w}, ray[3] = {rc.x, rc.y, rc.z};
    const float row[3] = {edge_row_term(py, e0.y, e0.z), edge_row_term(py, e1.y, e1.z), edge_row_term(py, e2.y, e2.z)};
    const unsigned tl = (unsigned)meta & RASTER_META_TL_MASK;
    const int lx0 = (meta >> RASTER_META_BX0) & 31, lx1 = (meta >> RASTER_META_BX1) & 31;
    int lxs, lxe;  // covered interval, tile-local
    if (!(meta & RASTER_META_WILD)) {
      float lo, hi;
      row_span_fast(ox, ay, row, ray, tl, x_lo_f + small_i2f(lx0), x_lo_f + small_i2f(lx1), lo, hi);
      lxs = small_f2i(lo - x_lo_f);
      lxe = small_f2i((hi - x_lo_f) + 1.f) - 1;  // hi may be lo0 - 1
    } else {  // huge coordinates: per-sample tests (the covered set is still one interval)
      lxs = kTile; lxe = -1;
      for (int lx = lx0; lx <= lx1; ++lx)
        if (sample_covered(ox, ay, row, tl, (float)(x_lo + lx))) { lxs = min(lxs, lx); lxe = max(lxe, lx); }
    }
    if (lxs > lxe) continue;
    // Register the span at its start cell.  If another span of this pass already starts on that pixel (overlapping
    // surfaces), shade that ONE pixel here (64-bit minimum) and register the rest of the span one pixel further right.
    const uint32_t mine = (uint32_t)slot | ((uint32_t)lxe << 8);
    while (lxs <= lxe && atomicCAS(&S.cell[cell_slot(ly, lxs)], kCellNone, mine) != kCellNone) {
      atomicMin(&S.zbuf[(ly << kTileLog) + lxs], shade_key(rec, x_lo_f + small_i2f(lxs), py));
      S.collided = 1;
      ++lxs;
    }
  }
  __syncthreads();

  // ---- C. phase 2: warp = tile row, lane = pixel: find the spans covering the pixel, shade, keep the minimum ----
  const float px = x_lo_f + small_i2f(lane);
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    const int ly = wid * kRowsPerWarp + r;
    const int cs = cell_slot(ly, lane);
    const uint32_t c = S.cell[cs];
    const bool is_start = c != kCellNone;
    const unsigned starts = __ballot_sync(0xffffffffu, is_start);
    if (starts == 0u) continue;            // nothing in this row (uniform)
    if (is_start) S.cell[cs] = kCellNone;  // for the next pass
    const float py = y_lo_f + small_i2f(ly);
    // nearest span start at or left of this pixel, and the one before it
    const unsigned left = starts & ((2u << lane) - 1u);
    const int s1 = left ? 31 - __clz(left) : lane;
    const uint32_t c1 = __shfl_sync(0xffffffffu, c, s1);
    const bool hit1 = left != 0u && lane <= (int)((c1 >> 8) & 31u);
    // Do spans of this row overlap (several surfaces)?  A start lane looks at the span that starts before it.
    const unsigned before = starts & ((1u << lane) - 1u);
    const int s0 = before ? 31 - __clz(before) : lane;
    const uint32_t c0 = __shfl_sync(0xffffffffu, c, s0);
    // running maximum of the span ends left of this start would be exact; the previous span's end is enough to
    // detect "some overlap in this row", which selects the general loop below
    const bool overlap = is_start && before != 0u && (int)((c0 >> 8) & 31u) >= lane;
    if (!__any_sync(0xffffffffu, overlap)) {  // one surface in this row: the nearest start is the only candidate
      if (hit1) {
        const unsigned long long key = shade_key(S.rec + (c1 & 0xFFu) * kRecF4, px, py);
        best[r] = key < best[r] ? key : best[r];
      }
      continue;
    }
    // Several surfaces: every start within the longest span length to the left may cover this pixel.  The covering
    // spans are first COLLECTED (up to four record slots per pixel, one byte each), then shaded together -- all lanes
    // shade their k-th span in the same step, whatever the order in which they found it.
    uint32_t hits = hit1 ? (0xFFFFFF00u | (c1 & 0xFFu)) : 0xFFFFFFFFu;
    {
      const int len = is_start ? (int)((c >> 8) & 31u) - lane + 1 : 0;
      const int L = __reduce_max_sync(0xffffffffu, len);
      const int first = max(lane - L + 1, 0);
      unsigned cand = left & ~((1u << first) - 1u);
      if (left) cand &= ~(1u << s1);  // already collected
      while (__any_sync(0xffffffffu, cand != 0u)) {
        const int s = cand ? 31 - __clz(cand) : lane;
        const uint32_t sc = __shfl_sync(0xffffffffu, c, s);
        const bool hit = cand != 0u && lane <= (int)((sc >> 8) & 31u);
        cand &= ~(1u << s);
        if (hit) {
          if ((hits >> 24) != 0xFFu) {  // a fifth span on this pixel (rare): make room by shading the oldest now
            const unsigned long long key = shade_key(S.rec + (hits >> 24) * kRecF4, px, py);
            best[r] = key < best[r] ? key : best[r];
          }
          hits = (hits << 8) | (sc & 0xFFu);
        }
      }
    }
    while (__any_sync(0xffffffffu, (hits & 0xFFu) != 0xFFu)) {
      if ((hits & 0xFFu) != 0xFFu) {
        const unsigned long long key = shade_key(S.rec + (hits & 0xFFu) * kRecF4, px, py);
        best[r] = key < best[r] ? key : best[r];
      }
      hits = (hits >> 8) | 0xFF000000u;
    }
  }
  __syncthreads();  // S.rec may be overwritten now
}

__global__ void __launch_bounds__(kRasterThreads, RV_MIN_CTAS) raster_tiles_kernel(
    RasterArgs a, const BinLists L, float* __restrict__ depth_img, int32_t* __restrict__ index_img) {
  const uint32_t* __restrict__ tile_count = L.tile_count;
  const uint32_t* __restrict__ tile_offset = L.tile_offset;
  const float4* __restrict__ recs = L.recs;
  __shared__ TileSmem S;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y, n = blockIdx.z;
  const int64_t t = ((int64_t)n * a.tilesY + tile_y) * a.tilesX + tile_x;
  const int x_lo = tile_x << kTileLog, y_lo = tile_y << kTileLog;
  const int x_hi = min(x_lo + kTile - 1, a.W - 1), y_hi = min(y_lo + kTile - 1, a.H - 1);
  const float x_lo_f = (float)x_lo, y_lo_f = (float)y_lo;
  uint64_t* bar = reinterpret_cast<uint64_t*>(&S.bar);

  const uint32_t cnt = a.F > 0 ? tile_count[t] : 0u;  // after bin_kernel<true>: entries of this tile
  const uint32_t off = a.F > 0 ? tile_offset[t] : 0u;
  const int64_t st = ((int64_t)n * a.superY + (tile_y >> kSuperLog)) * a.superX + (tile_x >> kSuperLog);
  const uint32_t nmed = a.F > 0 ? L.med_count[st] : 0u;
  const uint32_t med_off = a.F > 0 ? L.med_offset[st] : 0u;
  const uint32_t nlarge_img = a.F > 0 ? L.large_count[n] : 0u;
  if (cnt == 0u && nmed == 0u && nlarge_img == 0u) {
    // nothing can touch this tile (uniform over the CTA): write the background and leave -- an object in front of an
    // empty background leaves most tiles here, and the full tile prologue + resolve cost ~8 us per tile
    const int x = x_lo + lane;
    if (x <= x_hi) {
      for (int y = y_lo + wid; y <= y_hi; y += kRasterThreads / 32) {
        const int64_t o = ((int64_t)n * a.H + y) * a.W + x;
        index_img[o] = -1;
        depth_img[o] = 0.f;
      }
    }
    return;
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    S.nmatch = 0;
    S.collided = 0;
    if (cnt) {  // first pass in flight while the CTA initialises its pixel state
      const uint32_t bytes = min(cnt, (uint32_t)kPassRecs) * (uint32_t)(kRecF4 * 16);
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(S.rec, recs + (size_t)off * kRecF4, bytes, bar);
    }
  }
  for (int i = tid * 4; i < kTilePix; i += kRasterThreads * 4) {
    *reinterpret_cast<uint4*>(&S.cell[i]) = make_uint4(kCellNone, kCellNone, kCellNone, kCellNone);
    *reinterpret_cast<ulonglong2*>(&S.zbuf[i]) = make_ulonglong2(~0ull, ~0ull);
    *reinterpret_cast<ulonglong2*>(&S.zbuf[i + 2]) = make_ulonglong2(~0ull, ~0ull);
  }
  unsigned long long best[kRowsPerWarp];
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) best[r] = ~0ull;
  __syncthreads();  // barrier initialised, pixel state initialised

  // (1) small triangles: the tile's own record list, kPassRecs at a time
  uint32_t parity = 0;
  for (uint32_t base = 0; base < cnt; base += kPassRecs) {
    const int m = (int)min(cnt - base, (uint32_t)kPassRecs);
    if (base && tid == 0) {  // later passes (tiles with more than kPassRecs triangles)
      fence_proxy_async_smem();
      mbar_arrive_expect_tx(bar, (uint32_t)m * (kRecF4 * 16));
      bulk_g2s(S.rec, recs + (size_t)(off + base) * kRecF4, (uint32_t)m * (kRecF4 * 16), bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    tile_pass(S, m, x_lo, y_lo, x_lo_f, y_lo_f, best);
  }

  // (2) medium triangles (the list of this tile's super-tile), then large triangles (the list of the image): scan the
  // bounding boxes, build the records of those touching the tile.
  // The scan is optimistic: every thread tests kScanPerThread boxes per round (kScanBlock boxes between two barriers).
  // Matches are appended to S.mlist; should a wide round find more than the list can hold, it is undone, the matches
  // gathered so far are rasterised, and the round is redone in steps of kLargeBlock, which always fit.
  auto flush_matches = [&](int m) {  // uniform over the CTA, after a barrier; m <= kPassRecs
    if (tid < m) {
      TriFull s;
      float4* dst = S.rec + tid * kRecF4;
      if (tri_full(a, n, (int)S.mlist[tid], s) && max(s.bx0, x_lo) <= min(s.bx1, x_hi) && max(s.by0, y_lo) <= min(s.by1, y_hi)) {
        write_record(s, (int)S.mlist[tid], record_meta(s, x_lo, y_lo, x_hi, y_hi, a.W), [&](int q, float4 val) { dst[q] = val; });
      } else {  // cannot happen for a listed triangle; an empty record keeps the pass well defined
        dst[0] = dst[1] = dst[2] = dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[4] = make_float4(0.f, __int_as_float((1 << RASTER_META_BY0) | (0 << RASTER_META_BY1)), 0.f, 0.f);
      }
    }
    __syncthreads();
    tile_pass(S, m, x_lo, y_lo, x_lo_f, y_lo_f, best);
    if (tid == 0) S.nmatch = 0;
    __syncthreads();
  };
  auto scan_list = [&](const uint32_t* __restrict__ lid, const int4* __restrict__ lbb, uint32_t count) {
    auto test_append = [&](uint32_t idx, bool active) {  // convergent: called by whole warps
      bool hit = false;
      uint32_t f = 0;
      if (active) {
        const int4 bb = lbb[idx];
        hit = bb.x <= x_hi && bb.z >= x_lo && bb.y <= y_hi && bb.w >= y_lo;
        f = lid[idx];
      }
      const unsigned bm = __ballot_sync(0xffffffffu, hit);
      if (bm) {
        int wbase = 0;
        if (lane == 0) wbase = atomicAdd(&S.nmatch, __popc(bm));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        const int pos = wbase + __popc(bm & ((1u << lane) - 1u));
        if (hit && pos < kPassRecs) S.mlist[pos] = f;  // beyond the list: dropped, the round is redone (see below)
      }
    };
    // `pending` = matches gathered and not yet rasterised.  It mirrors S.nmatch at the start of every round but lives in
    // a register: S.nmatch is only ever READ between two barriers that no atomicAdd / reset can cross, so that all
    // threads see the same value and take the same flush / no-flush branch (the barriers inside the flush would fall
    // out of step otherwise -- racecheck found exactly that on a scene with 300 overlapping triangles).
    int pending = 0;
    for (uint32_t base = 0; base < count; base += kScanBlock) {
      const uint32_t end = min(base + (uint32_t)kScanBlock, count);
      const bool last = end == count;
      const int m0 = pending;
#pragma unroll
      for (int k = 0; k < kScanPerThread; ++k) {
        const uint32_t idx = base + (uint32_t)(k * kRasterThreads + tid);
        test_append(idx, idx < end);
      }
      __syncthreads();
      const int m = S.nmatch;
      __syncthreads();  // every thread has read the counter before anybody changes it again
      if (m > kPassRecs) {  // (uniform) the wide round overflowed the list: undo it
        if (tid == 0) S.nmatch = m0;
        __syncthreads();
        if (m0 > 0) flush_matches(m0);
        pending = 0;
        for (uint32_t b2 = base; b2 < end; b2 += kLargeBlock) {  // the always-fitting form: kLargeBlock boxes per barrier
          if (wid < kLargeBlock / 32) test_append(b2 + tid, b2 + tid < end);
          __syncthreads();
          const int m2 = S.nmatch;
          __syncthreads();
          pending = m2;
          if (m2 > kPassRecs - kLargeBlock || (last && b2 + kLargeBlock >= end && m2 > 0)) {
            flush_matches(m2);
            pending = 0;
          }
        }
      } else if (m > 0 && (last || m > kPassRecs - kLargeBlock)) {
        flush_matches(m);
        pending = 0;
      } else {
        pending = m;
      }
    }
  };
  scan_list(L.med_id + med_off, L.med_bbox + med_off, nmed);
  scan_list(L.large_id + (int64_t)n * a.F, L.large_bbox + (int64_t)n * a.F, nlarge_img);

  // (3) resolve + store (:402-415): empty -> index -1 (low word all ones), depth 0.  Lane = pixel of a row: every warp
  // store is one full 128-byte line of index_img / depth_img.
  const int x = x_lo + lane;
  if (x > x_hi) return;
  const int row0 = wid * kRowsPerWarp;
  const int64_t o0 = ((int64_t)n * a.H + (y_lo + row0)) * a.W + x;
  int32_t* ip = index_img + o0;
  float* dp = depth_img + o0;
  const int nrows = min(kRowsPerWarp, y_hi - (y_lo + row0) + 1);
  const bool merge = S.collided != 0;  // uniform over the CTA (every pass ends with a barrier)
#pragma unroll
  for (int r = 0; r < kRowsPerWarp; ++r) {
    if (r >= nrows) break;
    unsigned long long k = best[r];
    if (merge) {
      const unsigned long long zb = S.zbuf[((row0 + r) << kTileLog) + lane];
      k = zb < k ? zb : k;
    }
    const uint32_t d = (uint32_t)(k >> 32);
    *ip = (int)(uint32_t)k;
    *dp = d == 0xFFFFFFFFu ? 0.f : __uint_as_float(d);
    ip += a.W;
    dp += a.W;
  }
}

// ------------------------------------------------------------------------------------------
// validation path (algo 1): triangle-parallel walk with 64-bit global atomicMin + unpack.
// Same per-sample arithmetic, the reference's work organisation; used by the tests to
// cross-check the tiled path and as a second opinion against the oracle.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raster_atomic_kernel(RasterArgs a, int64_t total,
                                                            unsigned long long* packed_img) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx / a.F);
  const int f = (int)(idx - (int64_t)n * a.F);
  TriFull s;
  if (!tri_full(a, n, f, s)) return;
  unsigned long long* img = packed_img + (int64_t)n * a.H * a.W;
  for (int y = s.by0; y <= s.by1; ++y) {
    float row[3];
    row_terms(s, (float)y, row);
    for (int x = s.bx0; x <= s.bx1; ++x) {
      uint32_t db;
      if (sample(s, (float)x, row, db))
        atomicMin(img + (int64_t)y * a.W + x, ((unsigned long long)db << 32) | (uint32_t)f);
    }
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const unsigned long long* __restrict__ packed,
                                                     int64_t count, float* __restrict__ depth_img,
                                                     int32_t* __restrict__ index_img) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const unsigned long long z = packed[i];
  const uint32_t d = (uint32_t)(z >> 32);
  index_img[i] = (int)(uint32_t)z;
  depth_img[i] = d == 0xFFFFFFFFu ? 0.f : __uint_as_float(d);
}

// ------------------------------------------------------------------------------------------
// wireframe (src/rasterize/rasterize_kernel.cu:171-400 of the reference)
// ------------------------------------------------------------------------------------------
// A pixel belongs to the wireframe when one of the (visible) edges of a triangle crosses the diamond
// |dx| + |dy| = 0.5 around the pixel centre (:220-259); the triangle's interior still writes depth with id
// 0xFFFFFFFF (= -1) so that it occludes lines behind it (:376-393).  Edge visibility = bits 0-2 of the top
// nibble of vi[..., 0] (:293-303); bounding box padded by 2 and clamped to [1, W-2] x [1, H-2] (:333-337).
//
// Arithmetic: pinned with intrinsics to the reference build's compiled form (see below), like the fill path;
// tests compare index_img / depth_img bit for bit with the reference CUDA kernel.
// Organisation: one WARP per triangle, lanes stride over the padded bounding box (the reference walks it
// with one thread), 64-bit global atomicMin into the packed image, then unpack_kernel.
// The arithmetic of the diamond test is pinned to the reference build's sm_100 SASS (rasterize_lines_kernel
// <float,int>, all .FTZ), like the fill path: a knife-edge crossing point decides whether a pixel carries a
// triangle id, so the roundings must be the reference's (a version left to the compiler's own contraction
// choices differed in 2 of 2 M pixels at config 3).  For a triangle edge (p1, p2) with line (:171-181)
//   a1 = p1.y - p2.y,  b1 = p2.x - p1.x,  c1 = fma(p1.x, p2.y, -rn(p1.y * p2.x))
// and a diamond side (s0, s1) with a2 = s0.y - s1.y, b2 = s1.x - s0.x, c2 (see DiamondSides):
//   d  = fma(a1, b2, -rn(b1 * a2));  r = MUFU.RCP(d)
//   cx = rn(fma(b1, c2, -rn(c1 * b2)) * r);  cy = rn(fma(c1, a2, -rn(a1 * c2)) * r)       (:193-203)
//   d == 0  ->  (FLT_MAX, 0)
#ifndef DRTK_LINES_CHUNK
#define DRTK_LINES_CHUNK 0x40000000ull  // samples per pass of a triangle's padded box (a test build uses 96 to exercise the multi-pass path)
#endif
#ifndef DRTK_LINES_MINCTAS
#define DRTK_LINES_MINCTAS 3
#endif
struct Line { float a, b, c; };
__device__ __forceinline__ Line line_through(float p1x, float p1y, float p2x, float p2y) {
  Line l;
  l.a = sub_rn(p1y, p2y);
  l.b = sub_rn(p2x, p1x);
  l.c = fma_rn(p1x, p2y, -mul_rn(p1y, p2x));
  return l;
}
__device__ __forceinline__ bool within(float p1x, float p1y, float p2x, float p2y, float cx, float cy) {  // (:183-191)
  return (((p2x >= cx) && (cx >= p1x)) || ((p2x <= cx) && (cx <= p1x))) &&
         (((p2y >= cy) && (cy >= p1y)) || ((p2y <= cy) && (cy <= p1y)));
}
// the four sides of the diamond around (px, py), shared by the three edges of a triangle (:229-256):
// side k runs s0 -> s1 with  1: (px, py-.5) -> (px+.5, py)   2: (px+.5, py) -> (px, py+.5)
//                            3: (px, py+.5) -> (px-.5, py)   4: (px-.5, py) -> (px, py-.5)
// c2 = s0.x * s1.y - s1.x * s0.y shares the rounded product pp = rn(py * px) between the sides.
struct DiamondSides {
  float s0x[4], s0y[4], s1x[4], s1y[4], a2[4], b2[4], c2[4];
};
__device__ __forceinline__ void diamond_sides(float px, float py, DiamondSides& s) {
  const float xh = __fadd_rn(px, 0.5f), xl = __fadd_rn(px, -0.5f);
  const float yh = __fadd_rn(py, 0.5f), yl = __fadd_rn(py, -0.5f);
  const float pp = mul_rn(py, px);
  s.s0x[0] = px; s.s0y[0] = yl; s.s1x[0] = xh; s.s1y[0] = py;
  s.s0x[1] = xh; s.s0y[1] = py; s.s1x[1] = px; s.s1y[1] = yh;
  s.s0x[2] = px; s.s0y[2] = yh; s.s1x[2] = xl; s.s1y[2] = py;
  s.s0x[3] = xl; s.s0y[3] = py; s.s1x[3] = px; s.s1y[3] = yl;
  s.a2[0] = sub_rn(yl, py); s.b2[0] = sub_rn(xh, px); s.c2[0] = fma_rn(-yl, xh, pp);
  s.a2[1] = sub_rn(py, yh); s.b2[1] = sub_rn(px, xh); s.c2[1] = fma_rn(yh, xh, -pp);
  s.a2[2] = sub_rn(yh, py); s.b2[2] = sub_rn(xl, px); s.c2[2] = fma_rn(-yh, xl, pp);
  s.a2[3] = sub_rn(py, yl); s.b2[3] = sub_rn(px, xl); s.c2[3] = fma_rn(yl, xl, -pp);
}
__device__ __forceinline__ bool crosses_diamond(const Line& l, float p1x, float p1y, float p2x, float p2y,
                                                const DiamondSides& s) {
  bool hit = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float d = fma_rn(l.a, s.b2[k], -mul_rn(l.b, s.a2[k]));
    float cx = 3.402823466e+38f, cy = 0.f;  // TVec2{max}: x = FLT_MAX, y = 0 (:198)
    if (d != 0.f) {
      const float r = rcp_approx(d);
      cx = mul_rn(fma_rn(l.b, s.c2[k], -mul_rn(l.c, s.b2[k])), r);
      cy = mul_rn(fma_rn(l.c, s.a2[k], -mul_rn(l.a, s.c2[k])), r);
    }
    hit |= within(s.s0x[k], s.s0y[k], s.s1x[k], s.s1y[k], cx, cy) && within(p1x, p1y, p2x, p2y, cx, cy);
  }
  return hit;
}
// edge_function(a -> b, p) = v_ap.y * v_ab.x - v_ap.x * v_ab.y (:19-27) as the reference build rounds it in THIS
// kernel (read from the sm_100 SASS of rasterize_lines_kernel<float,int>): the product with the x-dependent
// factor is rounded and the other one fused, fma(v_ab.x, v_ap.y, -rn(v_ab.y * v_ap.x)) -- except for edge 0 in
// its canonical orientation, whose y product the compiler hoisted out of the x loop:
// fma(-v_ab.y, v_ap.x, rn(v_ap.y * v_ab.x)) (the form of the fill kernel).
__device__ __forceinline__ float edge_fn(float pax, float pay, float pbx, float pby, float px, float py) {
  const float abx = sub_rn(pbx, pax), aby = sub_rn(pby, pay);
  const float apx = sub_rn(px, pax), apy = sub_rn(py, pay);
  return fma_rn(abx, apy, -mul_rn(aby, apx));
}
__device__ __forceinline__ float edge_fn_hoisted(float pax, float pay, float pbx, float pby, float px, float py) {
  const float abx = sub_rn(pbx, pax), aby = sub_rn(pby, pay);
  const float apx = sub_rn(px, pax), apy = sub_rn(py, pay);
  return fma_rn(-aby, apx, mul_rn(apy, abx));
}
template <bool EDGE0>
__device__ __forceinline__ float canon_edge_fn(int ia, int ib, float pax, float pay, float pbx, float pby, float px,
                                               float py) {  // (:29-40)
  if (ia <= ib) return EDGE0 ? edge_fn_hoisted(pax, pay, pbx, pby, px, py) : edge_fn(pax, pay, pbx, pby, px, py);
  return -edge_fn(pbx, pby, pax, pay, px, py);
}

__global__ void __launch_bounds__(256, DRTK_LINES_MINCTAS) raster_lines_kernel(RasterArgs a, int64_t total,
                                                           unsigned long long* __restrict__ packed_img) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t idx = warp0; idx < total; idx += nwarps) {
    const int n = (int)(idx / a.F);
    const int f = (int)(idx - (int64_t)n * a.F);
    const int32_t* vip = a.vi + (int64_t)n * a.vis.s0 + (int64_t)f * a.vis.s1;
    const uint32_t raw0 = (uint32_t)vip[0];
    const int flag = (int)((raw0 & 0xF0000000u) >> 28);
    const int i0 = (int)(raw0 & 0x0FFFFFFFu), i1 = vip[a.vis.s2], i2 = vip[2 * a.vis.s2];
    if (i0 == i1 && i1 == i2) continue;  // (:296)
    const bool vis0 = (flag & 1) != 0, vis1 = (flag & 2) != 0, vis2 = (flag & 4) != 0;
    const float* vp = a.v + (int64_t)n * a.vs.s0;
    const float* q0 = vp + (int64_t)i0 * a.vs.s1;
    const float* q1 = vp + (int64_t)i1 * a.vs.s1;
    const float* q2 = vp + (int64_t)i2 * a.vs.s1;
    const float p0x = q0[0], p0y = q0[a.vs.s2], z0 = q0[2 * a.vs.s2];
    const float p1x = q1[0], p1y = q1[a.vs.s2], z1 = q1[2 * a.vs.s2];
    const float p2x = q2[0], p2y = q2[a.vs.s2], z2 = q2[2 * a.vs.s2];
    if (!(z0 > 1e-8f && z1 > 1e-8f && z2 > 1e-8f)) continue;  // (:321)
    const float mnx = fminf(fminf(p0x, p1x), p2x), mny = fminf(fminf(p0y, p1y), p2y);
    const float mxx = fmaxf(fmaxf(p0x, p1x), p2x), mxy = fmaxf(fmaxf(p0y, p1y), p2y);
    if (!(mnx <= (float)(a.W - 1) && mny <= (float)(a.H - 1) && mxx > 0.f && mxy > 0.f)) continue;  // (:322-323)
    const float v01x = sub_rn(p1x, p0x), v01y = sub_rn(p1y, p0y);
    const float v02x = sub_rn(p2x, p0x), v02y = sub_rn(p2y, p0y);
    const float v12x = sub_rn(p2x, p1x), v12y = sub_rn(p2y, p1y);
    const float den = diff_of_products(v01x, v02y, v01y, v02x);  // (:330)
    if (den == 0.f) continue;
    const int bx0 = max(1, (int)mnx - 2), by0 = max(1, (int)mny - 2);  // (:333-337)
    const int bx1 = min(a.W - 2, (int)mxx + 2), by1 = min(a.H - 2, (int)mxy + 2);
    if (bx0 > bx1 || by0 > by1) continue;
    const float sgn = den > 0.f ? 1.f : (den < 0.f ? -1.f : 0.f);
    bool tl0, tl1, tl2;  // (:361-369)
    if (den > 0.f) {
      tl0 = (v12y < 0.f) || (v12y == 0.f && v12x > 0.f);
      tl1 = (v02y > 0.f) || (v02y == 0.f && v02x < 0.f);
      tl2 = (v01y < 0.f) || (v01y == 0.f && v01x > 0.f);
    } else {
      tl0 = (v12y > 0.f) || (v12y == 0.f && v12x < 0.f);
      tl1 = (v02y < 0.f) || (v02y == 0.f && v02x > 0.f);
      tl2 = (v01y > 0.f) || (v01y == 0.f && v01x < 0.f);
    }
    const Line l01 = line_through(p0x, p0y, p1x, p1y), l12 = line_through(p1x, p1y, p2x, p2y),
               l02 = line_through(p0x, p0y, p2x, p2y);
    unsigned long long* img = packed_img + (int64_t)n * a.H * a.W;
    // Work split: the padded bounding box is walked in row-major order, 32 consecutive samples per step (a 9 x 9 box
    // of a 4-px triangle keeps 27 of 32 lanes busy; the previous (row, half-row) split kept 18).
    const int bw = bx1 - bx0 + 1, rows = by1 - by0 + 1;
    const uint64_t area = (uint64_t)bw * (uint64_t)rows;
    for (uint64_t s0 = 0; s0 < area; s0 += DRTK_LINES_CHUNK) {  // one pass unless the box holds > 2^30 samples: 32-bit index math inside
      const uint32_t cnt = (uint32_t)min((unsigned long long)(area - s0), (unsigned long long)DRTK_LINES_CHUNK);
      const uint32_t row0 = s0 ? (uint32_t)(s0 / (uint32_t)bw) : 0u;
      const uint32_t rem0 = s0 ? (uint32_t)(s0 - (uint64_t)row0 * (uint32_t)bw) : 0u;
      for (uint32_t s = (uint32_t)lane; s < cnt; s += 32u) {
      const uint32_t t = s + rem0;
      const uint32_t ry = t / (uint32_t)bw;
      const int y = by0 + (int)(row0 + ry), x = bx0 + (int)(t - ry * (uint32_t)bw);
      const float py = (float)y;
      {
      const float px = (float)x;
      DiamondSides ds;
      diamond_sides(px, py, ds);
      // An edge can only be hit through an intersection point c that lies `within` a diamond side AND `within` the
      // edge: c.x in [xl, xh] and in [min, max] of the edge's x (same for y).  When those closed ranges are disjoint
      // no value of c passes both tests, whatever the rounding of c -- the edge is skipped for this sample (exact,
      // not a tolerance).  xl / xh / yl / yh are the very values the sides are built from.
      const float xl = ds.s0x[3], xh = ds.s0x[1], yl = ds.s0y[0], yh = ds.s0y[2];
      auto reach = [&](float ax, float ay, float bx, float by) {
        return !(xh < fminf(ax, bx) || xl > fmaxf(ax, bx) || yh < fminf(ay, by) || yl > fmaxf(ay, by));
      };
      bool hit = false;  // (:343-346)
      if (vis0 && reach(p0x, p0y, p1x, p1y)) hit |= crosses_diamond(l01, p0x, p0y, p1x, p1y, ds);
      if (vis1 && reach(p1x, p1y, p2x, p2y)) hit |= crosses_diamond(l12, p1x, p1y, p2x, p2y, ds);
      if (vis2 && reach(p0x, p0y, p2x, p2y)) hit |= crosses_diamond(l02, p0x, p0y, p2x, p2y, ds);
      float b0 = canon_edge_fn<true>(i1, i2, p1x, p1y, p2x, p2y, px, py);  // (:348-353)
      float b1 = canon_edge_fn<false>(i2, i0, p2x, p2y, p0x, p0y, px, py);
      float b2 = canon_edge_fn<false>(i0, i1, p0x, p0y, p1x, p1y, px, py);
      b0 = mul_rn(b0, sgn); b1 = mul_rn(b1, sgn); b2 = mul_rn(b2, sgn);
      const bool inside = (b0 >= 0.f) && (b1 >= 0.f) && (b2 >= 0.f);
      const bool keep = inside && !((b0 == 0.f && !tl0) || (b1 == 0.f && !tl1) || (b2 == 0.f && !tl2));
      if (keep || hit) {  // (:375-393)
        // as compiled: FFMA.SAT(b, RCP(|den|), 0); (b0 + b1) + b2; b * RCP(sum); dot = FFMA(b2,d2, FFMA(b0,d0, FMUL(b1,d1)))
        const float rad = rcp_approx(fabsf(den));
        b0 = __saturatef(mul_rn(b0, rad)); b1 = __saturatef(mul_rn(b1, rad)); b2 = __saturatef(mul_rn(b2, rad));
        const float rs = rcp_approx(__fadd_rn(b2, __fadd_rn(b0, b1)));
        b0 = mul_rn(b0, rs); b1 = mul_rn(b1, rs); b2 = mul_rn(b2, rs);
        const float d0 = rcp_approx(epsclamp(z0)), d1 = rcp_approx(epsclamp(z1)), d2 = rcp_approx(epsclamp(z2));
        const float inv = fma_rn(b2, d2, fma_rn(b0, d0, mul_rn(b1, d1)));
        const float depth = rcp_approx(epsclamp(inv));
        const unsigned long long val = ((unsigned long long)__float_as_uint(depth) << 32) |
                                       (hit ? (unsigned long long)(uint32_t)f : 0xFFFFFFFFull);
        atomicMin(img + (int64_t)y * a.W + x, val);
      }
      }
      }
    }
  }
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct TiledWorkspace {
  size_t off_count, off_med_count, off_large_count, zero_bytes, off_offset, off_med_offset, off_med_id, off_med_bbox,
      off_large_id, off_large_bbox, off_recs, total;
  int64_t M;   // tiles of the batch
  int64_t MS;  // super-tiles of the batch
};

inline TiledWorkspace tiled_layout(int64_t N, int64_t F, int64_t H, int64_t W) {
  TiledWorkspace w;
  const int64_t tilesX = (W + kTile - 1) >> kTileLog, tilesY = (H + kTile - 1) >> kTileLog;
  const int64_t superX = (tilesX + (1 << kSuperLog) - 1) >> kSuperLog, superY = (tilesY + (1 << kSuperLog) - 1) >> kSuperLog;
  w.M = N * tilesX * tilesY;
  w.MS = N * superX * superY;
  size_t o = 0;
  w.off_count = o;       o = align_up(o + sizeof(uint32_t) * (size_t)w.M, 256);
  w.off_med_count = o;   o = align_up(o + sizeof(uint32_t) * (size_t)w.MS, 256);
  w.off_large_count = o; o = align_up(o + sizeof(uint32_t) * (size_t)N, 256);
  w.zero_bytes = o;      // [0, zero_bytes) is memset to 0 per call
  w.off_offset = o;      o = align_up(o + sizeof(uint32_t) * (size_t)w.M, 256);
  w.off_med_offset = o;  o = align_up(o + sizeof(uint32_t) * (size_t)w.MS, 256);
  w.off_med_id = o;      o = align_up(o + sizeof(uint32_t) * (size_t)(4 * N * F), 256);  // a medium triangle: <= 4 super-tiles
  w.off_med_bbox = o;    o = align_up(o + sizeof(int4) * (size_t)(4 * N * F), 256);
  w.off_large_id = o;    o = align_up(o + sizeof(uint32_t) * (size_t)(N * F), 256);
  w.off_large_bbox = o;  o = align_up(o + sizeof(int4) * (size_t)(N * F), 256);
  w.off_recs = o;        o = align_up(o + (size_t)(kRecF4 * 16) * (size_t)(4 * N * F), 256);
  w.total = o;
  return w;
}

}  // namespace
}  // namespace drtk

using namespace drtk;

extern "C" size_t drtk_b200_rasterize_workspace_bytes(int64_t N, int64_t F, int64_t H, int64_t W,
                                                       int algo) {
  if (N <= 0 || H <= 0 || W <= 0 || F < 0) return 0;
  if (N > kMaxBatchPerLaunch) N = kMaxBatchPerLaunch;  // larger batches are processed in slices that reuse the workspace
  if (algo == 1) return sizeof(unsigned long long) * (size_t)(N * H * W) + 256;
  const size_t lines = sizeof(unsigned long long) * (size_t)(N * H * W) + 256;  // wireframe mode uses the packed image
  size_t tiled = tiled_layout(N, F, H, W).total + 256;
  if (use_v1()) tiled = rasterize_v1_workspace_bytes(N, F, H, W);
  return tiled > lines ? tiled : lines;
}

extern "C" int drtk_b200_rasterize(const float* v, const int64_t* v_strides, const int32_t* vi,
                                   const int64_t* vi_strides, int64_t N, int64_t V, int64_t F,
                                   int64_t H, int64_t W, int wireframe, int algo, float* depth_img,
                                   int32_t* index_img, void* workspace, size_t workspace_bytes,
                                   void* stream_) {
  if (N < 0 || F < 0 || V < 0 || H <= 0 || W <= 0) return DRTK_B200_EINVAL;
  if (N == 0) return 0;
  if (!depth_img || !index_img || !v_strides || !vi_strides) return DRTK_B200_EINVAL;
  if ((F > 0 && (!v || !vi))) return DRTK_B200_EINVAL;
  if (N > kMaxBatchPerLaunch) {  // batch index rides on gridDim.z: slices of the batch (the workspace is reused)
    for (int64_t n0 = 0; n0 < N; n0 += kMaxBatchPerLaunch) {
      const int64_t nn = (N - n0 < kMaxBatchPerLaunch) ? N - n0 : kMaxBatchPerLaunch;
      const int rc = drtk_b200_rasterize(v ? v + n0 * v_strides[0] : nullptr, v_strides, vi ? vi + n0 * vi_strides[0] : nullptr,
                                         vi_strides, nn, V, F, H, W, wireframe, algo, depth_img + n0 * H * W,
                                         index_img + n0 * H * W, workspace, workspace_bytes, stream_);
      if (rc) return rc;
    }
    return 0;
  }
  if (H > (1 << 30) || W > (1 << 30) || N * F > (int64_t)0x1FFFFFFF || V >= 0x10000000LL)
    return DRTK_B200_EUNSUPPORTED;
  if (workspace_bytes < drtk_b200_rasterize_workspace_bytes(N, F, H, W, algo) || !workspace)
    return DRTK_B200_EWORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);

  RasterArgs a;
  a.v = v; a.vs = make3(v_strides); a.vi = vi; a.vis = make3(vi_strides);
  a.N = (int)N; a.V = (int)V; a.F = (int)F; a.H = (int)H; a.W = (int)W;
  a.tilesX = (int)((W + kTile - 1) >> kTileLog);
  a.tilesY = (int)((H + kTile - 1) >> kTileLog);
  a.superX = (a.tilesX + (1 << kSuperLog) - 1) >> kSuperLog;
  a.superY = (a.tilesY + (1 << kSuperLog) - 1) >> kSuperLog;
  const int64_t total = N * F;
  char* ws = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));

  if (wireframe) {
    unsigned long long* packed = reinterpret_cast<unsigned long long*>(ws);
    const int64_t npx = N * H * W;
    DRTK_CUDA(cudaMemsetAsync(packed, 0xFF, sizeof(unsigned long long) * (size_t)npx, stream));
    if (total > 0) {
      const int64_t want = (total * 32 + 255) / 256;
      const int64_t cap = (int64_t)num_sms() * 32;
      raster_lines_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(a, total, packed);
      DRTK_CHECK_LAUNCH();
    }
    unpack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, stream>>>(packed, npx, depth_img, index_img);
    DRTK_CHECK_LAUNCH();
    return 0;
  }

  if (algo == 1) {
    unsigned long long* packed = reinterpret_cast<unsigned long long*>(ws);
    const int64_t npx = N * H * W;
    DRTK_CUDA(cudaMemsetAsync(packed, 0xFF, sizeof(unsigned long long) * (size_t)npx, stream));
    if (total > 0) {
      raster_atomic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, total, packed);
      DRTK_CHECK_LAUNCH();
    }
    unpack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, stream>>>(packed, npx, depth_img, index_img);
    DRTK_CHECK_LAUNCH();
    return 0;
  }

  if (use_v1()) return rasterize_v1(v, v_strides, vi, vi_strides, N, V, F, H, W, depth_img, index_img, workspace, stream);
  const TiledWorkspace w = tiled_layout(N, F, H, W);
  if (w.M > 0x7FFFFFFFLL || a.tilesY > 65535) return DRTK_B200_EUNSUPPORTED;
  BinLists L;
  L.tile_count = reinterpret_cast<uint32_t*>(ws + w.off_count);
  L.tile_offset = reinterpret_cast<uint32_t*>(ws + w.off_offset);
  L.recs = reinterpret_cast<float4*>(ws + w.off_recs);
  L.med_count = reinterpret_cast<uint32_t*>(ws + w.off_med_count);
  L.med_offset = reinterpret_cast<uint32_t*>(ws + w.off_med_offset);
  L.med_id = reinterpret_cast<uint32_t*>(ws + w.off_med_id);
  L.med_bbox = reinterpret_cast<int4*>(ws + w.off_med_bbox);
  L.large_count = reinterpret_cast<uint32_t*>(ws + w.off_large_count);
  L.large_id = reinterpret_cast<uint32_t*>(ws + w.off_large_id);
  L.large_bbox = reinterpret_cast<int4*>(ws + w.off_large_bbox);

  DRTK_CUDA(cudaMemsetAsync(ws, 0, w.zero_bytes, stream));  // tile_count + med_count + large_count
  if (total > 0) {
    const unsigned blocks = (unsigned)((total + 255) / 256);
    bin_kernel<false><<<blocks, 256, 0, stream>>>(a, total, L);
    DRTK_CHECK_LAUNCH();
    const int64_t T = w.M / N, TS = w.MS / N;  // tiles / super-tiles per image; 16-B aligned per-image segments allow the uint4 path
    if (4 * F * N > 0xFFFFFFFFLL) return DRTK_B200_EUNSUPPORTED;
    scan_kernel<<<dim3((unsigned)N, 2u), 1024, 0, stream>>>(L.tile_count, const_cast<uint32_t*>(L.tile_offset), T, (T & 3) == 0,
                                                            L.med_count, const_cast<uint32_t*>(L.med_offset), TS, (TS & 3) == 0,
                                                            (uint32_t)(4 * F));
    DRTK_CHECK_LAUNCH();
    bin_kernel<true><<<blocks, 256, 0, stream>>>(a, total, L);
    DRTK_CHECK_LAUNCH();
  }
  raster_tiles_kernel<<<dim3((unsigned)a.tilesX, (unsigned)a.tilesY, (unsigned)N), kRasterThreads, 0, stream>>>(
      a, L, depth_img, index_img);
  DRTK_CHECK_LAUNCH();
  return 0;
}
