// raster_core.cuh -- the sample arithmetic of the rasteriser and the EXACT row-span solver built on it.
//
// Compiled twice: by nvcc into the kernels of rasterize.cu, and by g++ (-mfma, FTZ/DAZ set at run time) into
// tests/span_harness.cpp, which checks on the CPU -- millions of random, degenerate and knife-edge triangles --
// that the span solver selects exactly the samples the per-sample test selects.  The per-sample test is the
// reference's (src/rasterize/rasterize_kernel.cu:118-145 of facebookresearch/DRTK) in its compiled sm_100 form:
//
//     b_k(x, y) = FFMA(-ay_k, FADD(x, -ox_k), FMUL(FADD(y, -oy_k), ax_k))          (all .FTZ)
//     inside    = all b_k >= 0  and  every b_k == 0 belongs to a top/left edge
//
// (ox, oy) = origin of canonical edge k (the endpoint with the lower vertex index), (ax, ay) = edge direction
// times sign(den) * (swapped ? -1 : 1); see rasterize.cu for how they are derived.
//
// Exact spans.  FFMA rounds ONCE, so sign(b_k) = sign(row_k - ay_k * dx) exactly, with dx = rn(x - ox_k) monotone in
// x: along one image row edge k passes on a half line of pixels, and the covered pixels of a (triangle, row) are
// ONE interval.  Its ends are found without testing the samples in between: the real crossing x* = ox + row/ay
// is estimated with MUFU.RCP (error << 0.26 px under the magnitude guards below), which brackets the first / last
// passing pixel to two candidates, and ONE evaluation of the true b_k (same instructions as the per-sample test)
// decides between them.  Edges whose crossing is farther than 2^17 px from their origin (this includes ay == 0)
// have one sign over the whole row segment and are evaluated once.  Triangles with |coordinates| >= 2^16 or images
// wider than 2^16 px take the per-sample path instead (RASTER_META_WILD).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define DRTK_HD __host__ __device__ __forceinline__
#else
#define DRTK_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define DRTK_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define DRTK_MUL(a, b) __fmul_rn((a), (b))
#define DRTK_SUB(a, b) __fsub_rn((a), (b))
#define DRTK_ADD(a, b) __fadd_rn((a), (b))
#define DRTK_FLOOR(a) floorf(a)
namespace drtk {
__device__ __forceinline__ float core_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

}  // namespace drtk
#else
#include <math.h>
// host build (harness): g++ -O2 -mfma -ffp-contract=off, MXCSR FTZ|DAZ set by the caller
#define DRTK_FMA(a, b, c) fmaf((a), (b), (c))
#define DRTK_MUL(a, b) ((a) * (b))
#define DRTK_SUB(a, b) ((a) - (b))
#define DRTK_ADD(a, b) ((a) + (b))
#define DRTK_FLOOR(a) floorf(a)
namespace drtk {
// the harness perturbs the reciprocal by up to +-2 ulp to stand in for MUFU.RCP's approximation error
extern int g_rcp_ulp_noise;
inline float core_rcp(float x) {
  float r = 1.0f / x;
  if (g_rcp_ulp_noise && r == r && r != 0.f && fabsf(r) < 3.0e38f) {
    union { float f; int32_t i; } u;
    u.f = r;
    u.i += g_rcp_ulp_noise;
    r = u.f;
  }
  return r;
}

}  // namespace drtk
#endif

namespace drtk {

// tile-local meta word of a triangle record
constexpr int RASTER_META_TL_MASK = 7;         // bits 0-2: edge k is a top/left edge
constexpr int RASTER_META_BX0 = 3;             // 5 bits each: clipped bounding box relative to the tile
constexpr int RASTER_META_BX1 = 8;
constexpr int RASTER_META_BY0 = 13;
constexpr int RASTER_META_BY1 = 18;
constexpr int RASTER_META_WILD = 1 << 23;      // coordinates too large for the span solver: per-sample path

constexpr float kSpanNear = 131072.f;          // 2^17: |row/ay| beyond this -> one sign over the row segment
constexpr float kSpanCoordMax = 65536.f;       // 2^16: |ox|, image width allowed on the span path

// smallest normal float, negated: with FTZ arithmetic b is 0 or |b| >= FLT_MIN, so (b > -FLT_MIN) == (b >= 0)
#define DRTK_NEG_FLT_MIN (-1.17549435e-38f)

// Edge setup of one triangle (src/rasterize/rasterize_kernel.cu:29-40, :105-107, :120-141): canonical edges
// k = 0,1,2 <-> (v1,v2), (v2,v0), (v0,v1); origin = endpoint with the lower vertex index; (ax, ay) = direction
// times s = sign(den) * (swapped ? -1 : 1), so that b_k = s * fma(-ab.y, p.x-o.x, rn((p.y-o.y)*ab.x)) equals
// fma(-ay, p.x-o.x, rn((p.y-o.y)*ax)) bit for bit (round-to-nearest is sign symmetric).  Returns den
// (= FFMA(v01.x, v02.y, -FMUL(v01.y, v02.x)) as compiled); the caller rejects den == 0.
struct EdgeSetup {
  float ox[3], oy[3], ax[3], ay[3];
  unsigned tl_bits;
  float den;
};

DRTK_HD void canon_edge(int ia, int ib, float pax, float pay, float pbx, float pby, float sgn, float& ox, float& oy,
                        float& ax, float& ay) {
  if (ia <= ib) {
    ox = pax; oy = pay;
    ax = DRTK_MUL(DRTK_SUB(pbx, pax), sgn); ay = DRTK_MUL(DRTK_SUB(pby, pay), sgn);
  } else {
    ox = pbx; oy = pby;
    ax = DRTK_MUL(DRTK_SUB(pax, pbx), -sgn); ay = DRTK_MUL(DRTK_SUB(pay, pby), -sgn);
  }
}

DRTK_HD void edge_setup(int i0, int i1, int i2, float p0x, float p0y, float p1x, float p1y, float p2x, float p2y,
                        EdgeSetup& s) {
  const float v01x = DRTK_SUB(p1x, p0x), v01y = DRTK_SUB(p1y, p0y);
  const float v02x = DRTK_SUB(p2x, p0x), v02y = DRTK_SUB(p2y, p0y);
  const float v12x = DRTK_SUB(p2x, p1x), v12y = DRTK_SUB(p2y, p1y);
  s.den = DRTK_FMA(v01x, v02y, -DRTK_MUL(v01y, v02x));
  const float sgn = s.den > 0.f ? 1.f : -1.f;
  canon_edge(i1, i2, p1x, p1y, p2x, p2y, sgn, s.ox[0], s.oy[0], s.ax[0], s.ay[0]);
  canon_edge(i2, i0, p2x, p2y, p0x, p0y, sgn, s.ox[1], s.oy[1], s.ax[1], s.ay[1]);
  canon_edge(i0, i1, p0x, p0y, p1x, p1y, sgn, s.ox[2], s.oy[2], s.ax[2], s.ay[2]);
  bool t0, t1, t2;  // top-left classification (:133-141)
  if (s.den > 0.f) {
    t0 = (v12y < 0.f) || (v12y == 0.f && v12x > 0.f);
    t1 = (v02y > 0.f) || (v02y == 0.f && v02x < 0.f);
    t2 = (v01y < 0.f) || (v01y == 0.f && v01x > 0.f);
  } else {
    t0 = (v12y > 0.f) || (v12y == 0.f && v12x < 0.f);
    t1 = (v02y < 0.f) || (v02y == 0.f && v02x > 0.f);
    t2 = (v01y > 0.f) || (v01y == 0.f && v01x < 0.f);
  }
  s.tl_bits = (t0 ? 1u : 0u) | (t1 ? 2u : 0u) | (t2 ? 4u : 0u);
}

// row term of edge k: rn(rn(y - oy) * ax)   (hoisted per row by the reference compiler; same value either way)
DRTK_HD float edge_row_term(float py, float oy, float ax) { return DRTK_MUL(DRTK_SUB(py, oy), ax); }

// b_k at pixel column px
DRTK_HD float edge_value(float ay, float ox, float row, float px) { return DRTK_FMA(-ay, DRTK_SUB(px, ox), row); }

// The reference's per-sample decision (:127-145).
DRTK_HD bool sample_covered(const float (&ox)[3], const float (&ay)[3], const float (&row)[3], unsigned tl_bits,
                            float px) {
  const float b0 = edge_value(ay[0], ox[0], row[0], px);
  const float b1 = edge_value(ay[1], ox[1], row[1], px);
  const float b2 = edge_value(ay[2], ox[2], row[2], px);
  if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) return false;
  if ((b0 == 0.f && !(tl_bits & 1u)) || (b1 == 0.f && !(tl_bits & 2u)) || (b2 == 0.f && !(tl_bits & 4u))) return false;
  return true;
}

// Exact covered interval [xs, xe] of one (triangle, row) inside [xs0, xe0] (empty when xs > xe).
// Preconditions (else use the per-sample path): |ox_k| < 2^16, 0 <= xs0 <= xe0 < 2^16.
DRTK_HD void row_span_exact(const float (&ox)[3], const float (&ay)[3], const float (&row)[3], unsigned tl_bits,
                            int xs0, int xe0, int& xs, int& xe) {
  const float lo0 = (float)xs0, hi0 = (float)xe0;
  float lo = lo0, hi = hi0;
  const float cl = lo0 - 1.f, ch = hi0 + 1.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float thr = ((tl_bits >> k) & 1u) ? DRTK_NEG_FLT_MIN : 0.f;  // pass <=> b > thr
    const bool inc = ay[k] < 0.f;                                      // b grows with x
    const float rho = row[k] * core_rcp(ay[k]);                        // dx at which b changes sign
    const bool near = fabsf(rho) <= kSpanNear;                         // false for inf / NaN (ay == 0, ...)
    float u = DRTK_FLOOR((rho + ox[k]) + (inc ? 0.26f : 0.74f));
    u = near ? u : lo0;
    u = fminf(fmaxf(u, cl), ch);                                       // fmaxf(NaN, cl) = cl
    const bool p = edge_value(ay[k], ox[k], row[k], u) > thr;
    if (near) {
      if (inc) lo = fmaxf(lo, p ? u : u + 1.f);   // first passing column
      else hi = fminf(hi, p ? u : u - 1.f);       // last passing column
    } else if (!p) {
      hi = cl;                                    // this edge rejects the whole row segment
    }
  }
  xs = (int)lo;
  xe = (int)hi;
}

// int <-> float for small integers on the ALU / FMA pipes (I2F / F2I / FRND are XU instructions, the most loaded pipe
// of the tile kernel).  small_i2f: exact for 0 <= i < 2^23; small_f2i: exact for integer-valued |f| < 2^22.
#if defined(__CUDA_ARCH__)
#define DRTK_I2F_BITS(i) __int_as_float(i)
#define DRTK_F2I_BITS(f) __float_as_int(f)
#else
inline float drtk_bits_to_float(int32_t i) { union { int32_t i; float f; } u; u.i = i; return u.f; }
inline int32_t drtk_float_to_bits(float f) { union { int32_t i; float f; } u; u.f = f; return u.i; }
#define DRTK_I2F_BITS(i) drtk_bits_to_float(i)
#define DRTK_F2I_BITS(f) drtk_float_to_bits(f)
#endif
DRTK_HD float small_i2f(int i) { return DRTK_SUB(DRTK_I2F_BITS(0x4B000000 | i), 8388608.f); }
DRTK_HD int small_f2i(float f) { return DRTK_F2I_BITS(DRTK_ADD(f, 12582912.f)) - 0x4B400000; }

// row_span_exact as the tile kernel runs it: the three reciprocals ray[k] = MUFU.RCP(ay[k]) are supplied (computed once
// per record), floor() is done with the add-magic-number trick and the bounds stay in float (no XU instruction).
// Same decisions as row_span_exact; tests/span_harness.cpp checks both against the per-sample test.
// lo0 / hi0: the clipped bounding-box columns as floats; result: first / last covered column (empty when lo > hi).
DRTK_HD void row_span_fast(const float (&ox)[3], const float (&ay)[3], const float (&row)[3], const float (&ray)[3],
                           unsigned tl_bits, float lo0, float hi0, float& lo_out, float& hi_out) {
  float lo = lo0, hi = hi0;
  const float cl = DRTK_SUB(lo0, 1.f), ch = DRTK_ADD(hi0, 1.f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float thr = ((tl_bits >> k) & 1u) ? DRTK_NEG_FLT_MIN : 0.f;
    const bool inc = ay[k] < 0.f;
    const float rho = row[k] * ray[k];
    const bool near = fabsf(rho) <= kSpanNear;
    const float t = (rho + ox[k]) + (inc ? 0.26f : 0.74f);
    const float r = DRTK_SUB(DRTK_ADD(t, 12582912.f), 12582912.f);  // round to nearest integer (|t| < 2^22 when near)
    float u = r > t ? DRTK_SUB(r, 1.f) : r;                          // floor(t)
    u = near ? u : lo0;
    u = fminf(fmaxf(u, cl), ch);
    const bool p = edge_value(ay[k], ox[k], row[k], u) > thr;
    if (near) {
      if (inc) lo = fmaxf(lo, p ? u : DRTK_ADD(u, 1.f));
      else hi = fminf(hi, p ? u : DRTK_SUB(u, 1.f));
    } else if (!p) {
      hi = cl;
    }
  }
  lo_out = lo;
  hi_out = hi;
}

}  // namespace drtk
