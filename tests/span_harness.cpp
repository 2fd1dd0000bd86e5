// span_harness.cpp -- CPU check of the rasteriser's exact row-span solver (drtk_b200/csrc/raster_core.cuh).
//
// TEST INFRASTRUCTURE.  Built and run by tests/test_raster_span.py:
//     g++ -O2 -mfma -ffp-contract=off -I drtk_b200/csrc tests/span_harness.cpp -o <tmp>/span_harness
// For random / degenerate / knife-edge triangles and every image row of a tile-sized window it compares
//   (a) the interval returned by row_span_exact()            (what the CUDA kernel uses)
//   (b) the set of columns accepted by sample_covered()      (the reference's per-sample test)
// under FTZ/DAZ arithmetic with hardware FMA, the reciprocal perturbed by -2..+2 ulp (MUFU.RCP is an
// approximation).  Any difference is a bug in the solver.  Prints "rows=<n> covered=<c> mismatches=<m>".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <xmmintrin.h>
#include <pmmintrin.h>

#include "raster_core.cuh"

namespace drtk { int g_rcp_ulp_noise = 0; }
using namespace drtk;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() {  // xorshift64*
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return rng_state * 2685821657736338717ull;
}
static inline float uni() { return (float)((rnd() >> 40) * (1.0 / 16777216.0)); }  // [0,1)
static inline int irand(int n) { return (int)(rnd() % (uint64_t)n); }

struct Tri { float x[3], y[3]; int idx[3]; };

// families of test triangles; W = image width the coordinates live in
static Tri make_tri(int family, int W) {
  Tri t;
  for (int k = 0; k < 3; ++k) t.idx[k] = irand(1000);
  if (irand(8) == 0) t.idx[1] = t.idx[0];  // equal indices on an edge: canonical order falls back to "ia <= ib"
  const float cx = uni() * W, cy = uni() * W;
  switch (family) {
    case 0: {  // small, arbitrary (a few pixels)
      const float s = 0.5f + uni() * 12.f;
      for (int k = 0; k < 3; ++k) { t.x[k] = cx + (uni() - 0.5f) * s; t.y[k] = cy + (uni() - 0.5f) * s; }
      break;
    }
    case 1: {  // medium / large
      const float s = 20.f + uni() * 400.f;
      for (int k = 0; k < 3; ++k) { t.x[k] = cx + (uni() - 0.5f) * s; t.y[k] = cy + (uni() - 0.5f) * s; }
      break;
    }
    case 2: {  // vertices on pixel centres / half pixels: exact ties, horizontal and vertical edges
      const int s = 1 + irand(20);
      for (int k = 0; k < 3; ++k) {
        t.x[k] = (float)((int)cx + irand(2 * s + 1) - s) + (irand(4) == 0 ? 0.5f : 0.f);
        t.y[k] = (float)((int)cy + irand(2 * s + 1) - s) + (irand(4) == 0 ? 0.5f : 0.f);
      }
      break;
    }
    case 3: {  // slivers: nearly collinear, nearly horizontal / vertical edges
      const float s = 1.f + uni() * 60.f;
      const float dx = (uni() - 0.5f) * s, dy = (uni() - 0.5f) * s * (irand(2) ? 1e-3f : 1.f);
      t.x[0] = cx; t.y[0] = cy;
      t.x[1] = cx + dx; t.y[1] = cy + dy;
      const float a = uni() * 2.f;
      t.x[2] = cx + a * dx + (uni() - 0.5f) * 1e-2f * (irand(3) ? 1.f : 100.f);
      t.y[2] = cy + a * dy + (uni() - 0.5f) * 1e-2f * (irand(3) ? 1.f : 100.f);
      break;
    }
    case 4: {  // tiny edge slopes: ay of magnitude 1e-6 .. 1e-30 (far crossings), huge triangles
      const float s = 10.f + uni() * 1000.f;
      const float eps = powf(10.f, -(1.f + uni() * 30.f));
      t.x[0] = cx - s; t.y[0] = cy;
      t.x[1] = cx + s; t.y[1] = cy + (irand(2) ? eps : -eps) * (irand(2) ? 1.f : 0.f);
      t.x[2] = cx + (uni() - 0.5f) * s; t.y[2] = cy + (uni() - 0.5f) * s;
      break;
    }
    default: {  // shared-edge pairs are covered by family 0..3 through the canonical form; here: off-screen / partially visible
      const float s = 5.f + uni() * 200.f;
      for (int k = 0; k < 3; ++k) { t.x[k] = (uni() * 1.4f - 0.2f) * W + (uni() - 0.5f) * s; t.y[k] = (uni() * 1.4f - 0.2f) * W + (uni() - 0.5f) * s; }
      break;
    }
  }
  return t;
}

int main(int argc, char** argv) {
  _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
  _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
  const long ntri = argc > 1 ? atol(argv[1]) : 2000000;
  const int W = argc > 2 ? atoi(argv[2]) : 2048;
  rng_state ^= (uint64_t)(argc > 3 ? atol(argv[3]) : 1) * 0xD1B54A32D192ED03ull;
  long rows = 0, covered = 0, mism = 0, printed = 0;
  for (long it = 0; it < ntri; ++it) {
    const Tri t = make_tri((int)(it % 6), W);
    EdgeSetup s;
    edge_setup(t.idx[0], t.idx[1], t.idx[2], t.x[0], t.y[0], t.x[1], t.y[1], t.x[2], t.y[2], s);
    if (s.den == 0.f) continue;
    // the reference's bounding box (:109-113), then a random 32-wide tile window intersecting it
    float mnx = fminf(fminf(t.x[0], t.x[1]), t.x[2]), mxx = fmaxf(fmaxf(t.x[0], t.x[1]), t.x[2]);
    float mny = fminf(fminf(t.y[0], t.y[1]), t.y[2]), mxy = fmaxf(fmaxf(t.y[0], t.y[1]), t.y[2]);
    if (!(mnx <= (float)(W - 1) && mny <= (float)(W - 1) && mxx > 0.f && mxy > 0.f)) continue;
    int bx0 = (int)mnx < 0 ? 0 : (int)mnx, by0 = (int)mny < 0 ? 0 : (int)mny;
    int bx1 = (int)mxx + 1 > W - 1 ? W - 1 : (int)mxx + 1, by1 = (int)mxy + 1 > W - 1 ? W - 1 : (int)mxy + 1;
    if (bx0 > bx1 || by0 > by1) continue;
    const int tx = (bx0 + irand(bx1 - bx0 + 1)) >> 5, ty = (by0 + irand(by1 - by0 + 1)) >> 5;
    const int xs0 = bx0 > tx * 32 ? bx0 : tx * 32, xe0 = bx1 < tx * 32 + 31 ? bx1 : tx * 32 + 31;
    const int ys0 = by0 > ty * 32 ? by0 : ty * 32, ye0 = by1 < ty * 32 + 31 ? by1 : ty * 32 + 31;
    g_rcp_ulp_noise = irand(5) - 2;
    for (int y = ys0; y <= ye0; ++y) {
      float row[3];
      for (int k = 0; k < 3; ++k) row[k] = edge_row_term((float)y, s.oy[k], s.ax[k]);
      int xs, xe;
      row_span_exact(s.ox, s.ay, row, s.tl_bits, xs0, xe0, xs, xe);
      // the tile kernel's variant (supplied reciprocals, float bounds, no XU instructions) must agree
      {
        const float ray[3] = {core_rcp(s.ay[0]), core_rcp(s.ay[1]), core_rcp(s.ay[2])};
        const float x_lo_f = (float)(tx * 32);
        float lo, hi;
        row_span_fast(s.ox, s.ay, row, ray, s.tl_bits, x_lo_f + small_i2f(xs0 - tx * 32), x_lo_f + small_i2f(xe0 - tx * 32), lo, hi);
        const int fxs = tx * 32 + small_f2i(lo - x_lo_f), fxe = tx * 32 + small_f2i((hi - x_lo_f) + 1.f) - 1;
        const bool e1 = xs > xe, e2 = fxs > fxe;
        if (e1 != e2 || (!e1 && (fxs != xs || fxe != xe))) { ++mism; if (printed++ < 10) fprintf(stderr, "FAST variant differs: [%d,%d] vs [%d,%d]\n", fxs, fxe, xs, xe); }
      }
      bool bad = false;
      int n_in = 0;
      for (int x = xs0; x <= xe0; ++x) {
        const bool in = sample_covered(s.ox, s.ay, row, s.tl_bits, (float)x);
        n_in += in;
        if (in != (x >= xs && x <= xe)) bad = true;
      }
      ++rows;
      covered += n_in;
      if (bad) {
        ++mism;
        if (printed++ < 10) {
          fprintf(stderr, "MISMATCH family %ld y=%d window [%d,%d] span [%d,%d] brute:", it % 6, y, xs0, xe0, xs, xe);
          for (int x = xs0; x <= xe0; ++x) fputc(sample_covered(s.ox, s.ay, row, s.tl_bits, (float)x) ? '#' : '.', stderr);
          fprintf(stderr, "\n  v = (%.9g,%.9g) (%.9g,%.9g) (%.9g,%.9g) idx %d %d %d noise %d\n", t.x[0], t.y[0], t.x[1], t.y[1],
                  t.x[2], t.y[2], t.idx[0], t.idx[1], t.idx[2], g_rcp_ulp_noise);
        }
      }
    }
  }
  printf("rows=%ld covered=%ld mismatches=%ld\n", rows, covered, mism);
  return mism ? 1 : 0;
}
