/*
 * drtk_oracle_impl.h -- CPU restatement of the DRTK rasterisation hot path.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under drtk_b200/ may include, link or call this.
 * It exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check
 * the CUDA kernels.  Included twice by drtk_oracle.c with REAL = float / double.
 *
 * Every function restates (in plain C, contiguous layouts, sequential accumulation) the
 * algorithm of the reference file:line it cites.  Paths are relative to /root/reference.
 *
 * Parity pinning: tests/test_oracle_golden.py checks these functions against fixtures in
 * tests/golden/ that were produced by the reference's own CPU kernels (oracle/_ref,
 * built from the unmodified reference sources by oracle/build_ref.py) and against the
 * known answers listed in SURVEY.md section 4.
 */

#ifndef REAL
#error "include from drtk_oracle.c"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

/* src/include/cuda_math_helper.h:1036-1041 (eps :63-69): keep |v| >= eps, preserve sign,
 * non-negative (and NaN, for which v<0 is false) go to the +eps branch via max().       */
static inline REAL FN(epsclamp)(REAL v) {
  const REAL eps = EPSVAL;
  if (v < 0) return v < -eps ? v : -eps; /* min(v,-eps) */
  return v > eps ? v : eps;              /* max(v, eps) */
}

/* src/include/cuda_math_helper.h:994 */
static inline REAL FN(sgn)(REAL v) { return v > 0 ? (REAL)1 : (v < 0 ? (REAL)-1 : (REAL)0); }

/* ------------------------------------------------------------------------------------ */
/* Edge functions.                                                                       */
/* mode 0: reference CPU twin arithmetic -- two rounded products, one rounded subtraction */
/*         (src/rasterize/rasterize_kernel_cpu.cpp:32-36).                                */
/* mode 1: reference CUDA arithmetic as compiled with --use_fast_math for sm_100          */
/*         (src/rasterize/rasterize_kernel.cu:19-27; SASS read with cuobjdump from         */
/*         oracle/_ref/rasterize_ext.so): t = rn(v_ap.y * v_ab.x); e = fma(-v_ab.y, v_ap.x, t) */
/* ------------------------------------------------------------------------------------ */
static inline REAL FN(edge_fn)(REAL ax, REAL ay, REAL bx, REAL by, REAL px, REAL py, int mode) {
  const REAL abx = bx - ax, aby = by - ay;
  const REAL apx = px - ax, apy = py - ay;
  if (mode == 1) {
    const volatile REAL t = apy * abx;
    return FMA(-apx, aby, t);
  } else {
    const volatile REAL t1 = apy * abx;
    const volatile REAL t2 = apx * aby;
    return t1 - t2;
  }
}

/* src/rasterize/rasterize_kernel.cu:29-40: evaluate the shared edge from the lower vertex
 * *index* to the higher one so both triangles sharing it see the same magnitude.          */
static inline REAL FN(canon_edge)(int32_t ia, int32_t ib, REAL ax, REAL ay, REAL bx, REAL by,
                                  REAL px, REAL py, int mode) {
  if (ia <= ib) return FN(edge_fn)(ax, ay, bx, by, px, py, mode);
  return -FN(edge_fn)(bx, by, ax, ay, px, py, mode);
}

/* top-left classification of the three edges; src/rasterize/rasterize_kernel.cu:133-141 */
static inline void FN(top_left)(REAL den, REAL v01x, REAL v01y, REAL v02x, REAL v02y, REAL v12x,
                                REAL v12y, int tl[3]) {
  if (den > 0) {
    tl[0] = (v12y < 0) || (v12y == 0 && v12x > 0);
    tl[1] = (v02y > 0) || (v02y == 0 && v02x < 0);
    tl[2] = (v01y < 0) || (v01y == 0 && v01x > 0);
  } else {
    tl[0] = (v12y > 0) || (v12y == 0 && v12x < 0);
    tl[1] = (v02y < 0) || (v02y == 0 && v02x > 0);
    tl[2] = (v01y > 0) || (v01y == 0 && v01x < 0);
  }
}

static inline int FN(f2i_trunc)(REAL f) {
  /* C `int(f)`; mirror cvt.rzi saturation so huge coordinates do not invoke UB */
  if (!(f == f)) return 0;
  if (f >= (REAL)2147483647.0) return 2147483647;
  if (f <= (REAL)-2147483648.0) return (int)(-2147483647 - 1);
  return (int)f;
}

/*
 * rasterize: z-buffered coverage.  Restates rasterize_kernel + unpack_kernel + memset
 * (src/rasterize/rasterize_kernel.cu:42-168, :402-415, :484-488) and the CPU twin
 * (src/rasterize/rasterize_kernel_cpu.cpp:54-204).
 *
 *  v          [N,V,3] REAL contiguous;  vi [N or 1, F, 3] int32 (vi_batched selects)
 *  depth_img  [N,H,W] float (always float, reference :481); index_img [N,H,W] int32
 *  margin_img optional [N,H,W] uint32: ulp distance (float bits) between the winning and
 *             the runner-up depth at that pixel, 0xFFFFFFFF when fewer than 2 candidates.
 *             Tests use it to mask pixels where approximate reciprocals (MUFU.RCP on the
 *             GPU, not reproducible on a CPU) can legitimately flip the z-test.
 *  mode       0 = CPU-twin arithmetic, 1 = CUDA arithmetic (FMA edge function, reciprocal
 *             multiply, fma-chained depth); reciprocals are IEEE 1/x in both.
 */
void FN(oracle_rasterize)(const REAL* v, const int32_t* vi, int64_t N, int64_t V, int64_t F,
                          int64_t H, int64_t W, int vi_batched, int mode, float* depth_img,
                          int32_t* index_img, uint32_t* margin_img) {
  (void)V;
  const int64_t HW = H * W;
  for (int64_t n = 0; n < N; ++n) {
    uint64_t* best = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)HW);
    uint64_t* second = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)HW);
    memset(best, 0xFF, sizeof(uint64_t) * (size_t)HW); /* :484-488 */
    memset(second, 0xFF, sizeof(uint64_t) * (size_t)HW);
    const REAL* vn = v + n * V * 3;
    const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
    for (int64_t id = 0; id < F; ++id) {
      const int32_t i0 = (int32_t)(((uint32_t)vin[id * 3 + 0]) & 0x0FFFFFFFu); /* :74 */
      const int32_t i1 = vin[id * 3 + 1];
      const int32_t i2 = vin[id * 3 + 2];
      if (i0 == i1 && i1 == i2) continue; /* :81 */
      const REAL p0x = vn[i0 * 3 + 0], p0y = vn[i0 * 3 + 1], z0 = vn[i0 * 3 + 2];
      const REAL p1x = vn[i1 * 3 + 0], p1y = vn[i1 * 3 + 1], z1 = vn[i1 * 3 + 2];
      const REAL p2x = vn[i2 * 3 + 0], p2y = vn[i2 * 3 + 1], z2 = vn[i2 * 3 + 2];
      if (!(z0 > (REAL)1e-8f && z1 > (REAL)1e-8f && z2 > (REAL)1e-8f)) continue; /* :96 */
      const REAL mnx = FMIN(FMIN(p0x, p1x), p2x), mny = FMIN(FMIN(p0y, p1y), p2y);
      const REAL mxx = FMAX(FMAX(p0x, p1x), p2x), mxy = FMAX(FMAX(p0y, p1y), p2y);
      if (!(mnx <= (REAL)(W - 1) && mny <= (REAL)(H - 1) && mxx > 0 && mxy > 0)) continue; /* :97-98 */
      const REAL v01x = p1x - p0x, v01y = p1y - p0y;
      const REAL v02x = p2x - p0x, v02y = p2y - p0y;
      const REAL v12x = p2x - p1x, v12y = p2y - p1y;
      REAL den;
      if (mode == 1) {
        /* sm_100 SASS of the reference: FMUL t = v01.y*v02.x ; FFMA den = v01.x*v02.y - t */
        const volatile REAL t = v01y * v02x;
        den = FMA(v01x, v02y, -t);
      } else {
        const volatile REAL t1 = v01x * v02y;
        const volatile REAL t2 = v01y * v02x;
        den = t1 - t2;
      }
      if (den == 0) continue; /* :107 */
      int bx0 = FN(f2i_trunc)(mnx); if (bx0 < 0) bx0 = 0;           /* :109-113 */
      int by0 = FN(f2i_trunc)(mny); if (by0 < 0) by0 = 0;
      int64_t bx1 = (int64_t)FN(f2i_trunc)(mxx) + 1; if (bx1 > W - 1) bx1 = W - 1;
      int64_t by1 = (int64_t)FN(f2i_trunc)(mxy) + 1; if (by1 > H - 1) by1 = H - 1;
      int tl[3];
      FN(top_left)(den, v01x, v01y, v02x, v02y, v12x, v12y, tl);
      const REAL s = FN(sgn)(den);
      const REAL aden = den < 0 ? -den : den;
      const REAL rden = (REAL)1 / aden;
      const REAL d0 = (REAL)1 / FN(epsclamp)(z0), d1 = (REAL)1 / FN(epsclamp)(z1),
                 d2 = (REAL)1 / FN(epsclamp)(z2);
      for (int64_t y = by0; y <= by1; ++y) {
        for (int64_t x = bx0; x <= bx1; ++x) {
          const REAL px = (REAL)x, py = (REAL)y;
          REAL b0 = FN(canon_edge)(i1, i2, p1x, p1y, p2x, p2y, px, py, mode) * s; /* :120-125 */
          REAL b1 = FN(canon_edge)(i2, i0, p2x, p2y, p0x, p0y, px, py, mode) * s;
          REAL b2 = FN(canon_edge)(i0, i1, p0x, p0y, p1x, p1y, px, py, mode) * s;
          if (!(b0 >= 0 && b1 >= 0 && b2 >= 0)) continue; /* :127 */
          if ((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2])) continue; /* :143-145 */
          REAL depth;
          if (mode == 1) {
            b0 = b0 * rden; b1 = b1 * rden; b2 = b2 * rden;
            /* sm_100 SASS of the reference: FMUL t = b1*d1 ; FFMA (b0,d0,t) ; FFMA (b2,d2,.) */
            const volatile REAL t = b1 * d1;
            const REAL inv = FMA(b2, d2, FMA(b0, d0, t));
            depth = (REAL)1 / FN(epsclamp)(inv);
          } else {
            b0 = b0 / aden; b1 = b1 / aden; b2 = b2 / aden; /* cpu :176-178 */
            const volatile REAL t0 = d0 * b0;
            const volatile REAL t1 = d1 * b1;
            const volatile REAL t2 = d2 * b2;
            const volatile REAL s01 = t0 + t1;
            const REAL inv = s01 + t2; /* cpu :181 */
            depth = (REAL)1 / FN(epsclamp)(inv);
          }
          const float depth_f = (float)depth;
          uint32_t dbits; memcpy(&dbits, &depth_f, 4);
          const uint64_t packed = ((uint64_t)dbits << 32) | (uint64_t)(uint32_t)id; /* :155-157 */
          const int64_t o = y * W + x;
          if (packed < best[o]) { second[o] = best[o]; best[o] = packed; }
          else if (packed < second[o]) { second[o] = packed; }
        }
      }
    }
    for (int64_t o = 0; o < HW; ++o) { /* unpack :409-413 */
      const uint32_t du = (uint32_t)(best[o] >> 32);
      float df; memcpy(&df, &du, 4);
      depth_img[n * HW + o] = (du == 0xFFFFFFFFu) ? 0.0f : df;
      index_img[n * HW + o] = (int32_t)(uint32_t)(best[o] & 0xFFFFFFFFu);
      if (margin_img) {
        const uint32_t d2u = (uint32_t)(second[o] >> 32);
        margin_img[n * HW + o] = (d2u == 0xFFFFFFFFu) ? 0xFFFFFFFFu : (d2u - du);
      }
    }
    free(best); free(second);
  }
}

/*
 * rasterize, wireframe mode.  Restates rasterize_lines_kernel + unpack_kernel + memset
 * (src/rasterize/rasterize_kernel.cu:171-400, :402-415, :484-488).  The reference has NO CPU twin for this
 * mode (rasterize_kernel_cpu.cpp:257 raises), so this restatement follows the CUDA kernel -- with the FMA
 * contractions of the reference's sm_100 build (noted per expression) and IEEE 1/x where the GPU uses
 * MUFU.RCP -- and is pinned against tests/golden/wire_*.npz, which were written by the reference CUDA
 * kernel on a B200 (tests/golden/make_golden_wireframe.py).  A pixel whose crossing point lands within
 * rounding of a segment end may legitimately differ from the GPU (approximate reciprocal); the tests bound
 * the number of such pixels and compare depths to a few ulp.
 */
static inline int FN(within)(REAL p1x, REAL p1y, REAL p2x, REAL p2y, REAL cx, REAL cy) { /* :183-191 */
  return (((p2x >= cx) && (cx >= p1x)) || ((p2x <= cx) && (cx <= p1x))) &&
         (((p2y >= cy) && (cy >= p1y)) || ((p2y <= cy) && (cy <= p1y)));
}

/* is_crossing_dimond (:220-259) for the edge (p1, p2) with line a1 x + b1 y + c1 = 0 (:171-181) */
static inline int FN(crosses_diamond)(REAL a1, REAL b1, REAL c1, REAL p1x, REAL p1y, REAL p2x, REAL p2y,
                                      REAL px, REAL py) {
  const REAL h = (REAL)0.5;
  const REAL xh = px + h, xl = px - h, yh = py + h, yl = py - h;
  const volatile REAL pp = py * px;
  /* side k: s0 -> s1 */
  const REAL s0x[4] = {px, xh, px, xl}, s0y[4] = {yl, py, yh, py};
  const REAL s1x[4] = {xh, px, xl, px}, s1y[4] = {py, yh, py, yl};
  /* c2 = s0.x*s1.y - s1.x*s0.y: the product py*px is rounded, the other one fused */
  const REAL c2[4] = {FMA(-yl, xh, pp), FMA(yh, xh, -pp), FMA(-yh, xl, pp), FMA(yl, xl, -pp)};
  int hit = 0;
  for (int k = 0; k < 4; ++k) {
    const REAL a2 = s0y[k] - s1y[k], b2 = s1x[k] - s0x[k];
    const volatile REAL t0 = b1 * a2;
    const REAL d = FMA(a1, b2, -t0); /* :196 / :210 */
    REAL cx = (sizeof(REAL) == 4) ? (REAL)3.402823466e+38f : (REAL)1.7976931348623157e308, cy = 0; /* :198 */
    if (d != 0) {
      const REAL r = (REAL)1 / d;
      const volatile REAL t1 = c1 * b2;
      const volatile REAL t2 = a1 * c2[k];
      const volatile REAL nx = FMA(b1, c2[k], -t1), ny = FMA(c1, a2, -t2);
      cx = nx * r; cy = ny * r;
    }
    hit |= FN(within)(s0x[k], s0y[k], s1x[k], s1y[k], cx, cy) && FN(within)(p1x, p1y, p2x, p2y, cx, cy);
  }
  return hit;
}

/* edge function as rounded in the lines kernel: fma(v_ab.x, v_ap.y, -rn(v_ab.y*v_ap.x)); `hoisted` selects
 * the other contraction, used for edge 0 in canonical orientation (see drtk_b200/csrc/rasterize.cu)        */
static inline REAL FN(edge_fn_lines)(REAL ax, REAL ay, REAL bx, REAL by, REAL px, REAL py, int hoisted) {
  const REAL abx = bx - ax, aby = by - ay, apx = px - ax, apy = py - ay;
  if (hoisted) { const volatile REAL t = apy * abx; return FMA(-aby, apx, t); }
  const volatile REAL t = aby * apx;
  return FMA(abx, apy, -t);
}

void FN(oracle_rasterize_lines)(const REAL* v, const int32_t* vi, int64_t N, int64_t V, int64_t F, int64_t H,
                                int64_t W, int vi_batched, float* depth_img, int32_t* index_img) {
  const int64_t HW = H * W;
  for (int64_t n = 0; n < N; ++n) {
    uint64_t* best = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)HW);
    memset(best, 0xFF, sizeof(uint64_t) * (size_t)HW);
    const REAL* vn = v + n * V * 3;
    const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
    for (int64_t id = 0; id < F; ++id) {
      const uint32_t raw0 = (uint32_t)vin[id * 3 + 0];
      const int flag = (int)((raw0 & 0xF0000000u) >> 28);                                /* :293 */
      const int32_t i0 = (int32_t)(raw0 & 0x0FFFFFFFu), i1 = vin[id * 3 + 1], i2 = vin[id * 3 + 2];
      if (i0 == i1 && i1 == i2) continue;                                                /* :299 */
      const int vis0 = (flag & 1) != 0, vis1 = (flag & 2) != 0, vis2 = (flag & 4) != 0;  /* :301-303 */
      const REAL p0x = vn[i0 * 3 + 0], p0y = vn[i0 * 3 + 1], z0 = vn[i0 * 3 + 2];
      const REAL p1x = vn[i1 * 3 + 0], p1y = vn[i1 * 3 + 1], z1 = vn[i1 * 3 + 2];
      const REAL p2x = vn[i2 * 3 + 0], p2y = vn[i2 * 3 + 1], z2 = vn[i2 * 3 + 2];
      if (!(z0 > (REAL)1e-8f && z1 > (REAL)1e-8f && z2 > (REAL)1e-8f)) continue;        /* :321 */
      const REAL mnx = FMIN(FMIN(p0x, p1x), p2x), mny = FMIN(FMIN(p0y, p1y), p2y);
      const REAL mxx = FMAX(FMAX(p0x, p1x), p2x), mxy = FMAX(FMAX(p0y, p1y), p2y);
      if (!(mnx <= (REAL)(W - 1) && mny <= (REAL)(H - 1) && mxx > 0 && mxy > 0)) continue; /* :322-323 */
      const REAL v01x = p1x - p0x, v01y = p1y - p0y, v02x = p2x - p0x, v02y = p2y - p0y;
      const REAL v12x = p2x - p1x, v12y = p2y - p1y;
      const volatile REAL tden = v01y * v02x;
      const REAL den = FMA(v01x, v02y, -tden);                                           /* :330 */
      if (den == 0) continue;
      int64_t bx0 = (int64_t)FN(f2i_trunc)(mnx) - 2; if (bx0 < 1) bx0 = 1;               /* :333-337 */
      int64_t by0 = (int64_t)FN(f2i_trunc)(mny) - 2; if (by0 < 1) by0 = 1;
      int64_t bx1 = (int64_t)FN(f2i_trunc)(mxx) + 2; if (bx1 > W - 2) bx1 = W - 2;
      int64_t by1 = (int64_t)FN(f2i_trunc)(mxy) + 2; if (by1 > H - 2) by1 = H - 2;
      int tl[3];
      FN(top_left)(den, v01x, v01y, v02x, v02y, v12x, v12y, tl);
      const REAL s = FN(sgn)(den);
      const REAL rad = (REAL)1 / (den < 0 ? -den : den);
      const REAL d0 = (REAL)1 / FN(epsclamp)(z0), d1 = (REAL)1 / FN(epsclamp)(z1), d2 = (REAL)1 / FN(epsclamp)(z2);
      /* lines of the three edges (:171-181): c = fma(p1.x, p2.y, -rn(p1.y*p2.x)) */
      const volatile REAL tc01 = p0y * p1x, tc12 = p1y * p2x, tc02 = p0y * p2x;
      const REAL a01 = p0y - p1y, b01 = p1x - p0x, c01 = FMA(p0x, p1y, -tc01);
      const REAL a12 = p1y - p2y, b12 = p2x - p1x, c12 = FMA(p1x, p2y, -tc12);
      const REAL a02 = p0y - p2y, b02 = p2x - p0x, c02 = FMA(p0x, p2y, -tc02);
      for (int64_t y = by0; y <= by1; ++y) {
        for (int64_t x = bx0; x <= bx1; ++x) {
          const REAL px = (REAL)x, py = (REAL)y;
          int hit = 0;                                                                  /* :343-346 */
          hit |= FN(crosses_diamond)(a01, b01, c01, p0x, p0y, p1x, p1y, px, py) && vis0;
          hit |= FN(crosses_diamond)(a12, b12, c12, p1x, p1y, p2x, p2y, px, py) && vis1;
          hit |= FN(crosses_diamond)(a02, b02, c02, p0x, p0y, p2x, p2y, px, py) && vis2;
          REAL b0 = (i1 <= i2) ? FN(edge_fn_lines)(p1x, p1y, p2x, p2y, px, py, 1) : -FN(edge_fn_lines)(p2x, p2y, p1x, p1y, px, py, 0);
          REAL b1 = (i2 <= i0) ? FN(edge_fn_lines)(p2x, p2y, p0x, p0y, px, py, 0) : -FN(edge_fn_lines)(p0x, p0y, p2x, p2y, px, py, 0);
          REAL b2 = (i0 <= i1) ? FN(edge_fn_lines)(p0x, p0y, p1x, p1y, px, py, 0) : -FN(edge_fn_lines)(p1x, p1y, p0x, p0y, px, py, 0);
          b0 *= s; b1 *= s; b2 *= s;                                                     /* :354 */
          const int inside = (b0 >= 0) && (b1 >= 0) && (b2 >= 0);
          const int keep = inside && !((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2]));
          if (!(keep || hit)) continue;                                                  /* :375 */
          b0 *= rad; b1 *= rad; b2 *= rad;                                               /* :376-379 */
          b0 = b0 < 0 ? 0 : (b0 > 1 ? 1 : b0); b1 = b1 < 0 ? 0 : (b1 > 1 ? 1 : b1); b2 = b2 < 0 ? 0 : (b2 > 1 ? 1 : b2);
          const volatile REAL s01 = b0 + b1;
          const REAL rs = (REAL)1 / (b2 + s01);
          b0 *= rs; b1 *= rs; b2 *= rs;
          const volatile REAL t = b1 * d1;
          const REAL inv = FMA(b2, d2, FMA(b0, d0, t));                                  /* :382-383 */
          const float depth_f = (float)((REAL)1 / FN(epsclamp)(inv));
          uint32_t dbits; memcpy(&dbits, &depth_f, 4);
          const uint64_t packed = ((uint64_t)dbits << 32) | (hit ? (uint64_t)(uint32_t)id : 0xFFFFFFFFull); /* :386-388 */
          const int64_t o = y * W + x;
          if (packed < best[o]) best[o] = packed;
        }
      }
    }
    for (int64_t o = 0; o < HW; ++o) { /* unpack :409-413 */
      const uint32_t du = (uint32_t)(best[o] >> 32);
      float df; memcpy(&df, &du, 4);
      depth_img[n * HW + o] = (du == 0xFFFFFFFFu) ? 0.0f : df;
      index_img[n * HW + o] = (int32_t)(uint32_t)(best[o] & 0xFFFFFFFFu);
    }
    free(best);
  }
}

/* Per-pixel triangle fetch shared by render/interpolate: vi rows are NOT nibble-masked there
 * (src/render/render_kernel.cu:69-72).                                                     */
#define FETCH_TRI(vin, t, i0, i1, i2) \
  const int32_t i0 = (vin)[(int64_t)(t) * 3 + 0], i1 = (vin)[(int64_t)(t) * 3 + 1], i2 = (vin)[(int64_t)(t) * 3 + 2]

/*
 * render forward: perspective-correct barycentrics + depth.
 * Restates render_kernel (src/render/render_kernel.cu:19-117; CPU twin
 * src/render/render_kernel_cpu.cpp:18-117).  bary_img is PLANAR [N,3,H,W] (:346).
 */
void FN(oracle_render_fwd)(const REAL* v, const int32_t* vi, const int32_t* index_img, int64_t N,
                           int64_t V, int64_t F, int64_t H, int64_t W, int vi_batched,
                           REAL* depth_img, REAL* bary_img) {
  const int64_t HW = H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    for (int64_t h = 0; h < H; ++h) {
      const REAL* vn = v + n * V * 3;
      const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
      for (int64_t w = 0; w < W; ++w) {
        const int64_t o = h * W + w;
        const int32_t t = index_img[n * HW + o];
        REAL* bo = bary_img + n * 3 * HW + o;
        if (t != -1) {
          FETCH_TRI(vin, t, i0, i1, i2);
          const REAL p0x = vn[i0 * 3], p0y = vn[i0 * 3 + 1], z0 = vn[i0 * 3 + 2];
          const REAL p1x = vn[i1 * 3], p1y = vn[i1 * 3 + 1], z1 = vn[i1 * 3 + 2];
          const REAL p2x = vn[i2 * 3], p2y = vn[i2 * 3 + 1], z2 = vn[i2 * 3 + 2];
          const REAL v01x = p1x - p0x, v01y = p1y - p0y, v02x = p2x - p0x, v02y = p2y - p0y;
          const REAL den = FN(epsclamp)(v01x * v02y - v01y * v02x); /* :88 */
          const REAL qx = (REAL)w - p0x, qy = (REAL)h - p0y;        /* :90 */
          const REAL b1 = (qx * v02y - qy * v02x) / den;             /* :92-96 */
          const REAL b2 = (qy * v01x - qx * v01y) / den;
          const REAL b0 = (REAL)1 - b1 - b2;                          /* :97 */
          const REAL d0 = (REAL)1 / FN(epsclamp)(z0), d1 = (REAL)1 / FN(epsclamp)(z1),
                     d2 = (REAL)1 / FN(epsclamp)(z2);
          const REAL dinv = d0 * b0 + d1 * b1 + d2 * b2;             /* :102 */
          const REAL depth = (REAL)1 / FN(epsclamp)(dinv);            /* :103 */
          bo[0] = d0 * b0 * depth; bo[HW] = d1 * b1 * depth; bo[2 * HW] = d2 * b2 * depth; /* :105 */
          depth_img[n * HW + o] = depth;
        } else {
          bo[0] = 0; bo[HW] = 0; bo[2 * HW] = 0; depth_img[n * HW + o] = 0; /* :110-115 */
        }
      }
    }
  }
}

/*
 * render backward: analytic chain rule with clamp masks, sequential scatter into grad_v.
 * Restates render_backward_kernel (src/render/render_kernel.cu:119-281; CPU twin
 * src/render/render_kernel_cpu.cpp:119-297).  grad_v [N,V,3] is zeroed here (:397).
 * grad_depth / grad_bary may be NULL (treated as zeros: the reference materialises them,
 * src/render/render_module.cpp:34).
 */
void FN(oracle_render_bwd)(const REAL* v, const int32_t* vi, const int32_t* index_img,
                           const REAL* grad_depth, const REAL* grad_bary, int64_t N, int64_t V,
                           int64_t F, int64_t H, int64_t W, int vi_batched, REAL* grad_v) {
  const int64_t HW = H * W;
  memset(grad_v, 0, sizeof(REAL) * (size_t)(N * V * 3));
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const REAL* vn = v + n * V * 3;
    REAL* gvn = grad_v + n * V * 3;
    const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
    for (int64_t h = 0; h < H; ++h)
      for (int64_t w = 0; w < W; ++w) {
        const int64_t o = h * W + w;
        const int32_t t = index_img[n * HW + o];
        if (t == -1) continue;
        FETCH_TRI(vin, t, i0, i1, i2);
        const REAL p0x = vn[i0 * 3], p0y = vn[i0 * 3 + 1], z0 = vn[i0 * 3 + 2];
        const REAL p1x = vn[i1 * 3], p1y = vn[i1 * 3 + 1], z1 = vn[i1 * 3 + 2];
        const REAL p2x = vn[i2 * 3], p2y = vn[i2 * 3 + 1], z2 = vn[i2 * 3 + 2];
        const REAL v01x = p1x - p0x, v01y = p1y - p0y, v02x = p2x - p0x, v02y = p2y - p0y;
        const REAL den_raw = v01x * v02y - v01y * v02x;
        const REAL den = FN(epsclamp)(den_raw);
        const int den_clamped = den != den_raw; /* :198 */
        const REAL qx = (REAL)w - p0x, qy = (REAL)h - p0y;
        const REAL b1 = (qx * v02y - qy * v02x) / den;
        const REAL b2 = (qy * v01x - qx * v01y) / den;
        const REAL b0 = (REAL)1 - b1 - b2;
        const REAL z0e = FN(epsclamp)(z0), z1e = FN(epsclamp)(z1), z2e = FN(epsclamp)(z2);
        const int c0 = z0e != z0, c1 = z1e != z1, c2 = z2e != z2; /* :211-213 */
        const REAL d0 = (REAL)1 / z0e, d1 = (REAL)1 / z1e, d2 = (REAL)1 / z2e;
        const REAL dinv = d0 * b0 + d1 * b1 + d2 * b2;
        const REAL dinv_e = FN(epsclamp)(dinv);
        const int dinv_clamped = dinv_e != dinv; /* :219 */
        const REAL depth = (REAL)1 / dinv_e;
        const REAL g0 = grad_bary ? grad_bary[n * 3 * HW + o] : 0;
        const REAL g1 = grad_bary ? grad_bary[n * 3 * HW + HW + o] : 0;
        const REAL g2 = grad_bary ? grad_bary[n * 3 * HW + 2 * HW + o] : 0;
        const REAL gd = grad_depth ? grad_depth[n * HW + o] : 0;
        const REAL dL_depth = gd + (g0 * d0 * b0 + g1 * d1 * b1 + g2 * d2 * b2);          /* :226 */
        const REAL dL_dinv = dinv_clamped ? 0 : (-dL_depth / (dinv * dinv));               /* :228-229 */
        const REAL dLd0 = g0 * b0 * depth + dL_dinv * b0;                                  /* :230 */
        const REAL dLd1 = g1 * b1 * depth + dL_dinv * b1;
        const REAL dLd2 = g2 * b2 * depth + dL_dinv * b2;
        gvn[i0 * 3 + 2] += c0 ? 0 : -dLd0 / (z0e * z0e);                                    /* :231-250 */
        gvn[i1 * 3 + 2] += c1 ? 0 : -dLd1 / (z1e * z1e);
        gvn[i2 * 3 + 2] += c2 ? 0 : -dLd2 / (z2e * z2e);
        const REAL dLb0 = g0 * d0 * depth + dL_dinv * d0;                                  /* :252 */
        const REAL dLb1 = g1 * d1 * depth + dL_dinv * d1;
        const REAL dLb2 = g2 * d2 * depth + dL_dinv * d2;
        const REAL e1 = (-dLb0 + dLb1) / den, e2 = (-dLb0 + dLb2) / den;                   /* :253-254 */
        const REAL dL_den = den_clamped ? 0 : -(e1 * b1 + e2 * b2);                        /* :256 */
        const REAL dqx = e1 * v02y - e2 * v01y, dqy = -e1 * v02x + e2 * v01x;              /* :258-260 */
        const REAL dv02x = -e1 * qy - dL_den * v01y, dv02y = e1 * qx + dL_den * v01x;      /* :262-264 */
        const REAL dv01x = e2 * qy + dL_den * v02y, dv01y = -e2 * qx - dL_den * v02x;      /* :265-267 */
        gvn[i0 * 3 + 0] += -dv02x - dv01x - dqx;                                           /* :269-278 */
        gvn[i0 * 3 + 1] += -dv02y - dv01y - dqy;
        gvn[i1 * 3 + 0] += dv01x; gvn[i1 * 3 + 1] += dv01y;
        gvn[i2 * 3 + 0] += dv02x; gvn[i2 * 3 + 1] += dv02y;
      }
  }
}

/*
 * interpolate forward.  Restates interpolate_kernel (src/interpolate/interpolate_kernel.cu:38-111;
 * CPU twin src/interpolate/interpolate_kernel_cpu.cpp:32-120).  Empty pixels receive the
 * coordinate sweep (:104-109), computed in float like the reference (2.0f literals).
 */
void FN(oracle_interpolate_fwd)(const REAL* attr, const int32_t* vi, const int32_t* index_img,
                                const REAL* bary_img, int64_t N, int64_t V, int64_t F, int64_t C,
                                int64_t H, int64_t W, int vi_batched, REAL* out) {
  const int64_t HW = H * W;
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t n = 0; n < N; ++n)
    for (int64_t h = 0; h < H; ++h) {
      const REAL* an = attr + n * V * C;
      const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
      for (int64_t w = 0; w < W; ++w) {
        const int64_t o = h * W + w;
        const int32_t t = index_img[n * HW + o];
        REAL* op = out + n * C * HW + o;
        if (t != -1) {
          FETCH_TRI(vin, t, i0, i1, i2);
          const REAL b0 = bary_img[n * 3 * HW + o], b1 = bary_img[n * 3 * HW + HW + o],
                     b2 = bary_img[n * 3 * HW + 2 * HW + o];
          for (int64_t c = 0; c < C; ++c)
            op[c * HW] = an[i0 * C + c] * b0 + an[i1 * C + c] * b1 + an[i2 * C + c] * b2; /* :102 */
        } else {
          const REAL sx = (REAL)(((float)w * 2.0f + 1.0f) / (float)W - 1.0f);
          const REAL sy = (REAL)(((float)h * 2.0f + 1.0f) / (float)H - 1.0f);
          for (int64_t c = 0; c < C; ++c) op[c * HW] = (c % 2) ? sy : sx; /* :106-107 */
        }
      }
    }
}

/*
 * interpolate backward.  Restates interpolate_backward_kernel
 * (src/interpolate/interpolate_kernel.cu:113-299; CPU twin
 * src/interpolate/interpolate_kernel_cpu.cpp:122-228) with a plain sequential scatter in
 * place of the warp-segmented reduction + atomics (same sums, fixed order).
 * Either output may be NULL (the reference computes only what requires grad, :610-639).
 * attr_grad [N,V,C] is zeroed here (:661); bary_grad [N,3,H,W] is fully written (:282-297).
 */
void FN(oracle_interpolate_bwd)(const REAL* grad_out, const REAL* attr, const int32_t* vi,
                                const int32_t* index_img, const REAL* bary_img, int64_t N,
                                int64_t V, int64_t F, int64_t C, int64_t H, int64_t W,
                                int vi_batched, REAL* attr_grad, REAL* bary_grad) {
  const int64_t HW = H * W;
  if (attr_grad) memset(attr_grad, 0, sizeof(REAL) * (size_t)(N * V * C));
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const REAL* an = attr + n * V * C;
    const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
    for (int64_t o = 0; o < HW; ++o) {
      const int32_t t = index_img[n * HW + o];
      if (t == -1) {
        if (bary_grad) {
          bary_grad[n * 3 * HW + o] = 0; bary_grad[n * 3 * HW + HW + o] = 0;
          bary_grad[n * 3 * HW + 2 * HW + o] = 0;
        }
        continue;
      }
      FETCH_TRI(vin, t, i0, i1, i2);
      const REAL b0 = bary_img[n * 3 * HW + o], b1 = bary_img[n * 3 * HW + HW + o],
                 b2 = bary_img[n * 3 * HW + 2 * HW + o];
      REAL gb0 = 0, gb1 = 0, gb2 = 0;
      for (int64_t c = 0; c < C; ++c) {
        const REAL g = grad_out[n * C * HW + c * HW + o];
        gb0 += g * an[i0 * C + c]; gb1 += g * an[i1 * C + c]; gb2 += g * an[i2 * C + c]; /* :256-258 */
        if (attr_grad) {
          REAL* gn = attr_grad + n * V * C;
          gn[i0 * C + c] += g * b0; gn[i1 * C + c] += g * b1; gn[i2 * C + c] += g * b2;  /* :262-279 */
        }
      }
      if (bary_grad) {
        bary_grad[n * 3 * HW + o] = gb0; bary_grad[n * 3 * HW + HW + o] = gb1;
        bary_grad[n * 3 * HW + 2 * HW + o] = gb2;
      }
    }
  }
}

/* ---------------- edge_grad helpers (src/edge_grad/edge_grad_kernel.cu:18-203) ---------- */
typedef struct { REAL p0x, p0y, p1x, p1y, v01x, v01y, v02x, v02y, v12x, v12y, den; } FN(TriInfo);

static inline FN(TriInfo) FN(tri_info)(const REAL* vn, int32_t i0, int32_t i1, int32_t i2) { /* :72-87 */
  FN(TriInfo) t;
  t.p0x = vn[i0 * 3]; t.p0y = vn[i0 * 3 + 1]; t.p1x = vn[i1 * 3]; t.p1y = vn[i1 * 3 + 1];
  const REAL p2x = vn[i2 * 3], p2y = vn[i2 * 3 + 1];
  t.v01x = t.p1x - t.p0x; t.v01y = t.p1y - t.p0y;
  t.v02x = p2x - t.p0x; t.v02y = p2y - t.p0y;
  t.v12x = p2x - t.p1x; t.v12y = p2y - t.p1y;
  t.den = t.v01x * t.v02y - t.v01y * t.v02x;
  return t;
}

/* :30-70 -- same top-left rule as rasterize but with PLAIN (non-canonical) edge functions */
static inline int FN(pix_in_tri)(const FN(TriInfo)* t, int64_t x, int64_t y) {
  if (t->den == 0) return 0;
  const REAL px = (REAL)x, py = (REAL)y;
  const REAL q0x = px - t->p0x, q0y = py - t->p0y, q1x = px - t->p1x, q1y = py - t->p1y;
  const REAL s = FN(sgn)(t->den);
  const REAL b0 = (q1y * t->v12x - q1x * t->v12y) * s;
  const REAL b1 = (q0x * t->v02y - q0y * t->v02x) * s;
  const REAL b2 = (q0y * t->v01x - q0x * t->v01y) * s;
  if (!(b0 >= 0 && b1 >= 0 && b2 >= 0)) return 0;
  int tl[3];
  FN(top_left)(t->den, t->v01x, t->v01y, t->v02x, t->v02y, t->v12x, t->v12y, tl);
  return !((b0 == 0 && !tl[0]) || (b1 == 0 && !tl[1]) || (b2 == 0 && !tl[2]));
}

/* :89-100  normalize(cross(p0 - p2, p1 - p0)) */
static inline void FN(tri_normal)(const REAL* vn, int32_t i0, int32_t i1, int32_t i2, REAL nrm[3]) {
  const REAL ax = vn[i0 * 3] - vn[i2 * 3], ay = vn[i0 * 3 + 1] - vn[i2 * 3 + 1],
             az = vn[i0 * 3 + 2] - vn[i2 * 3 + 2];
  const REAL bx = vn[i1 * 3] - vn[i0 * 3], by = vn[i1 * 3 + 1] - vn[i0 * 3 + 1],
             bz = vn[i1 * 3 + 2] - vn[i0 * 3 + 2];
  const REAL cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
  const REAL r = (REAL)1 / SQRT(cx * cx + cy * cy + cz * cz);
  nrm[0] = cx * r; nrm[1] = cy * r; nrm[2] = cz * r;
}

/* :102-203  d(edge position)/d(fragment position) for the intersection case */
static inline void FN(dp_dr)(REAL nvx, REAL nvy, REAL nfx, REAL nfy, REAL max_mag, REAL out[2]) {
  const REAL rv = (REAL)1 / SQRT(nvx * nvx + nvy * nvy);
  const REAL rf = (REAL)1 / SQRT(nfx * nfx + nfy * nfy);
  nvx *= rv; nvy *= rv; nfx *= rf; nfy *= rf;
  const REAL bx = -nfy, by = nfx;
  const REAL d = bx * nvx + by * nvy;
  REAL k;
  if (max_mag > 0) {
    const REAL ad = d < 0 ? -d : d;
    const REAL abm = (bx < 0 ? -bx : bx) / max_mag;
    const REAL safe = (d >= 0 ? (REAL)1 : (REAL)-1) * FN(epsclamp)(ad > abm ? ad : abm); /* :197-198 */
    k = bx / safe;
  } else {
    k = bx / FN(epsclamp)(d); /* :201 */
  }
  out[0] = k * nvx; out[1] = k * nvy;
}

/*
 * edge_grad backward.  Restates edge_grad_backward_kernel
 * (src/edge_grad/edge_grad_kernel.cu:217-449; CPU twin
 * src/edge_grad/edge_grad_kernel_cpu.cpp:139-359).  Output grad_v_pix_img [N,3,H,W] is
 * zeroed then accumulated in pixel order (the reference uses atomics, :427-445).
 */
void FN(oracle_edge_grad_bwd)(const REAL* v_pix, const REAL* img, const int32_t* index_img,
                              const int32_t* vi, const REAL* grad_output, int64_t N, int64_t V,
                              int64_t F, int64_t C, int64_t H, int64_t W, int vi_batched,
                              REAL max_dp_dr, REAL* grad_v_pix_img) {
  const int64_t HW = H * W;
  memset(grad_v_pix_img, 0, sizeof(REAL) * (size_t)(N * 3 * HW));
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const REAL* vn = v_pix + n * V * 3;
    const int32_t* vin = vi + (vi_batched ? n * F * 3 : 0);
    const int32_t* idx = index_img + n * HW;
    const REAL* im = img + n * C * HW;
    const REAL* go = grad_output + n * C * HW;
    REAL* gout = grad_v_pix_img + n * 3 * HW;
    for (int64_t y = 0; y < H - 1; ++y)
      for (int64_t x = 0; x < W - 1; ++x) { /* :270 */
        const int32_t ci = idx[y * W + x], ri = idx[y * W + x + 1], di = idx[(y + 1) * W + x];
        const int cv = ci >= 0, rv = ri >= 0, dv = di >= 0; /* :290-292 */
        int32_t c0 = 0, c1 = 0, c2 = 0, r0 = 0, r1 = 0, r2 = 0, e0 = 0, e1 = 0, e2 = 0; /* :296-301 */
        if (cv) { c0 = vin[(int64_t)ci * 3]; c1 = vin[(int64_t)ci * 3 + 1]; c2 = vin[(int64_t)ci * 3 + 2]; }
        if (rv) { r0 = vin[(int64_t)ri * 3]; r1 = vin[(int64_t)ri * 3 + 1]; r2 = vin[(int64_t)ri * 3 + 2]; }
        if (dv) { e0 = vin[(int64_t)di * 3]; e1 = vin[(int64_t)di * 3 + 1]; e2 = vin[(int64_t)di * 3 + 2]; }
        const int lr_diff = ci != ri, ud_diff = ci != di;
        const int xb = cv && rv, yb = cv && dv;
        const FN(TriInfo) tc = FN(tri_info)(vn, c0, c1, c2), tr = FN(tri_info)(vn, r0, r1, r2),
                          td = FN(tri_info)(vn, e0, e1, e2);
        const int c_in_r = lr_diff && xb && FN(pix_in_tri)(&tr, x, y);      /* :320-325 */
        const int r_in_c = lr_diff && xb && FN(pix_in_tri)(&tc, x + 1, y);
        const int c_in_d = ud_diff && yb && FN(pix_in_tri)(&td, x, y);
        const int d_in_c = ud_diff && yb && FN(pix_in_tri)(&tc, x, y + 1);
        const int l_over_r = c_in_r && !r_in_c, r_over_l = r_in_c && !c_in_r; /* :328-331 */
        const int u_over_d = c_in_d && !d_in_c, d_over_u = d_in_c && !c_in_d;
        const int horiz_int = c_in_r && r_in_c, vert_int = c_in_d && d_in_c; /* :334-335 */
        const int horiz_adj = lr_diff && xb && !c_in_r && !r_in_c;           /* :338-341 */
        const int vert_adj = ud_diff && yb && !c_in_d && !d_in_c;
        REAL gdx = 0, gdy = 0;
        if (lr_diff)                                                          /* :353-366 */
          for (int64_t c = 0; c < C; ++c)
            gdx += (im[c * HW + y * W + x + 1] - im[c * HW + y * W + x]) *
                   ((REAL)0.5 * (go[c * HW + y * W + x + 1] + go[c * HW + y * W + x]));
        if (ud_diff)                                                          /* :367-380 */
          for (int64_t c = 0; c < C; ++c)
            gdy += (im[c * HW + (y + 1) * W + x] - im[c * HW + y * W + x]) *
                   ((REAL)0.5 * (go[c * HW + (y + 1) * W + x] + go[c * HW + y * W + x]));
        REAL gc[3] = {0, 0, 0}, gr[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
        if (!horiz_int) {                                                     /* :391-393 */
          gc[0] += (!cv || r_over_l || horiz_adj) ? 0 : gdx;
          gr[0] += (!rv || l_over_r || horiz_adj) ? 0 : gdx;
        } else {                                                              /* :394-406 */
          REAL nc[3], nr[3], o2[2];
          FN(tri_normal)(vn, c0, c1, c2, nc); FN(tri_normal)(vn, r0, r1, r2, nr);
          FN(dp_dr)(nc[0], nc[2], nr[0], nr[2], max_dp_dr, o2);
          gc[0] += gdx * o2[0]; gc[2] += gdx * o2[1];
          FN(dp_dr)(nr[0], nr[2], nc[0], nc[2], max_dp_dr, o2);
          gr[0] += gdx * o2[0]; gr[2] += gdx * o2[1];
        }
        if (!vert_int) {                                                      /* :408-410 */
          gc[1] += (!cv || d_over_u || vert_adj) ? 0 : gdy;
          gd[1] += (!dv || u_over_d || vert_adj) ? 0 : gdy;
        } else {                                                              /* :411-423 */
          REAL nc[3], nd[3], o2[2];
          FN(tri_normal)(vn, c0, c1, c2, nc); FN(tri_normal)(vn, e0, e1, e2, nd);
          FN(dp_dr)(nc[1], nc[2], nd[1], nd[2], max_dp_dr, o2);
          gc[1] += gdy * o2[0]; gc[2] += gdy * o2[1];
          FN(dp_dr)(nd[1], nd[2], nc[1], nc[2], max_dp_dr, o2);
          gd[1] += gdy * o2[0]; gd[2] += gdy * o2[1];
        }
        for (int k = 0; k < 3; ++k) {                                         /* :427-445 */
          gout[k * HW + y * W + x] += -gc[k];
          gout[k * HW + y * W + x + 1] += -gr[k];
          gout[k * HW + (y + 1) * W + x] += -gd[k];
        }
      }
  }
}

#undef FETCH_TRI
#undef FN
#undef CAT
#undef CAT_
