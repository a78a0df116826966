/* oracle/pnp_ref.c — CPU checker for the NEXT row of the hot path: GSLAM::Estimator::findPnP
 * (GSLAM/core/Estimator.h:158-164, factory :175-191), minimal P3P solver + RANSAC + non-linear refinement.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under gslam_b200/ may import, link or execute this file (tests/test_abi.py checks it).
 *
 * Parity status: UNPINNED by the reference — the reference ships the interface only (`plugins/estimator` is not in the tree,
 * CMakeLists.txt:45), no test, no vector.  What pins this file: cv2.solveP3P on minimal problems (solution SETS agree to 1e-9),
 * ground-truth recovery on synthetic 2D-3D sets with outliers, cv2.solvePnPRansac on the same sets (pose agreement at the noise
 * level) — tests/test_oracle_pnp.py.
 *
 * Algorithm (our definition; chosen so that a GPU version evaluates hypotheses in parallel and still gives THIS result):
 *  P3P   Grunert's distance formulation: eliminate u = s2/s1 between the two ratio equations -> quartic in v = s3/s1 (built by
 *        polynomial arithmetic, no hand-expanded coefficients), quartic roots by bracketing between the critical points
 *        (bisection + two Newton steps; near-double roots are reported), pose from the three camera-frame points by frame alignment.
 *  RANSAC hypothesis h (h = 0,1,2,...) draws its three indices from a counter-based generator (splitmix64 of seed, h) — no
 *        sequential state, so hypotheses are independent work items; score = number of points with z > 0 and squared normalised
 *        reprojection error < threshold^2; best = (most inliers, then lowest hypothesis index, then lowest root index);
 *        hypotheses are consumed in batches of 64 and the run stops at the first batch boundary where
 *        h >= log(1-confidence)/log(1-w^3), w = best inlier ratio so far (or at max_hypotheses).
 *  refine Levenberg-Marquardt on the inliers (orc_ba_pnp: the same solver as Optimizer::optimizePnP), inlier set recomputed once.
 * Pose convention: the result is world2camera (T_cw) as the reference signature asks, layout {qx,qy,qz,qw,tx,ty,tz}.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gslam_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int orc_ba_pnp(int n, const double* xyz, const double* xy1, double* pose_wc, int dof, double* info6x6, const gb_ba_options* opt,
               gb_ba_result* res);
void orc_se3_inverse(const double* in, double* out);

/* ---- small helpers ---------------------------------------------------------------------------------------------------- */
static void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double norm3(const double* a) { return sqrt(dot3(a, a)); }

/* all real roots of x^3 + a x^2 + b x + c (trigonometric form for three real roots, Cardano otherwise); returns the count */
static int cubic_real_roots(double a, double b, double c, double* x) {
  const double q = (a * a - 3.0 * b) / 9.0, r = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  const double q3 = q * q * q;
  if (r * r < q3) {
    const double t = acos(fmax(-1.0, fmin(1.0, r / sqrt(q3)))), m = -2.0 * sqrt(q);
    x[0] = m * cos(t / 3.0) - a / 3.0;
    x[1] = m * cos((t + 2.0 * M_PI) / 3.0) - a / 3.0;
    x[2] = m * cos((t - 2.0 * M_PI) / 3.0) - a / 3.0;
    return 3;
  }
  double A = -cbrt(fabs(r) + sqrt(r * r - q3));
  if (r < 0) A = -A;
  const double B = (A == 0.0) ? 0.0 : q / A;
  x[0] = (A + B) - a / 3.0;
  return 1;
}

/* real roots of c[0] + c[1] x + ... + c[4] x^4 (c[4] != 0), ascending; returns the count.
 * Bracketing method (robust at the near-double roots P3P produces, branch-light and of fixed cost for a GPU version): the real roots
 * of a quartic are separated by the real roots of its derivative, so every interval between consecutive critical points (and the
 * Cauchy bound on both sides) holds at most one root — found by 64 bisection steps + two guarded Newton steps; a critical point at
 * which the polynomial (almost) touches zero without a sign change is reported as a double root. */
int orc_quartic_roots(const double c[5], double roots[4]) {
  if (c[4] == 0.0) return 0;
  const double a3 = c[3] / c[4], a2 = c[2] / c[4], a1 = c[1] / c[4], a0 = c[0] / c[4];
#define QF(x) (((((x) + a3) * (x) + a2) * (x) + a1) * (x) + a0)
#define QD(x) (((4.0 * (x) + 3.0 * a3) * (x) + 2.0 * a2) * (x) + a1)
  double crit[3];
  int nc = cubic_real_roots(0.75 * a3, 0.5 * a2, 0.25 * a1, crit);
  for (int i = 0; i < nc; ++i)  /* insertion sort */
    for (int j = i; j > 0 && crit[j] < crit[j - 1]; --j) { const double t = crit[j]; crit[j] = crit[j - 1]; crit[j - 1] = t; }
  const double B = 1.0 + fmax(fmax(fabs(a3), fabs(a2)), fmax(fabs(a1), fabs(a0)));  /* Cauchy bound */
  double brk[5];
  int nb = 0;
  brk[nb++] = -B;
  for (int i = 0; i < nc; ++i)
    if (crit[i] > -B && crit[i] < B) brk[nb++] = crit[i];
  brk[nb++] = B;
  int n = 0;
  for (int i = 0; i + 1 < nb; ++i) {
    double lo = brk[i], hi = brk[i + 1];
    double flo = QF(lo), fhi = QF(hi);
    if ((flo < 0.0) == (fhi < 0.0) && flo != 0.0 && fhi != 0.0) {
      /* no sign change: a (near-)tangent root at an interior critical point?  reported once, at the left end of the next interval */
      if (i > 0) {
        const double x = lo, scale = (((fabs(x) + fabs(a3)) * fabs(x) + fabs(a2)) * fabs(x) + fabs(a1)) * fabs(x) + fabs(a0);
        const double fprev = QF(brk[i - 1]);
        if (fabs(flo) <= 1e-12 * scale && (fprev < 0.0) == (flo < 0.0) && n < 4 && (n == 0 || roots[n - 1] != x)) roots[n++] = x;
      }
      continue;
    }
    if (flo == 0.0) { if (n < 4 && (n == 0 || roots[n - 1] != lo)) roots[n++] = lo; if (fhi != 0.0) continue; }
    if (fhi == 0.0) { if (i + 2 == nb && n < 4) roots[n++] = hi; continue; }
    for (int it = 0; it < 64; ++it) {
      const double mid = 0.5 * (lo + hi), fm = QF(mid);
      if (mid == lo || mid == hi) break;
      if ((fm < 0.0) == (flo < 0.0)) { lo = mid; flo = fm; } else { hi = mid; fhi = fm; }
    }
    double x = 0.5 * (lo + hi);
    for (int it = 0; it < 2; ++it) {
      const double d = QD(x);
      if (d != 0.0) {
        const double xn = x - QF(x) / d;
        if (xn >= lo && xn <= hi) x = xn;
      }
    }
    if (n < 4) roots[n++] = x;
  }
#undef QF
#undef QD
  return n;
}

/* polynomials in v as coefficient arrays (index = power) */
static void poly_mul(const double* a, int da, const double* b, int db, double* o) {
  for (int i = 0; i <= da + db; ++i) o[i] = 0.0;
  for (int i = 0; i <= da; ++i)
    for (int j = 0; j <= db; ++j) o[i + j] += a[i] * b[j];
}

/* P3P: X = three world points (row-major 3x3), f = three UNIT bearing vectors in the camera frame (row-major 3x3).
 * Writes up to 4 solutions as Rt[12] = R (row-major, world->camera) | t; returns the count. */
int orc_p3p(const double* X, const double* f, double* Rt_out) {
  const double *P1 = X, *P2 = X + 3, *P3 = X + 6, *f1 = f, *f2 = f + 3, *f3 = f + 6;
  double v12[3], v13[3], v23[3];
  for (int k = 0; k < 3; ++k) { v12[k] = P2[k] - P1[k]; v13[k] = P3[k] - P1[k]; v23[k] = P3[k] - P2[k]; }
  const double a2 = dot3(v23, v23), b2 = dot3(v13, v13), c2 = dot3(v12, v12);
  double nrm[3];
  cross3(v12, v13, nrm);
  if (a2 == 0.0 || b2 == 0.0 || c2 == 0.0 || dot3(nrm, nrm) < 1e-24 * b2 * c2) return 0; /* degenerate triangle */
  const double ca = dot3(f2, f3), cb = dot3(f1, f3), cg = dot3(f1, f2);
  /* q(v) = 1 + v^2 - 2 v cb;  E2 - E1:  u = N(v) / D(v),  N = b2 (1 - v^2) + (a2 - c2) q,  D = 2 b2 (cg - ca v)
   * E1 * D^2:  b2 N^2 - 2 b2 cg N D + (b2 - c2 q) D^2 = 0 */
  const double qv[3] = {1.0, -2.0 * cb, 1.0};
  double N[3], D[2];
  for (int k = 0; k < 3; ++k) N[k] = (a2 - c2) * qv[k];
  N[0] += b2; N[2] -= b2;
  D[0] = 2.0 * b2 * cg; D[1] = -2.0 * b2 * ca;
  double NN[5], ND[4], DD[3], K[3], KDD[5], poly[5];
  poly_mul(N, 2, N, 2, NN);
  poly_mul(N, 2, D, 1, ND);
  poly_mul(D, 1, D, 1, DD);
  for (int k = 0; k < 3; ++k) K[k] = -c2 * qv[k];
  K[0] += b2;
  poly_mul(K, 2, DD, 2, KDD);
  for (int k = 0; k < 5; ++k) poly[k] = b2 * NN[k] + KDD[k];
  for (int k = 0; k < 4; ++k) poly[k] -= 2.0 * b2 * cg * ND[k];
  double roots[4];
  int nr;
  if (fabs(poly[4]) < 1e-14 * (fabs(poly[0]) + fabs(poly[1]) + fabs(poly[2]) + fabs(poly[3]) + 1e-300)) return 0;
  nr = orc_quartic_roots(poly, roots);
  /* world frame of the triangle */
  double ex1[3], ex2[3], ex3[3];
  {
    const double l = norm3(v12);
    for (int k = 0; k < 3; ++k) ex1[k] = v12[k] / l;
    const double ln = norm3(nrm);
    for (int k = 0; k < 3; ++k) ex3[k] = nrm[k] / ln;
    cross3(ex3, ex1, ex2);
  }
  int ns = 0;
  for (int r = 0; r < nr; ++r) {
    const double v = roots[r];
    if (!(v > 0.0)) continue;
    const double q = 1.0 + v * v - 2.0 * v * cb;
    if (!(q > 0.0)) continue;
    const double den = 2.0 * b2 * (cg - ca * v);
    double u;
    if (fabs(den) > 1e-12 * b2) u = (b2 * (1.0 - v * v) + (a2 - c2) * q) / den;
    else { /* D(v) ~ 0: take u from E1 directly (quadratic in u), the root consistent with E2 */
      const double A = b2, B = -2.0 * b2 * cg, Cc = b2 - c2 * q, disc = B * B - 4.0 * A * Cc;
      if (disc < 0.0) continue;
      const double u0 = (-B + sqrt(disc)) / (2.0 * A), u1 = (-B - sqrt(disc)) / (2.0 * A);
      const double e0 = fabs(b2 * (u0 * u0 + v * v - 2.0 * u0 * v * ca) - a2 * q), e1 = fabs(b2 * (u1 * u1 + v * v - 2.0 * u1 * v * ca) - a2 * q);
      u = e0 <= e1 ? u0 : u1;
    }
    if (!(u > 0.0)) continue;
    /* the quartic is the SQUARED system: clustered roots lose ~6 digits there.  Three Newton steps on the original pair
     *   E1 = b2 (1 + u^2 - 2 u cg) - c2 q(v) = 0,   E2 = b2 (u^2 + v^2 - 2 u v ca) - a2 q(v) = 0
     * restore full accuracy (the pair is as well conditioned as the geometry allows). */
    double uu = u, vv = v;
    for (int it = 0; it < 3; ++it) {
      const double qq = 1.0 + vv * vv - 2.0 * vv * cb;
      const double E1 = b2 * (1.0 + uu * uu - 2.0 * uu * cg) - c2 * qq, E2 = b2 * (uu * uu + vv * vv - 2.0 * uu * vv * ca) - a2 * qq;
      const double J11 = b2 * (2.0 * uu - 2.0 * cg), J12 = -c2 * (2.0 * vv - 2.0 * cb);
      const double J21 = b2 * (2.0 * uu - 2.0 * vv * ca), J22 = b2 * (2.0 * vv - 2.0 * uu * ca) - a2 * (2.0 * vv - 2.0 * cb);
      const double det = J11 * J22 - J12 * J21;
      if (det == 0.0) break;
      const double du = (E1 * J22 - E2 * J12) / det, dv = (J11 * E2 - J21 * E1) / det;
      if (!(fabs(du) < 0.1 * (1.0 + fabs(uu))) || !(fabs(dv) < 0.1 * (1.0 + fabs(vv)))) break; /* not a refinement any more */
      uu -= du; vv -= dv;
    }
    if (!(uu > 0.0) || !(vv > 0.0)) continue;
    const double qr = 1.0 + vv * vv - 2.0 * vv * cb;
    if (!(qr > 0.0)) continue;
    const double s1 = sqrt(b2 / qr), s2 = uu * s1, s3 = vv * s1;
    /* consistency of the dropped equation (guards spurious roots of the squared system) */
    const double chk = s1 * s1 + s2 * s2 - 2.0 * s1 * s2 * cg;
    if (fabs(chk - c2) > 1e-7 * c2) continue;
    double Y1[3], Y2[3], Y3[3], w12[3], w13[3], wn[3], ey1[3], ey2[3], ey3[3];
    for (int k = 0; k < 3; ++k) { Y1[k] = s1 * f1[k]; Y2[k] = s2 * f2[k]; Y3[k] = s3 * f3[k]; w12[k] = Y2[k] - Y1[k]; w13[k] = Y3[k] - Y1[k]; }
    cross3(w12, w13, wn);
    const double l1 = norm3(w12), l3 = norm3(wn);
    if (l1 == 0.0 || l3 == 0.0) continue;
    for (int k = 0; k < 3; ++k) { ey1[k] = w12[k] / l1; ey3[k] = wn[k] / l3; }
    cross3(ey3, ey1, ey2);
    double* Rt = Rt_out + 12 * ns;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rt[3 * i + j] = ey1[i] * ex1[j] + ey2[i] * ex2[j] + ey3[i] * ex3[j]; /* R = Ey Ex^T */
    for (int i = 0; i < 3; ++i) Rt[9 + i] = Y1[i] - (Rt[3 * i] * P1[0] + Rt[3 * i + 1] * P1[1] + Rt[3 * i + 2] * P1[2]);
    ++ns;
  }
  return ns;
}

/* ---- RANSAC --------------------------------------------------------------------------------------------------------------- */
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
/* the three distinct indices of hypothesis h (counter-based: independent of every other hypothesis) */
void orc_pnp_sample(uint64_t seed, int h, int n, int idx[3]) {
  const uint64_t r0 = splitmix64(seed ^ (0x100000001B3ull * (uint64_t)(3 * h + 1)));
  const uint64_t r1 = splitmix64(seed ^ (0x100000001B3ull * (uint64_t)(3 * h + 2)));
  const uint64_t r2 = splitmix64(seed ^ (0x100000001B3ull * (uint64_t)(3 * h + 3)));
  int i0 = (int)(r0 % (uint64_t)n), i1 = (int)(r1 % (uint64_t)(n - 1)), i2 = (int)(r2 % (uint64_t)(n - 2));
  if (i1 >= i0) ++i1;
  const int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
  if (i2 >= lo) ++i2;
  if (i2 >= hi) ++i2;
  idx[0] = i0; idx[1] = i1; idx[2] = i2;
}

static int count_inliers(int n, const double* xyz, const double* xy, const double* Rt, double thr2, uint8_t* mask) {
  int c = 0;
  for (int k = 0; k < n; ++k) {
    const double* P = xyz + 3 * k;
    const double x = Rt[0] * P[0] + Rt[1] * P[1] + Rt[2] * P[2] + Rt[9];
    const double y = Rt[3] * P[0] + Rt[4] * P[1] + Rt[5] * P[2] + Rt[10];
    const double z = Rt[6] * P[0] + Rt[7] * P[1] + Rt[8] * P[2] + Rt[11];
    int in = 0;
    if (z > 0.0) {
      const double du = x / z - xy[2 * k], dv = y / z - xy[2 * k + 1];
      in = du * du + dv * dv < thr2;
    }
    if (mask) mask[k] = (uint8_t)in;
    c += in;
  }
  return c;
}

static void R_to_quat(const double* R, double* q) { /* {x,y,z,w}, w >= 0 */
  const double tr = R[0] + R[4] + R[8];
  double x, y, z, w;
  if (tr > 0.0) { const double s = sqrt(tr + 1.0) * 2.0; w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s; }
  else { const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s; }
  if (w < 0.0) { x = -x; y = -y; z = -z; w = -w; }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

typedef struct {
  int hypotheses;       /* hypotheses evaluated */
  int best_hypothesis;  /* index of the winning hypothesis, -1 if none */
  int best_root;
  int inliers_minimal;  /* inliers of the winning minimal solution */
  int inliers_refined;  /* inliers after refinement (size of the returned mask) */
} orc_pnp_stats;

/* xyz: n world points; xy: n normalised image points (x/z, y/z); pose_cw: out {qx,qy,qz,qw,tx,ty,tz} world->camera;
 * mask: out n bytes (may be NULL).  Returns GB_OK when a pose with >= 4 inliers was found, GB_ERR_NUMERIC otherwise. */
int orc_pnp_ransac(int n, const double* xyz, const double* xy, double threshold, double confidence, int max_hypotheses, uint64_t seed,
                   double* pose_cw, uint8_t* mask, orc_pnp_stats* st) {
  orc_pnp_stats s;
  memset(&s, 0, sizeof s);
  s.best_hypothesis = -1;
  if (st) *st = s;
  if (n < 4 || !xyz || !xy || !pose_cw || max_hypotheses < 1) return GB_ERR_INVALID;
  const double thr2 = threshold * threshold;
  double best_Rt[12];
  int best = 0;
  int h = 0;
  double needed = (double)max_hypotheses;
  while (h < max_hypotheses && (double)h < needed) {
    const int h_end = h + 64 < max_hypotheses ? h + 64 : max_hypotheses;
    for (; h < h_end; ++h) {
      int idx[3];
      orc_pnp_sample(seed, h, n, idx);
      double X[9], f[9], sol[48];
      for (int k = 0; k < 3; ++k) {
        memcpy(X + 3 * k, xyz + 3 * idx[k], 3 * sizeof(double));
        const double bx = xy[2 * idx[k]], by = xy[2 * idx[k] + 1], l = sqrt(bx * bx + by * by + 1.0);
        f[3 * k] = bx / l; f[3 * k + 1] = by / l; f[3 * k + 2] = 1.0 / l;
      }
      const int ns = orc_p3p(X, f, sol);
      for (int r = 0; r < ns; ++r) {
        const int c = count_inliers(n, xyz, xy, sol + 12 * r, thr2, NULL);
        if (c > best) { best = c; memcpy(best_Rt, sol + 12 * r, sizeof best_Rt); s.best_hypothesis = h; s.best_root = r; }
      }
    }
    const double w = (double)best / (double)n, w3 = w * w * w;
    if (w3 >= 1.0) needed = 0.0;
    else if (w3 > 0.0) needed = log(1.0 - confidence) / log(1.0 - w3);
  }
  s.hypotheses = h;
  s.inliers_minimal = best;
  if (st) *st = s;
  if (best < 4) return GB_ERR_NUMERIC;
  /* refinement on the inliers: the optimizePnP solver, started at the minimal solution */
  uint8_t* m = (uint8_t*)malloc((size_t)n);
  count_inliers(n, xyz, xy, best_Rt, thr2, m);
  double* ixyz = (double*)malloc(sizeof(double) * 3 * (size_t)best);
  double* ixy1 = (double*)malloc(sizeof(double) * 3 * (size_t)best);
  int c = 0;
  for (int k = 0; k < n; ++k)
    if (m[k]) {
      memcpy(ixyz + 3 * c, xyz + 3 * k, 3 * sizeof(double));
      ixy1[3 * c] = xy[2 * k]; ixy1[3 * c + 1] = xy[2 * k + 1]; ixy1[3 * c + 2] = 1.0;
      ++c;
    }
  double cw[7], wc[7];
  R_to_quat(best_Rt, cw);
  cw[4] = best_Rt[9]; cw[5] = best_Rt[10]; cw[6] = best_Rt[11];
  orc_se3_inverse(cw, wc);
  gb_ba_options o;
  memset(&o, 0, sizeof o);
  o.projection = 0; o.huber_delta = 0.0; o.max_iterations = 20; o.function_tolerance = 1e-12; o.lambda_init = 1e-4; o.pcg_max_iters = 50; o.pcg_tol = 1e-12;
  gb_ba_result res;
  int rc = orc_ba_pnp(c, ixyz, ixy1, wc, 63, NULL, &o, &res);
  free(ixyz); free(ixy1);
  if (rc == GB_OK) {
    double rcw[7];
    orc_se3_inverse(wc, rcw);
    /* keep the refined pose only if it does not lose inliers */
    const double x = rcw[0], y = rcw[1], z = rcw[2], w = rcw[3];
    double Rt[12] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y), rcw[4], rcw[5], rcw[6]};
    const int c2 = count_inliers(n, xyz, xy, Rt, thr2, m);
    if (c2 >= best) { memcpy(cw, rcw, sizeof cw); s.inliers_refined = c2; }
    else { s.inliers_refined = count_inliers(n, xyz, xy, best_Rt, thr2, m); }
  } else {
    s.inliers_refined = count_inliers(n, xyz, xy, best_Rt, thr2, m);
  }
  memcpy(pose_cw, cw, sizeof cw);
  if (mask) memcpy(mask, m, (size_t)n);
  free(m);
  if (st) *st = s;
  return GB_OK;
}
