// k_final.cu -- a14 reconcile + ordering (one warp per frame) and a15
// pose (one thread per detection: homography initialisation, 50 orthogonal iterations, ambiguity resolution).
// Restates AprilRobotics apriltag.c (reconcile block of apriltag_detector_detect), common/g2d.c
// (g2d_polygon_overlaps_polygon) and apriltag_pose.c (estimate_tag_pose) -- SURVEY App. A.8-A.9 -- with the
// oracle's operation order (double precision throughout).
#include <math_constants.h>

#include "detector.h"

namespace b200at {

constexpr int MAXC = 1024;  // candidates per frame upper bound (Geo::cand_cap <= MAXC)

__device__ bool seg_intersect_dev(const double a0[2], const double a1[2], const double b0[2], const double b1[2]) {
  double ua[2] = {a1[0] - a0[0], a1[1] - a0[1]};
  double ub[2] = {b1[0] - b0[0], b1[1] - b0[1]};
  double la = sqrt(ua[0] * ua[0] + ua[1] * ua[1]), lb = sqrt(ub[0] * ub[0] + ub[1] * ub[1]);
  ua[0] /= la;
  ua[1] /= la;
  ub[0] /= lb;
  ub[1] /= lb;
  double m00 = ua[0], m01 = -ub[0], m10 = ua[1], m11 = -ub[1];
  double det = m00 * m11 - m01 * m10;
  if (fabs(det) < 0.00000001) return false;
  double i00 = m11 / det, i01 = -m01 / det;
  double b00 = b0[0] - a0[0], b10 = b0[1] - a0[1];
  double x00 = i00 * b00 + i01 * b10;
  double px = ua[0] * x00 + a0[0], py = ua[1] * x00 + a0[1];
  double ta = (px - a0[0]) * ua[0] + (py - a0[1]) * ua[1];
  double tb = (px - b0[0]) * ub[0] + (py - b0[1]) * ub[1];
  double a_hi = (a1[0] - a0[0]) * ua[0] + (a1[1] - a0[1]) * ua[1];
  double b_hi = (b1[0] - b0[0]) * ub[0] + (b1[1] - b0[1]) * ub[1];
  if (ta < fmin(0.0, a_hi) || ta > fmax(0.0, a_hi)) return false;
  if (tb < fmin(0.0, b_hi) || tb > fmax(0.0, b_hi)) return false;
  return true;
}
__device__ bool poly_contains_point_dev(const double poly[4][2], const double q[2]) {
  int pos = 0, neg = 0;
  for (int i = 0; i < 4; i++) {
    const double *a = poly[i], *b = poly[(i + 1) & 3];
    double cr = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0]);
    if (cr > 0) pos++;
    if (cr < 0) neg++;
  }
  return !(pos > 0 && neg > 0);
}
__device__ bool polygon_overlaps_dev(const double a[4][2], const double b[4][2]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (seg_intersect_dev(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
  if (poly_contains_point_dev(a, b[0])) return true;
  if (poly_contains_point_dev(b, a[0])) return true;
  return false;
}
__device__ __forceinline__ int prefer_smaller_dev(int pref, double q0, double q1) {
  if (pref) return pref;
  if (q0 < q1) return -1;
  if (q1 < q0) return 1;
  return 0;
}

// One WARP per frame.  The candidates' ordering keys are staged in shared memory; both orderings (processing order by
// (cluster key, family), output order by (id, family, c.y, c.x)) are computed by rank counting across the lanes; the reconcile
// loop keeps upstream's control flow (removal = swap with last, restart of the outer element when it loses) but finds the next
// partner with the same (family, id) of the current element by a ballot over 32 positions at a time -- partners are rare, so an
// outer element costs m / 32 steps instead of m, and the polygon test runs only on real partners.
constexpr int RW = 2;  // frames (warps) per CTA
__global__ void __launch_bounds__(32 * RW) k_reconcile(Geo g, const Cand *__restrict__ cands, const uint32_t *__restrict__ cand_count,
                                                       b200AprilTagsDetection_t *__restrict__ out, uint32_t *__restrict__ out_count,
                                                       uint32_t *__restrict__ counters, int nframes) {
  __shared__ uint16_t s_idx_a[RW][MAXC];
  __shared__ uint16_t s_ord_a[RW][MAXC];
  __shared__ unsigned long long s_tag_a[RW][MAXC];  // family << 32 | id of the candidate at sorted position
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int fr = blockIdx.x * RW + wid;
  if (fr >= nframes) return;  // (whole warps leave; no block-wide barrier below)
  uint16_t *idx = s_idx_a[wid], *ord = s_ord_a[wid];
  unsigned long long *tag = s_tag_a[wid];
  const Cand *cf = cands + (size_t)fr * g.cand_cap;
  int n = (int)min(cand_count[fr], g.cand_cap);
  if (n > MAXC) n = MAXC;
  // canonical processing order: (cluster key, family); ties cannot occur (one candidate per cluster and family)
  for (int i = lane; i < n; i += 32) {
    const unsigned long long kv = cf[i].key;
    const int fv = cf[i].family;
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const unsigned long long kj = cf[j].key;
      const int fj = cf[j].family;
      rank += (kj < kv || (kj == kv && (fj < fv || (fj == fv && j < i)))) ? 1 : 0;
    }
    idx[rank] = (uint16_t)i;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const Cand &c = cf[idx[i]];
    tag[i] = ((unsigned long long)(uint32_t)c.family << 32) | (uint32_t)c.id;
  }
  __syncwarp();
  // reconcile: same control flow as upstream, removal = swap with last
  int m = n;
  for (int i0 = 0; i0 < m; i0++) {
    int i1 = i0 + 1;
    bool restart = false;
    while (i1 < m) {
      // first position >= i1 whose (family, id) equals that of i0
      const unsigned long long t0 = tag[i0];
      int found = -1;
      for (int b0 = i1; b0 < m && found < 0; b0 += 32) {
        const int j = b0 + lane;
        const unsigned bal = __ballot_sync(0xffffffffu, j < m && tag[j] == t0);
        if (bal) found = b0 + __ffs(bal) - 1;
      }
      if (found < 0) break;
      i1 = found;
      int pref = 0;
      if (lane == 0) {
        const Cand &d0 = cf[idx[i0]], &d1 = cf[idx[i1]];
        if (polygon_overlaps_dev(d0.p, d1.p)) {
          pref = prefer_smaller_dev(pref, d0.hamming, d1.hamming);
          pref = prefer_smaller_dev(pref, -d0.decision_margin, -d1.decision_margin);
          for (int i = 0; i < 4; i++) {
            pref = prefer_smaller_dev(pref, d0.p[i][0], d1.p[i][0]);
            pref = prefer_smaller_dev(pref, d0.p[i][1], d1.p[i][1]);
          }
          pref = pref < 0 ? 1 : 2;  // 1: d1 goes, 2: d0 goes
        }
      }
      pref = __shfl_sync(0xffffffffu, pref, 0);
      if (pref == 1) {
        if (lane == 0) {
          idx[i1] = idx[m - 1];
          tag[i1] = tag[m - 1];
        }
        m--;
        __syncwarp();
        // (upstream: i1-- then i1++: the element swapped in is examined next)
      } else if (pref == 2) {
        if (lane == 0) {
          idx[i0] = idx[m - 1];
          tag[i0] = tag[m - 1];
        }
        m--;
        __syncwarp();
        restart = true;  // (upstream: i0--, break: the element swapped in is processed at the same position)
        break;
      } else {
        i1++;
      }
    }
    if (restart) i0--;
  }
  __syncwarp();
  // output order: (id, family, c.y, c.x), position as the last tie-break
  for (int i = lane; i < m; i += 32) {
    const Cand &cv = cf[idx[i]];
    int rank = 0;
    for (int j = 0; j < m; j++) {
      const Cand &cj = cf[idx[j]];
      bool lt;
      if (cj.id != cv.id)
        lt = cj.id < cv.id;
      else if (cj.family != cv.family)
        lt = cj.family < cv.family;
      else if (cj.c[1] != cv.c[1])
        lt = cj.c[1] < cv.c[1];
      else if (cj.c[0] != cv.c[0])
        lt = cj.c[0] < cv.c[0];
      else
        lt = j < i;
      rank += lt ? 1 : 0;
    }
    ord[rank] = idx[i];
  }
  __syncwarp();
  int no = m;
  if (no > (int)g.max_tags) {
    no = (int)g.max_tags;
    if (lane == 0) atomicOr(&counters[CNT_STATUS], (uint32_t)ST_OUT_TRUNC);
  }
  for (int i = lane; i < no; i += 32) {
    const Cand &c = cf[ord[i]];
    b200AprilTagsDetection_t d;
    d.family = c.family;
    d.id = c.id;
    d.hamming = c.hamming;
    d.decision_margin = c.decision_margin;
    for (int k = 0; k < 9; k++) d.H[k] = c.H[k];
    d.c[0] = c.c[0];
    d.c[1] = c.c[1];
    for (int k = 0; k < 4; k++) {
      d.p[k][0] = c.p[k][0];
      d.p[k][1] = c.p[k][1];
    }
    for (int k = 0; k < 9; k++) d.R[k] = 0;
    d.t[0] = d.t[1] = d.t[2] = 0;
    d.pose_err = 0;
    out[(size_t)fr * g.max_tags + i] = d;
  }
  if (lane == 0) {
    out_count[fr] = (uint32_t)no;
    atomicAdd(&counters[CNT_DETS], (uint32_t)no);
  }
}

// ------------------------------------------------------------------------------------------------
// pose
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double det33_dev(const double *m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}
__device__ void mm33_dev(const double *A, const double *B, double *C) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
  for (int i = 0; i < 9; i++) C[i] = t[i];
}
__device__ void mv33_dev(const double *A, const double *v, double *o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
  o[0] = t[0];
  o[1] = t[1];
  o[2] = t[2];
}
__device__ void tr33_dev(const double *A, double *T) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[j * 3 + i] = A[i * 3 + j];
  for (int i = 0; i < 9; i++) T[i] = t[i];
}
__device__ bool inv33_dev(const double *m, double *o) {
  double d = det33_dev(m);
  if (d == 0) return false;
  double id = 1.0 / d;
  double t[9];
  t[0] = (m[4] * m[8] - m[5] * m[7]) * id;
  t[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  t[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  t[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  t[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  t[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  t[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  for (int i = 0; i < 9; i++) o[i] = t[i];
  return true;
}

// polar factor U*V' via cyclic Jacobi on A'A (same sweep order / rank handling as the oracle's polar_UVt)
__device__ void polar_UVt_dev(const double *A, double *R) {
  double At[9], S[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  tr33_dev(A, At);
  mm33_dev(At, A, S);
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = fabs(S[1]) + fabs(S[2]) + fabs(S[5]);
    // converged to double precision: a further rotation would change eigenvalues by off^2 / gap and the polar factor by less
    // than one ulp.  (The oracle sweeps on to 1e-32 -- for the rank-2 matrices of orthogonal_iteration that is one more rotation
    // with c == 1, s ~ 1e-17 per call: a third of k_pose's dependent instruction chain for a change of <= 1e-17 in R.)
    if (off <= 2.5e-16 * (fabs(S[0]) + fabs(S[4]) + fabs(S[8]))) break;
    for (int pi = 0; pi < 3; pi++) {
      int p = pi == 2 ? 0 : pi, q = pi == 0 ? 1 : 2;
      double apq = S[p * 3 + q];
      if (apq == 0) continue;
      double app = S[p * 3 + p], aqq = S[q * 3 + q];
      double theta = (aqq - app) / (2 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
      double c = rsqrt(t * t + 1), s = t * c;  // (1 ulp from the oracle's 1 / sqrt: half the dependent chain)
      for (int k = 0; k < 3; k++) {
        double skp = S[k * 3 + p], skq = S[k * 3 + q];
        S[k * 3 + p] = c * skp - s * skq;
        S[k * 3 + q] = s * skp + c * skq;
      }
      for (int k = 0; k < 3; k++) {
        double spk = S[p * 3 + k], sqk = S[q * 3 + k];
        S[p * 3 + k] = c * spk - s * sqk;
        S[q * 3 + k] = s * spk + c * sqk;
      }
      for (int k = 0; k < 3; k++) {
        double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  int idx[3] = {0, 1, 2};
  double ev[3] = {S[0], S[4], S[8]};
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 3; j++)
      if (ev[idx[j]] > ev[idx[i]]) {
        int t = idx[i];
        idx[i] = idx[j];
        idx[j] = t;
      }
  double Vs[9], U[9];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) Vs[r * 3 + c] = V[r * 3 + idx[c]];
  double smax = sqrt(fmax(ev[idx[0]], 0.0));
  int rank = 0;
  for (int c = 0; c < 3; c++) {
    double s = sqrt(fmax(ev[idx[c]], 0.0));
    if (s > 1e-12 * smax && s > 0) {
      double v[3] = {Vs[0 * 3 + c], Vs[1 * 3 + c], Vs[2 * 3 + c]}, u[3];
      mv33_dev(A, v, u);
      const double rs = 1.0 / s;  // one division per column instead of three (1 ulp from the oracle's u / s)
      for (int r = 0; r < 3; r++) U[r * 3 + c] = u[r] * rs;
      rank = c + 1;
    } else {
      break;
    }
  }
  if (rank == 2) {
    double ax = U[0], ay = U[3], az = U[6], bx = U[1], by = U[4], bz = U[7];
    U[2] = ay * bz - az * by;
    U[5] = az * bx - ax * bz;
    U[8] = ax * by - ay * bx;
  } else if (rank < 2) {
    for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1 : 0;
    for (int i = 0; i < 9; i++) Vs[i] = (i % 4 == 0) ? 1 : 0;
  }
  double Vt[9];
  tr33_dev(Vs, Vt);
  mm33_dev(U, Vt, R);
}

// Polar factor of a matrix whose third column is exactly zero (the object points of a tag are planar: M3 of orthogonal_iteration
// always has this form), in closed form: the first two columns are Q = A (A'A)^(-1/2) with the 2 x 2 inverse square root
// (S + delta I)^-1 * tau, delta = sqrt(det S), tau = sqrt(tr S + 2 delta); the third is their cross product (what the generic path
// -- SVD, completion of the rank-2 factor, determinant fix -- ends with as well).  Three dependent long operations (sqrt, sqrt,
// divide) instead of the Jacobi path's eight: the polar factor was two thirds of k_pose's dependent instruction chain.  Same
// matrix up to rounding (~1e-15 for ordinary views); nearly rank-deficient input takes the generic path.
__device__ void polar_planar_dev(const double *A, double *R) {
  const double ax = A[0], ay = A[3], az = A[6], bx = A[1], by = A[4], bz = A[7];
  const double p = ax * ax + ay * ay + az * az, q = ax * bx + ay * by + az * bz, r = bx * bx + by * by + bz * bz;
  const double det = p * r - q * q, tr = p + r;
  if (!(A[2] == 0 && A[5] == 0 && A[8] == 0) || !(det > 1e-10 * tr * tr)) {
    polar_UVt_dev(A, R);
    return;
  }
  const double delta = sqrt(det), tau = sqrt(tr + 2 * delta), inv = 1.0 / (delta * tau);
  const double m00 = (r + delta) * inv, m01 = -q * inv, m11 = (p + delta) * inv;
  const double q0x = ax * m00 + bx * m01, q0y = ay * m00 + by * m01, q0z = az * m00 + bz * m01;
  const double q1x = ax * m01 + bx * m11, q1y = ay * m01 + by * m11, q1z = az * m01 + bz * m11;
  R[0] = q0x;
  R[3] = q0y;
  R[6] = q0z;
  R[1] = q1x;
  R[4] = q1y;
  R[7] = q1z;
  R[2] = q0y * q1z - q0z * q1y;
  R[5] = q0z * q1x - q0x * q1z;
  R[8] = q0x * q1y - q0y * q1x;
}

__device__ void homography_to_pose_dev(const double *H, double fx, double fy, double cx, double cy, double R[9], double T[3]) {
  double R20 = H[6];
  double R21 = H[7];
  double TZ = H[8];
  double R00 = (H[0] - cx * R20) / fx;
  double R01 = (H[1] - cx * R21) / fx;
  double TX = (H[2] - cx * TZ) / fx;
  double R10 = (H[3] - cy * R20) / fy;
  double R11 = (H[4] - cy * R21) / fy;
  double TY = (H[5] - cy * TZ) / fy;
  double length1 = (double)sqrtf((float)(R00 * R00 + R10 * R10 + R20 * R20));
  double length2 = (double)sqrtf((float)(R01 * R01 + R11 * R11 + R21 * R21));
  double s = 1.0 / (double)sqrtf((float)(length1 * length2));
  if (TZ > 0) s *= -1;
  R20 *= s;
  R21 *= s;
  TZ *= s;
  R00 *= s;
  R01 *= s;
  TX *= s;
  R10 *= s;
  R11 *= s;
  TY *= s;
  double R02 = R10 * R21 - R20 * R11;
  double R12 = R20 * R01 - R00 * R21;
  double R22 = R00 * R11 - R10 * R01;
  double M[9] = {R00, R01, R02, R10, R11, R12, R20, R21, R22};
  polar_UVt_dev(M, R);
  T[0] = TX;
  T[1] = TY;
  T[2] = TZ;
}

__device__ void calculate_F_dev(const double v[3], double F[9]) {
  double n = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) F[i * 3 + j] = v[i] * v[j] / n;
}

__device__ double orthogonal_iteration_dev(const double v[4][3], const double p[4][3], double t[3], double R[9], int n_steps) {
  const int n_points = 4;
  double p_mean[3] = {0, 0, 0};
  for (int i = 0; i < n_points; i++)
    for (int k = 0; k < 3; k++) p_mean[k] += p[i][k];
  for (int k = 0; k < 3; k++) p_mean[k] *= 1.0 / n_points;
  double p_res[4][3];
  for (int i = 0; i < n_points; i++)
    for (int k = 0; k < 3; k++) p_res[i][k] = p[i][k] - p_mean[k];
  double F[4][9], avg_F[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n_points; i++) {
    calculate_F_dev(v[i], F[i]);
    for (int k = 0; k < 9; k++) avg_F[k] += F[i][k];
  }
  for (int k = 0; k < 9; k++) avg_F[k] *= 1.0 / n_points;
  double M1[9], M1_inv[9];
  for (int k = 0; k < 9; k++) M1[k] = ((k % 4 == 0) ? 1.0 : 0.0) - avg_F[k];
  inv33_dev(M1, M1_inv);
  // Same arithmetic as the oracle's loop, minus what it recomputes: R * p[j] is evaluated once per R (the oracle forms it three
  // times per step), and the object-space error -- which the oracle evaluates in every step but only returns from the last --
  // is evaluated in the last step only.  Bit-identical results, ~35 % fewer dependent double-precision instructions.
  double Rp[4][3];
  for (int j = 0; j < n_points; j++) mv33_dev(R, p[j], Rp[j]);
  double prev_error = CUDART_INF;
  for (int it = 0; it < n_steps; it++) {
    double M2[3] = {0, 0, 0};
    for (int j = 0; j < n_points; j++) {
      double FmI[9], u[3];
      for (int k = 0; k < 9; k++) FmI[k] = F[j][k] - ((k % 4 == 0) ? 1.0 : 0.0);
      mv33_dev(FmI, Rp[j], u);
      for (int k = 0; k < 3; k++) M2[k] += u[k];
    }
    for (int k = 0; k < 3; k++) M2[k] *= 1.0 / n_points;
    mv33_dev(M1_inv, M2, t);
    double q[4][3], q_mean[3] = {0, 0, 0};
    for (int j = 0; j < n_points; j++) {
      double Rpt[3];
      for (int k = 0; k < 3; k++) Rpt[k] = Rp[j][k] + t[k];
      mv33_dev(F[j], Rpt, q[j]);
      for (int k = 0; k < 3; k++) q_mean[k] += q[j][k];
    }
    for (int k = 0; k < 3; k++) q_mean[k] *= 1.0 / n_points;
    double M3[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n_points; j++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) M3[a * 3 + b] += (q[j][a] - q_mean[a]) * p_res[j][b];
    polar_planar_dev(M3, R);
    if (det33_dev(R) < 0) {
      R[2] *= -1;
      R[5] *= -1;
      R[8] *= -1;
    }
    for (int j = 0; j < n_points; j++) mv33_dev(R, p[j], Rp[j]);
    if (it == n_steps - 1) {
      double error = 0;
      for (int j = 0; j < 4; j++) {
        double ImF[9], Rpt[3], e[3];
        for (int k = 0; k < 9; k++) ImF[k] = ((k % 4 == 0) ? 1.0 : 0.0) - F[j][k];
        for (int k = 0; k < 3; k++) Rpt[k] = Rp[j][k] + t[k];
        mv33_dev(ImF, Rpt, e);
        error += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
      }
      prev_error = error;
    }
  }
  return prev_error;
}

__device__ __forceinline__ double polyval_dev(const double *p, int degree, double x) {
  double ret = 0, xp = 1;
  for (int i = 0; i <= degree; i++) {
    ret += p[i] * xp;
    xp *= x;
  }
  return ret;
}

// solve_poly_approx without recursion: derivative chain bottom-up (degree 1 .. 4)
__device__ void solve_poly_approx_dev(const double *p4, double *roots_out, int *n_roots_out) {
  const double MAX_ROOT = 1000;
  double polys[5][5];  // polys[d] = d-th degree member of the derivative chain, polys[4] = p
  for (int i = 0; i <= 4; i++) polys[4][i] = p4[i];
  for (int d = 4; d > 1; d--)
    for (int i = 0; i < d; i++) polys[d - 1][i] = (i + 1) * polys[d][i + 1];
  double roots[4];
  int n_roots = 0;
  {
    const double *p = polys[1];
    if (fabs(p[0]) > MAX_ROOT * fabs(p[1])) {
      n_roots = 0;
    } else {
      roots[0] = -p[0] / p[1];
      n_roots = 1;
    }
  }
  for (int degree = 2; degree <= 4; degree++) {
    const double *p = polys[degree];
    const double *p_der = polys[degree - 1];
    double der_roots[4];
    const int n_der_roots = n_roots;
    for (int i = 0; i < n_der_roots; i++) der_roots[i] = roots[i];
    n_roots = 0;
    for (int i = 0; i <= n_der_roots; i++) {
      double mn = (i == 0) ? -MAX_ROOT : der_roots[i - 1];
      double mx = (i == n_der_roots) ? MAX_ROOT : der_roots[i];
      if (polyval_dev(p, degree, mn) * polyval_dev(p, degree, mx) < 0) {
        double lower, upper;
        if (polyval_dev(p, degree, mn) < polyval_dev(p, degree, mx)) {
          lower = mn;
          upper = mx;
        } else {
          lower = mx;
          upper = mn;
        }
        double root = 0.5 * (lower + upper);
        double dx_old = upper - lower;
        double dx = dx_old;
        double f = polyval_dev(p, degree, root);
        double df = polyval_dev(p_der, degree - 1, root);
        for (int j = 0; j < 100; j++) {
          if (((f + df * (upper - root)) * (f + df * (lower - root)) > 0) || (fabs(2 * f) > fabs(dx_old * df))) {
            dx_old = dx;
            dx = 0.5 * (upper - lower);
            root = lower + dx;
          } else {
            dx_old = dx;
            dx = -f / df;
            root += dx;
          }
          if (root == upper || root == lower) break;
          f = polyval_dev(p, degree, root);
          df = polyval_dev(p_der, degree - 1, root);
          if (f > 0)
            upper = root;
          else
            lower = root;
        }
        roots[n_roots++] = root;
      } else if (polyval_dev(p, degree, mx) == 0) {
        roots[n_roots++] = mx;
      }
    }
  }
  for (int i = 0; i < n_roots; i++) roots_out[i] = roots[i];
  *n_roots_out = n_roots;
}

__device__ bool fix_pose_ambiguities_dev(const double v[4][3], const double p[4][3], const double t[3], const double R[9],
                                         double R2[9]) {
  const int n_points = 4;
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double tn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  double Rt3[3] = {t[0] / tn, t[1] / tn, t[2] / tn};
  double ex[3] = {1, 0, 0};
  double d = ex[0] * Rt3[0] + ex[1] * Rt3[1] + ex[2] * Rt3[2];
  double tmp[3] = {ex[0] - d * Rt3[0], ex[1] - d * Rt3[1], ex[2] - d * Rt3[2]};
  double tmn = sqrt(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
  double Rt1[3] = {tmp[0] / tmn, tmp[1] / tmn, tmp[2] / tmn};
  double Rt2[3] = {Rt3[1] * Rt1[2] - Rt3[2] * Rt1[1], Rt3[2] * Rt1[0] - Rt3[0] * Rt1[2], Rt3[0] * Rt1[1] - Rt3[1] * Rt1[0]};
  double R_t[9] = {Rt1[0], Rt1[1], Rt1[2], Rt2[0], Rt2[1], Rt2[2], Rt3[0], Rt3[1], Rt3[2]};
  double R_1_prime[9];
  mm33_dev(R_t, R, R_1_prime);
  double r31 = R_1_prime[6];
  double r32 = R_1_prime[7];
  double hypotenuse = sqrt(r31 * r31 + r32 * r32);
  if (hypotenuse < 1e-100) {
    r31 = 1;
    r32 = 0;
    hypotenuse = 1;
  }
  double R_z[9] = {r31 / hypotenuse, -r32 / hypotenuse, 0, r32 / hypotenuse, r31 / hypotenuse, 0, 0, 0, 1};
  double R_trans[9];
  mm33_dev(R_1_prime, R_z, R_trans);
  double sin_gamma = -R_trans[1];
  double cos_gamma = R_trans[4];
  double R_gamma[9] = {cos_gamma, -sin_gamma, 0, sin_gamma, cos_gamma, 0, 0, 0, 1};
  double sin_beta = -R_trans[6];
  double cos_beta = R_trans[8];
  double t_initial = atan2(sin_beta, cos_beta);
  double v_trans[4][3], p_trans[4][3], F_trans[4][9], avg_F_trans[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double R_zT[9];
  tr33_dev(R_z, R_zT);
  for (int i = 0; i < n_points; i++) {
    mv33_dev(R_zT, p[i], p_trans[i]);
    mv33_dev(R_t, v[i], v_trans[i]);
    calculate_F_dev(v_trans[i], F_trans[i]);
    for (int k = 0; k < 9; k++) avg_F_trans[k] += F_trans[i][k];
  }
  for (int k = 0; k < 9; k++) avg_F_trans[k] *= 1.0 / n_points;
  double G[9], ImA[9];
  for (int k = 0; k < 9; k++) ImA[k] = I3[k] - avg_F_trans[k];
  inv33_dev(ImA, G);
  for (int k = 0; k < 9; k++) G[k] *= 1.0 / n_points;
  const double M1[9] = {0, 0, 2, 0, 0, 0, -2, 0, 0};
  const double M2[9] = {-1, 0, 0, 0, 1, 0, 0, 0, -1};
  double b0[3] = {0, 0, 0}, b1[3] = {0, 0, 0}, b2[3] = {0, 0, 0};
  double RgM1[9], RgM2[9];
  mm33_dev(R_gamma, M1, RgM1);
  mm33_dev(R_gamma, M2, RgM2);
  for (int i = 0; i < n_points; i++) {
    double FmI[9], a[3], o[3];
    for (int k = 0; k < 9; k++) FmI[k] = F_trans[i][k] - I3[k];
    mv33_dev(R_gamma, p_trans[i], a);
    mv33_dev(FmI, a, o);
    for (int k = 0; k < 3; k++) b0[k] += o[k];
    mv33_dev(RgM1, p_trans[i], a);
    mv33_dev(FmI, a, o);
    for (int k = 0; k < 3; k++) b1[k] += o[k];
    mv33_dev(RgM2, p_trans[i], a);
    mv33_dev(FmI, a, o);
    for (int k = 0; k < 3; k++) b2[k] += o[k];
  }
  double b0_[3], b1_[3], b2_[3];
  mv33_dev(G, b0, b0_);
  mv33_dev(G, b1, b1_);
  mv33_dev(G, b2, b2_);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (int i = 0; i < n_points; i++) {
    double ImF[9], a[3], c0[3], c1[3], c2[3];
    for (int k = 0; k < 9; k++) ImF[k] = I3[k] - F_trans[i][k];
    mv33_dev(R_gamma, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b0_[k];
    mv33_dev(ImF, a, c0);
    mv33_dev(RgM1, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b1_[k];
    mv33_dev(ImF, a, c1);
    mv33_dev(RgM2, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b2_[k];
    mv33_dev(ImF, a, c2);
#define DOT3(x, y) ((x)[0] * (y)[0] + (x)[1] * (y)[1] + (x)[2] * (y)[2])
    a0 += DOT3(c0, c0);
    a1 += 2 * DOT3(c0, c1);
    a2 += DOT3(c1, c1) + 2 * DOT3(c0, c2);
    a3 += 2 * DOT3(c1, c2);
    a4 += DOT3(c2, c2);
#undef DOT3
  }
  double poly[5] = {a1, 2 * a2 - 4 * a0, 3 * a3 - 3 * a1, 4 * a4 - 2 * a2, -a3};
  double roots[4];
  int n_roots;
  solve_poly_approx_dev(poly, roots, &n_roots);
  double minima[4];
  int n_minima = 0;
  for (int i = 0; i < n_roots; i++) {
    double t1 = roots[i];
    double t2 = t1 * t1;
    double t3 = t1 * t2;
    double t4 = t1 * t3;
    double t5 = t1 * t4;
    if (a2 - 2 * a0 + (3 * a3 - 6 * a1) * t1 + (6 * a4 - 8 * a2 + 10 * a0) * t2 + (-8 * a3 + 6 * a1) * t3 +
            (-6 * a4 + 3 * a2) * t4 + a3 * t5 >=
        0) {
      double tt = 2 * atan(roots[i]);
      if (fabs(tt - t_initial) > 0.1) minima[n_minima++] = roots[i];
    }
  }
  if (n_minima == 1) {
    double tt = minima[0];
    double R_beta[9];
    for (int k = 0; k < 9; k++) R_beta[k] = M2[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= tt;
    for (int k = 0; k < 9; k++) R_beta[k] += M1[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= tt;
    for (int k = 0; k < 9; k++) R_beta[k] += I3[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= 1 / (1 + tt * tt);
    double R_tT[9], A[9], B[9];
    tr33_dev(R_t, R_tT);
    mm33_dev(R_tT, R_gamma, A);
    mm33_dev(A, R_beta, B);
    mm33_dev(B, R_zT, R2);
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(64) k_pose(Geo g, FitParams fp, b200AprilTagsDetection_t *__restrict__ out,
                                             const uint32_t *__restrict__ out_count, int nframes) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int fr = t / (int)g.max_tags, slot = t % (int)g.max_tags;
  if (fr >= nframes || slot >= (int)out_count[fr]) return;
  b200AprilTagsDetection_t *d = out + (size_t)fr * g.max_tags + slot;
  const double fx = fp.fx, fy = fp.fy, cx = fp.cx, cy = fp.cy, tagsize = fp.tagsize;
  const double scale = tagsize / 2.0;
  const double p[4][3] = {{-scale, scale, 0}, {scale, scale, 0}, {scale, -scale, 0}, {-scale, -scale, 0}};
  double v[4][3];
  for (int i = 0; i < 4; i++) {
    v[i][0] = (d->p[i][0] - cx) / fx;
    v[i][1] = (d->p[i][1] - cy) / fy;
    v[i][2] = 1;
  }
  double H[9];
  for (int k = 0; k < 9; k++) H[k] = d->H[k];
  double R1[9], t1[3], e1, R2[9], t2[3] = {0, 0, 0}, e2;
  {
    double R[9], T[3];
    homography_to_pose_dev(H, -fx, fy, cx, cy, R, T);
    T[0] *= scale;
    T[1] *= scale;
    T[2] *= scale;
    for (int j = 0; j < 3; j++) {
      R1[0 * 3 + j] = R[0 * 3 + j];
      R1[1 * 3 + j] = -R[1 * 3 + j];
      R1[2 * 3 + j] = -R[2 * 3 + j];
    }
    t1[0] = T[0];
    t1[1] = -T[1];
    t1[2] = -T[2];
  }
  e1 = orthogonal_iteration_dev(v, p, t1, R1, 50);
  if (fix_pose_ambiguities_dev(v, p, t1, R1, R2)) {
    e2 = orthogonal_iteration_dev(v, p, t2, R2, 50);
  } else {
    e2 = CUDART_INF;
  }
  if (e1 <= e2) {
    for (int k = 0; k < 9; k++) d->R[k] = R1[k];
    for (int k = 0; k < 3; k++) d->t[k] = t1[k];
    d->pose_err = e1;
  } else {
    for (int k = 0; k < 9; k++) d->R[k] = R2[k];
    for (int k = 0; k < 3; k++) d->t[k] = t2[k];
    d->pose_err = e2;
  }
}

int launch_finalize(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  k_reconcile<<<(nframes + RW - 1) / RW, 32 * RW, 0, s>>>(g, ws.cands, ws.cand_count, ws.out, ws.out_count, ws.counters, nframes);
  const int total = nframes * (int)g.max_tags;
  k_pose<<<(total + 63) / 64, 64, 0, s>>>(g, ws.fp, ws.out, ws.out_count, nframes);
  return 2;
}

}  // namespace b200at
