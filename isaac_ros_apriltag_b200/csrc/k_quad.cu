// k_quad.cu -- fit_quads (a10): one CTA per boundary-point cluster.
// Restates AprilRobotics fit_quad / ptsort / compute_lfps / quad_segment_maxima / fit_line
// (apriltag_quad_thresh.c) with the SAME floating-point types and evaluation order as the
// CPU oracle, so accept/reject decisions and the float corners are bit-identical (built with -fmad=false):
//   * angle-proxy `slope` in float, sort key = (slope, y, x) as one u64 -> bitonic sort in shared memory
//   * weighted prefix moments in double, accumulated SEQUENTIALLY (a parallel scan would change the roundings):
//     the six moments run as six lanes of one warp, each walking the cluster once
//   * line-fit error per point, 7-tap smoothing, local maxima, top-(max_nmaxima) selection, the C(n,4)
//     corner search (pair table + parallel lexicographic arg-min), final four lines, corner intersections,
//     area and angle gates.
// Latency/atomic-bound stage (O(boundary points)); bytes are negligible next to the dense stages.
#include <math_constants.h>

#include "detector.h"

namespace b200at {

constexpr int QT = 128;            // threads per CTA
constexpr int SORT_SMEM = 4096;    // keys sorted in shared memory up to this many points (32 KB)
constexpr int MAXM = 16;           // max_nmaxima upper bound

__device__ __forceinline__ void fit_line_dev(const LineFitPt *__restrict__ lfps, int sz, int i0, int i1, double *lineparm,
                                             double *err, double *mse) {
  double Mx, My, Mxx, Myy, Mxy, W;
  int N;
  if (i0 < i1) {
    N = i1 - i0 + 1;
    LineFitPt a = lfps[i1];
    Mx = a.Mx;
    My = a.My;
    Mxx = a.Mxx;
    Mxy = a.Mxy;
    Myy = a.Myy;
    W = a.W;
    if (i0 > 0) {
      LineFitPt b = lfps[i0 - 1];
      Mx -= b.Mx;
      My -= b.My;
      Mxx -= b.Mxx;
      Mxy -= b.Mxy;
      Myy -= b.Myy;
      W -= b.W;
    }
  } else {
    LineFitPt e = lfps[sz - 1], b = lfps[i0 - 1], a = lfps[i1];
    Mx = e.Mx - b.Mx;
    My = e.My - b.My;
    Mxx = e.Mxx - b.Mxx;
    Mxy = e.Mxy - b.Mxy;
    Myy = e.Myy - b.Myy;
    W = e.W - b.W;
    Mx += a.Mx;
    My += a.My;
    Mxx += a.Mxx;
    Mxy += a.Mxy;
    Myy += a.Myy;
    W += a.W;
    N = sz - i0 + i1 + 1;
  }
  double Ex = Mx / W;
  double Ey = My / W;
  double Cxx = Mxx / W - Ex * Ex;
  double Cxy = Mxy / W - Ex * Ey;
  double Cyy = Myy / W - Ey * Ey;
  double disc = (double)sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
  double eig_small = 0.5 * (Cxx + Cyy - disc);
  if (lineparm) {
    lineparm[0] = Ex;
    lineparm[1] = Ey;
    double eig = 0.5 * (Cxx + Cyy + disc);
    double nx1 = Cxx - eig;
    double ny1 = Cxy;
    double M1 = nx1 * nx1 + ny1 * ny1;
    double nx2 = Cxy;
    double ny2 = Cyy - eig;
    double M2 = nx2 * nx2 + ny2 * ny2;
    double nx, ny, M;
    if (M1 > M2) {
      nx = nx1;
      ny = ny1;
      M = M1;
    } else {
      nx = nx2;
      ny = ny2;
      M = M2;
    }
    double length = (double)sqrtf((float)M);
    if (fabs(length) < 1e-12) {
      lineparm[2] = lineparm[3] = 0;
    } else {
      lineparm[2] = nx / length;
      lineparm[3] = ny / length;
    }
  }
  if (err) *err = N * eig_small;
  if (mse) *mse = eig_small;
}

__device__ __forceinline__ uint32_t float_orderable(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ascending bitonic network in the "mirror" formulation: every compare-exchange puts the smaller key at the
// lower index, so the virtual +inf padding above n never has to be stored (works in shared or global memory).
__device__ void sort_keys(unsigned long long *a, int n) {
  int N = 1;
  while (N < n) N <<= 1;
  for (int k = 2; k <= N; k <<= 1) {
    for (int i = threadIdx.x; i < N; i += QT) {
      int l = i ^ (k - 1);
      if (l > i && l < n) {
        unsigned long long x = a[i], y = a[l];
        if (x > y) {
          a[i] = y;
          a[l] = x;
        }
      }
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N; i += QT) {
        int l = i ^ j;
        if (l > i && l < n) {
          unsigned long long x = a[i], y = a[l];
          if (x > y) {
            a[i] = y;
            a[l] = x;
          }
        }
      }
      __syncthreads();
    }
  }
}

struct BBoxRed {
  int xmin, xmax, ymin, ymax, sgx, sgy;
  long long s1;
};

__global__ void __launch_bounds__(QT) k_quadfit(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters,
                                                const uint32_t *__restrict__ pts, unsigned long long *__restrict__ keys,
                                                LineFitPt *__restrict__ lfps_pool, double *__restrict__ errs_pool,
                                                const uint8_t *__restrict__ dec, QuadRec *__restrict__ quads,
                                                uint32_t *__restrict__ counters, int Wp) {
  extern __shared__ unsigned long long skeys[];
  __shared__ BBoxRed s_red[QT / 32];
  __shared__ int s_cluster;
  __shared__ int s_fm[MAXM];
  __shared__ int s_nm;
  __shared__ int s_scan[QT / 32];
  __shared__ int s_run;
  __shared__ double s_rv[QT / 32];
  __shared__ int s_ri[QT / 32];
  __shared__ double s_thresh;
  __shared__ double p_err[MAXM][MAXM], p_mse[MAXM][MAXM], p_nx[MAXM][MAXM], p_ny[MAXM][MAXM];
  __shared__ unsigned long long s_rr[QT / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t ncl = min(counters[CNT_CLUSTERS], g.clu_cap);

  for (;;) {
    __syncthreads();
    if (tid == 0) s_cluster = (int)atomicAdd(&counters[6], 1u);
    __syncthreads();
    const int c = s_cluster;
    if ((uint32_t)c >= ncl) break;
    const ClusterRec cr = clusters[c];
    const int sz = (int)cr.count;
    const uint32_t o = cr.offset;
    const uint8_t *im = dec + (size_t)cr.frame * g.Hd * Wp;

    // ---- Phase A: bounding box + integer sums for the border-polarity test ----
    BBoxRed r = {1 << 30, -1, 1 << 30, -1, 0, 0, 0};
    for (int i = tid; i < sz; i += QT) {
      uint32_t p = pts[o + i];
      int x = p & 0x3fff, y = (p >> 14) & 0x3fff;
      int cxg = (p >> 28) & 3, cyg = (p >> 30) & 3;
      int gx = cxg == 0 ? 0 : (cxg == 1 ? 255 : -255), gy = cyg == 0 ? 0 : (cyg == 1 ? 255 : -255);
      r.xmin = min(r.xmin, x);
      r.xmax = max(r.xmax, x);
      r.ymin = min(r.ymin, y);
      r.ymax = max(r.ymax, y);
      r.sgx += gx;
      r.sgy += gy;
      r.s1 += (long long)x * gx + (long long)y * gy;
    }
    for (int of = 16; of > 0; of >>= 1) {
      r.xmin = min(r.xmin, __shfl_xor_sync(0xffffffffu, r.xmin, of));
      r.xmax = max(r.xmax, __shfl_xor_sync(0xffffffffu, r.xmax, of));
      r.ymin = min(r.ymin, __shfl_xor_sync(0xffffffffu, r.ymin, of));
      r.ymax = max(r.ymax, __shfl_xor_sync(0xffffffffu, r.ymax, of));
      r.sgx += __shfl_xor_sync(0xffffffffu, r.sgx, of);
      r.sgy += __shfl_xor_sync(0xffffffffu, r.sgy, of);
      r.s1 += __shfl_xor_sync(0xffffffffu, r.s1, of);
    }
    if (lane == 0) s_red[wid] = r;
    __syncthreads();
    r = s_red[0];
    for (int w = 1; w < QT / 32; w++) {
      BBoxRed q = s_red[w];
      r.xmin = min(r.xmin, q.xmin);
      r.xmax = max(r.xmax, q.xmax);
      r.ymin = min(r.ymin, q.ymin);
      r.ymax = max(r.ymax, q.ymax);
      r.sgx += q.sgx;
      r.sgy += q.sgy;
      r.s1 += q.s1;
    }
    if ((r.xmax - r.xmin) * (r.ymax - r.ymin) < fp.tag_width) continue;
    const float cx = (float)((r.xmin + r.xmax) * 0.5 + 0.05118);
    const float cy = (float)((r.ymin + r.ymax) * 0.5 + -0.028581);
    // dot = sum (x-cx)*gx + (y-cy)*gy, evaluated exactly on the integer parts (order independent)
    const double dotd = (double)r.s1 - (double)cx * (double)r.sgx - (double)cy * (double)r.sgy;
    const bool reversed = dotd < 0;
    if (!fp.reversed_border && reversed) continue;
    if (!fp.normal_border && !reversed) continue;

    // ---- Phase C: sort keys (slope | y | x) ----
    const bool in_smem = sz <= SORT_SMEM;
    unsigned long long *ka = in_smem ? skeys : (keys + o);
    for (int i = tid; i < sz; i += QT) {
      uint32_t p = pts[o + i];
      int x = p & 0x3fff, y = (p >> 14) & 0x3fff;
      float dx = (float)x - cx;
      float dy = (float)y - cy;
      float quadrant;
      if (dy > 0)
        quadrant = (dx > 0) ? 65536.0f : 131072.0f;
      else
        quadrant = (dx > 0) ? 0.0f : -65536.0f;
      if (dy < 0) {
        dy = -dy;
        dx = -dx;
      }
      if (dx < 0) {
        float tmp = dx;
        dx = dy;
        dy = -tmp;
      }
      float slope = quadrant + dy / dx;
      ka[i] = ((unsigned long long)float_orderable(slope) << 32) | ((unsigned long long)y << 16) | (unsigned long long)x;
    }
    __syncthreads();
    sort_keys(ka, sz);
    if (in_smem) {
      for (int i = tid; i < sz; i += QT) keys[o + i] = skeys[i];
    }
    __syncthreads();

    // ---- Phase E: line-fit terms, then SEQUENTIAL prefix sums (six lanes, one per moment) ----
    LineFitPt *lfps = lfps_pool + o;
    for (int i = tid; i < sz; i += QT) {
      unsigned long long k = ka[i];
      int px = (int)(k & 0xffff), py = (int)((k >> 16) & 0xffff);
      double x = px * .5 + 0.5;
      double y = py * .5 + 0.5;
      int ix = (int)x, iy = (int)y;
      double W = 1;
      if (ix > 0 && ix + 1 < g.Wd && iy > 0 && iy + 1 < g.Hd) {
        int grad_x = (int)im[(size_t)iy * Wp + ix + 1] - (int)im[(size_t)iy * Wp + ix - 1];
        int grad_y = (int)im[(size_t)(iy + 1) * Wp + ix] - (int)im[(size_t)(iy - 1) * Wp + ix];
        W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
      }
      double fx = x, fy = y;
      LineFitPt t;
      t.Mx = W * fx;
      t.My = W * fy;
      t.Mxx = W * fx * fx;
      t.Mxy = W * fx * fy;
      t.Myy = W * fy * fy;
      t.W = W;
      lfps[i] = t;
    }
    __syncthreads();
    if (tid < 6) {
      double *base = reinterpret_cast<double *>(lfps) + tid;
      double acc = 0;
      int i = 0;
      for (; i + 4 <= sz; i += 4) {
        double t0 = base[(size_t)(i + 0) * 6], t1 = base[(size_t)(i + 1) * 6], t2 = base[(size_t)(i + 2) * 6],
               t3 = base[(size_t)(i + 3) * 6];
        acc += t0;
        base[(size_t)(i + 0) * 6] = acc;
        acc += t1;
        base[(size_t)(i + 1) * 6] = acc;
        acc += t2;
        base[(size_t)(i + 2) * 6] = acc;
        acc += t3;
        base[(size_t)(i + 3) * 6] = acc;
      }
      for (; i < sz; i++) {
        acc += base[(size_t)i * 6];
        base[(size_t)i * 6] = acc;
      }
    }
    __syncthreads();

    // ---- Phase F/G: per-point line-fit error over a +-ksz window, then 7-tap smoothing (circular) ----
    const int ksz = min(20, sz / 12);
    if (ksz < 2) continue;
    double *errA = errs_pool + (size_t)2 * o, *errB = errA + sz;
    for (int i = tid; i < sz; i += QT) {
      double e;
      fit_line_dev(lfps, sz, (i + sz - ksz) % sz, (i + ksz) % sz, nullptr, &e, nullptr);
      errA[i] = e;
    }
    __syncthreads();
    for (int i = tid; i < sz; i += QT) {
      double acc = 0;
#pragma unroll
      for (int k = 0; k < 7; k++) acc += errA[(i + k - 3 + sz) % sz] * fp.smooth[k];
      errB[i] = acc;
    }
    __syncthreads();

    // ---- Phase H: local maxima, compacted in index order ----
    double *merr = errA;                                                  // copy of maxima errs (top-k removal)
    uint32_t *midx = reinterpret_cast<uint32_t *>(errA + (sz + 1) / 2);    // maxima indices
    if (tid == 0) s_run = 0;
    __syncthreads();
    for (int i0 = 0; i0 < sz; i0 += QT) {
      const int i = i0 + tid;
      bool is_max = false;
      double e = 0;
      if (i < sz) {
        e = errB[i];
        is_max = e > errB[(i + 1) % sz] && e > errB[(i + sz - 1) % sz];
      }
      unsigned bal = __ballot_sync(0xffffffffu, is_max);
      if (lane == 0) s_scan[wid] = __popc(bal);
      __syncthreads();
      int woff = 0, tot = 0;
      for (int w = 0; w < QT / 32; w++) {
        if (w < wid) woff += s_scan[w];
        tot += s_scan[w];
      }
      const int run = s_run;
      if (is_max) {
        int pos = run + woff + __popc(bal & ((1u << lane) - 1));
        midx[pos] = (uint32_t)i;
        merr[pos] = e;
      }
      __syncthreads();
      if (tid == 0) s_run = run + tot;
      __syncthreads();
    }
    const int nmax_all = s_run;
    if (nmax_all < 4) continue;

    // ---- Phase I: keep the max_nmaxima best maxima (threshold = (max_nmaxima+1)-th largest error) ----
    if (nmax_all > fp.max_nmaxima) {
      for (int round = 0; round <= fp.max_nmaxima; round++) {
        double bv = -CUDART_INF;
        int bi = 0x7fffffff;
        for (int i = tid; i < nmax_all; i += QT) {
          double v = merr[i];
          if (v > bv || (v == bv && i < bi)) {
            bv = v;
            bi = i;
          }
        }
        for (int of = 16; of > 0; of >>= 1) {
          double ov = __shfl_xor_sync(0xffffffffu, bv, of);
          int oi = __shfl_xor_sync(0xffffffffu, bi, of);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        if (lane == 0) {
          s_rv[wid] = bv;
          s_ri[wid] = bi;
        }
        __syncthreads();
        if (tid == 0) {
          for (int w = 1; w < QT / 32; w++) {
            if (s_rv[w] > bv || (s_rv[w] == bv && s_ri[w] < bi)) {
              bv = s_rv[w];
              bi = s_ri[w];
            }
          }
          s_thresh = bv;
          if (bi != 0x7fffffff) merr[bi] = -CUDART_INF;
        }
        __syncthreads();
      }
      const double maxima_thresh = s_thresh;
      // ordered compaction of maxima with err > thresh (errB holds the untouched values)
      if (tid == 0) s_run = 0;
      __syncthreads();
      for (int i0 = 0; i0 < nmax_all; i0 += QT) {
        const int i = i0 + tid;
        bool keep = false;
        uint32_t idx = 0;
        if (i < nmax_all) {
          idx = midx[i];
          keep = !(errB[idx] <= maxima_thresh);
        }
        unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_scan[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < QT / 32; w++) {
          if (w < wid) woff += s_scan[w];
          tot += s_scan[w];
        }
        const int run = s_run;
        if (keep) {
          int pos = run + woff + __popc(bal & ((1u << lane) - 1));
          if (pos < MAXM) s_fm[pos] = (int)idx;
        }
        __syncthreads();
        if (tid == 0) s_run = run + tot;
        __syncthreads();
      }
      if (tid == 0) s_nm = min(s_run, MAXM);
    } else {
      if (tid < nmax_all) s_fm[tid] = (int)midx[tid];
      if (tid == 0) s_nm = nmax_all;
    }
    __syncthreads();
    const int nm = s_nm;
    if (nm < 4) continue;

    // ---- Phase J: line fits between every ordered pair of kept maxima ----
    for (int t = tid; t < nm * nm; t += QT) {
      int a = t / nm, b = t % nm;
      if (a == b) continue;
      double lp[4], e, m;
      fit_line_dev(lfps, sz, s_fm[a], s_fm[b], lp, &e, &m);
      p_err[a][b] = e;
      p_mse[a][b] = m;
      p_nx[a][b] = lp[2];
      p_ny[a][b] = lp[3];
    }
    __syncthreads();

    // ---- Phase K: best (m0<m1<m2<m3); ties resolved to the lexicographically first, like the serial loops ----
    double best = CUDART_INF;
    uint32_t brank = 0xffffffffu;
    {
      const double max_mse = (double)fp.max_line_fit_mse, max_dot = (double)fp.cos_critical_rad;
      const int total = nm * nm * nm * nm;
      for (int t = tid; t < total; t += QT) {
        int m3 = t % nm, q = t / nm;
        int m2 = q % nm;
        q /= nm;
        int m1 = q % nm, m0 = q / nm;
        if (!(m0 < m1 && m1 < m2 && m2 < m3)) continue;
        if (p_mse[m0][m1] > max_mse) continue;
        if (p_mse[m1][m2] > max_mse) continue;
        double dot = p_nx[m0][m1] * p_nx[m1][m2] + p_ny[m0][m1] * p_ny[m1][m2];
        if (fabs(dot) > max_dot) continue;
        if (p_mse[m2][m3] > max_mse) continue;
        if (p_mse[m3][m0] > max_mse) continue;
        double err = p_err[m0][m1] + p_err[m1][m2] + p_err[m2][m3] + p_err[m3][m0];
        uint32_t rank = (uint32_t)(((m0 * MAXM + m1) * MAXM + m2) * MAXM + m3);
        if (err < best || (err == best && rank < brank)) {
          best = err;
          brank = rank;
        }
      }
    }
    for (int of = 16; of > 0; of >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, best, of);
      uint32_t orank = __shfl_xor_sync(0xffffffffu, brank, of);
      if (ov < best || (ov == best && orank < brank)) {
        best = ov;
        brank = orank;
      }
    }
    if (lane == 0) {
      s_rv[wid] = best;
      s_rr[wid] = brank;
    }
    __syncthreads();

    // ---- Phase L: final lines, corners, area / angle gates (thread 0) ----
    if (tid == 0) {
      for (int w = 1; w < QT / 32; w++) {
        if (s_rv[w] < best || (s_rv[w] == best && (uint32_t)s_rr[w] < brank)) {
          best = s_rv[w];
          brank = (uint32_t)s_rr[w];
        }
      }
      bool ok = brank != 0xffffffffu;
      if (ok && !(best / sz < (double)fp.max_line_fit_mse)) ok = false;
      float qp[4][2];
      if (ok) {
        int mm[4] = {(int)(brank / (MAXM * MAXM * MAXM)), (int)(brank / (MAXM * MAXM)) % MAXM, (int)(brank / MAXM) % MAXM,
                     (int)(brank % MAXM)};
        int indices[4] = {s_fm[mm[0]], s_fm[mm[1]], s_fm[mm[2]], s_fm[mm[3]]};
        double lines[4][4];
        for (int i = 0; i < 4 && ok; i++) {
          double mse;
          fit_line_dev(lfps, sz, indices[i], indices[(i + 1) & 3], lines[i], nullptr, &mse);
          if (mse > (double)fp.max_line_fit_mse) ok = false;
        }
        for (int i = 0; i < 4 && ok; i++) {
          double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
          double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
          double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
          double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
          double det = A00 * A11 - A10 * A01;
          double W00 = A11 / det, W01 = -A01 / det;
          if (fabs(det) < 0.001) {
            ok = false;
            break;
          }
          double L0 = W00 * B0 + W01 * B1;
          qp[i][0] = (float)(lines[i][0] + L0 * A00);
          qp[i][1] = (float)(lines[i][1] + L0 * A10);
        }
      }
      if (ok) {
        double area = 0;
        double length[3], p;
        for (int i = 0; i < 3; i++) {
          int idxa = i, idxb = (i + 1) % 3;
          double ddx = (double)(qp[idxb][0] - qp[idxa][0]), ddy = (double)(qp[idxb][1] - qp[idxa][1]);
          length[i] = sqrt(ddx * ddx + ddy * ddy);
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        const int idxs[4] = {2, 3, 0, 2};
        for (int i = 0; i < 3; i++) {
          int idxa = idxs[i], idxb = idxs[i + 1];
          double ddx = (double)(qp[idxb][0] - qp[idxa][0]), ddy = (double)(qp[idxb][1] - qp[idxa][1]);
          length[i] = sqrt(ddx * ddx + ddy * ddy);
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        if (area < 0.95 * fp.tag_width * fp.tag_width) ok = false;
      }
      if (ok) {
        const double ccr = (double)fp.cos_critical_rad;
        for (int i = 0; i < 4; i++) {
          int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
          double dx1 = (double)(qp[i1][0] - qp[i0][0]);
          double dy1 = (double)(qp[i1][1] - qp[i0][1]);
          double dx2 = (double)(qp[i2][0] - qp[i1][0]);
          double dy2 = (double)(qp[i2][1] - qp[i1][1]);
          double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
          if ((cos_dtheta > ccr || cos_dtheta < -ccr) || dx1 * dy2 < dy1 * dx2) {
            ok = false;
            break;
          }
        }
      }
      if (ok) {
        uint32_t qi = atomicAdd(&counters[CNT_QUADS], 1u);
        if (qi < g.quad_cap) {
          QuadRec q;
          q.key = cr.key;
          for (int i = 0; i < 4; i++) {
            q.p[i][0] = qp[i][0];
            q.p[i][1] = qp[i][1];
          }
          q.frame = cr.frame;
          q.reversed_border = reversed ? 1u : 0u;
          quads[qi] = q;
        } else {
          atomicOr(&counters[CNT_STATUS], (uint32_t)ST_QUADS_FULL);
        }
      }
    }
  }
}

int launch_quadfit(const Workspace &ws, int nframes, cudaStream_t s) {
  (void)nframes;
  const Geo &g = ws.g;
  static bool attr_set = false;
  const size_t smem = (size_t)SORT_SMEM * sizeof(unsigned long long);
  if (!attr_set) {
    cudaFuncSetAttribute(k_quadfit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  k_quadfit<<<sms * 5, QT, smem, s>>>(g, ws.fp, ws.clusters, ws.pts, ws.keys, ws.lfps, ws.errs, ws.dec, ws.quads, ws.counters,
                                      at_Wp(g));
  return 1;
}

}  // namespace b200at
