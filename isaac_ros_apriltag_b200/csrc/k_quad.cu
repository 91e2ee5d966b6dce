// k_quad.cu -- fit_quads (a10).
// Restates AprilRobotics fit_quad / ptsort / compute_lfps / quad_segment_maxima / fit_line
// (apriltag_quad_thresh.c) with the SAME floating-point types and evaluation order as the
// CPU oracle, so accept/reject decisions and the float corners are bit-identical (built with -fmad=false).
//
// One CTA per boundary-point cluster, clusters binned by size so the CTA shape fits the work:
//   n <= 128 / 256     : 32-thread CTA (one warp, every barrier is warp-local), everything in shared memory (8 / 16 KB)
//   n <= 512 / 1024    : 64 / 128-thread CTA, everything in shared memory (33 / 66 KB)
//   n <= 2048 / larger : 256-thread CTA, sort keys in shared memory (n <= 4096), moments / errors in an L2-resident scratch
// Per cluster: slope keys (float) -> merge sort of u64 (slope|y|x) keys in shared memory -> line-fit terms -> SEQUENTIAL double prefix sums (a parallel scan would change the
// roundings; the six moments run as six lanes reading shared memory) -> per-point window error -> 7-tap smoothing ->
// local maxima -> top-(max_nmaxima) by rank counting -> pair table of line fits -> C(n,4) search over a precomputed
// combination table with lexicographic arg-min -> corners, area and angle gates.
// Latency-bound stage (O(boundary points)); HBM bytes are ~12 B per point (read packed point, write sorted key).
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "detector.h"

namespace b200at {

constexpr int MAXM = kMaxNMaxima;  // max_nmaxima upper bound

struct LF6 {
  double Mx, My, Mxx, Mxy, Myy, W;
};

// moments accessor: SoA in shared memory (stride = capacity + 1 doubles: the six scan lanes hit six different banks)
// or AoS in global memory
template <bool SMEM>
struct LfAcc {
  const double *m;        // smem SoA base
  int stride;
  const LineFitPt *g;     // global AoS base
  __device__ __forceinline__ LF6 get(int i) const {
    LF6 r;
    if (SMEM) {
      r.Mx = m[i];
      r.My = m[stride + i];
      r.Mxx = m[2 * stride + i];
      r.Mxy = m[3 * stride + i];
      r.Myy = m[4 * stride + i];
      r.W = m[5 * stride + i];
    } else {
      const LineFitPt p = g[i];
      r.Mx = p.Mx;
      r.My = p.My;
      r.Mxx = p.Mxx;
      r.Mxy = p.Mxy;
      r.Myy = p.Myy;
      r.W = p.W;
    }
    return r;
  }
};

template <bool SMEM>
__device__ __forceinline__ void fit_line_dev(const LfAcc<SMEM> &lf, int sz, int i0, int i1, double *lineparm, double *err,
                                             double *mse) {
  double Mx, My, Mxx, Myy, Mxy, W;
  int N;
  if (i0 < i1) {
    N = i1 - i0 + 1;
    LF6 a = lf.get(i1);
    Mx = a.Mx;
    My = a.My;
    Mxx = a.Mxx;
    Mxy = a.Mxy;
    Myy = a.Myy;
    W = a.W;
    if (i0 > 0) {
      LF6 b = lf.get(i0 - 1);
      Mx -= b.Mx;
      My -= b.My;
      Mxx -= b.Mxx;
      Mxy -= b.Mxy;
      Myy -= b.Myy;
      W -= b.W;
    }
  } else {
    LF6 e = lf.get(sz - 1), b = lf.get(i0 - 1), a = lf.get(i1);
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

template <int THREADS>
__device__ __forceinline__ void cta_sync() {
  if (THREADS == 32)
    __syncwarp();
  else
    __syncthreads();
}

// Sort of the u64 keys (unique within a cluster): every thread sorts ITEMS contiguous keys in registers (odd-even
// transposition network), then log2(n/ITEMS) merge passes between two buffers; in a pass each thread produces ITEMS
// consecutive outputs of its pair of runs, located with a merge-path binary search.  O(n log n) work instead of the
// O(n log^2 n) of a bitonic network, one barrier per pass.  Works on shared or global memory (generic pointers).
// The sorted sequence ends in `a`.
template <int THREADS, int ITEMS>
__device__ void sort_keys(unsigned long long *a, unsigned long long *tmp, int n) {
  constexpr unsigned long long INF = ~0ull;
  const int tid = threadIdx.x;
  for (int base = tid * ITEMS; base < n; base += THREADS * ITEMS) {
    unsigned long long r[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) r[k] = (base + k < n) ? a[base + k] : INF;
#pragma unroll
    for (int pass = 0; pass < ITEMS; pass++) {
#pragma unroll
      for (int k = (pass & 1); k + 1 < ITEMS; k += 2) {
        unsigned long long x = r[k], y = r[k + 1];
        r[k] = x < y ? x : y;
        r[k + 1] = x < y ? y : x;
      }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++)
      if (base + k < n) a[base + k] = r[k];
  }
  cta_sync<THREADS>();
  unsigned long long *src = a, *dst = tmp;
  for (int width = ITEMS; width < n; width <<= 1) {
    const int w2 = width << 1;
    for (int ob = tid * ITEMS; ob < n; ob += THREADS * ITEMS) {
      const int pair_lo = ob & ~(w2 - 1);
      const int a0 = pair_lo, a1 = min(pair_lo + width, n), b0 = a1, b1 = min(pair_lo + w2, n);
      const int na = a1 - a0, nb = b1 - b0;
      const int diag = ob - pair_lo;
      int lo = max(0, diag - nb), hi = min(diag, na);
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (src[a0 + mid] < src[b0 + diag - 1 - mid])
          lo = mid + 1;
        else
          hi = mid;
      }
      int ia = lo, ib = diag - lo;
      unsigned long long va = ia < na ? src[a0 + ia] : INF, vb = ib < nb ? src[b0 + ib] : INF;
#pragma unroll
      for (int k = 0; k < ITEMS; k++) {
        if (ob + k < b1) {
          const bool take_a = va <= vb;
          dst[ob + k] = take_a ? va : vb;
          if (take_a) {
            ia++;
            va = ia < na ? src[a0 + ia] : INF;
          } else {
            ib++;
            vb = ib < nb ? src[b0 + ib] : INF;
          }
        }
      }
    }
    cta_sync<THREADS>();
    unsigned long long *t = src;
    src = dst;
    dst = t;
  }
  if (src != a) {
    for (int i = tid; i < n; i += THREADS) a[i] = src[i];
    cta_sync<THREADS>();
  }
}

struct BBoxRed {
  int xmin, xmax, ymin, ymax, sgx, sgy;
  long long s1;
};

// combination table for the current nm: all (m0<m1<m2<m3) < nm in lexicographic order, one byte each
struct ComboTable {
  const uchar4 *c;
  int off[18];
};

template <int THREADS, int NCAP, bool ALL_SMEM, int ITEMS, int SCAN_CH>
__global__ void __launch_bounds__(THREADS) k_quadfit(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters,
                                                     const uint32_t *__restrict__ bin_idx, int bin, const uint32_t *__restrict__ pts,
                                                     unsigned long long *__restrict__ keys, LineFitPt *__restrict__ lfps_pool,
                                                     double *__restrict__ errs_pool, const uint8_t *__restrict__ dec,
                                                     QuadRec *__restrict__ quads, uint32_t *__restrict__ counters, ComboTable combos,
                                                     int Wp) {
  constexpr int NW = THREADS / 32;
  constexpr int MSTRIDE = (ALL_SMEM ? NCAP : SCAN_CH) + 1;
  extern __shared__ unsigned long long dsm[];
  unsigned long long *skeys = dsm;                                  // [NCAP]  keys (later: errA)
  double *s_errB = reinterpret_cast<double *>(dsm + NCAP);          // [NCAP]  sort scratch, later errB (ALL_SMEM)
  double *s_M = reinterpret_cast<double *>(dsm + 2 * NCAP);         // [6][MSTRIDE]
  __shared__ BBoxRed s_red[NW];
  __shared__ int s_cluster;
  __shared__ int s_fm[MAXM];
  __shared__ int s_nm;
  __shared__ int s_scan[NW];
  __shared__ int s_run;
  __shared__ double s_rv[NW];
  __shared__ int s_ri[NW];
  __shared__ unsigned int s_rr[NW];
  __shared__ double s_thresh;
  __shared__ double s_carry[6];
  __shared__ double p_err[MAXM][MAXM], p_mse[MAXM][MAXM], p_nx[MAXM][MAXM], p_ny[MAXM][MAXM];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t nbin = min(counters[CNT_BIN0 + bin], g.clu_cap);

  for (;;) {
    cta_sync<THREADS>();
    if (tid == 0) s_cluster = (int)atomicAdd(&counters[CNT_WORK0 + bin], 1u);
    cta_sync<THREADS>();
    const int cw = s_cluster;
    if ((uint32_t)cw >= nbin) break;
    const ClusterRec cr = clusters[bin_idx[(size_t)bin * g.clu_cap + cw]];
    const int sz = (int)cr.count;
    const uint32_t o = cr.offset;
    const uint8_t *im = dec + (size_t)cr.frame * g.Hd * Wp;
    const bool keys_in_smem = sz <= NCAP;  // bin C clusters larger than its capacity fall back to global memory

    // ---- Phase A: bounding box + integer sums for the border-polarity test ----
    BBoxRed r = {1 << 30, -1, 1 << 30, -1, 0, 0, 0};
    for (int i = tid; i < sz; i += THREADS) {
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
    if (NW > 1) {
      if (lane == 0) s_red[wid] = r;
      __syncthreads();
      r = s_red[0];
      for (int w = 1; w < NW; w++) {
        BBoxRed q = s_red[w];
        r.xmin = min(r.xmin, q.xmin);
        r.xmax = max(r.xmax, q.xmax);
        r.ymin = min(r.ymin, q.ymin);
        r.ymax = max(r.ymax, q.ymax);
        r.sgx += q.sgx;
        r.sgy += q.sgy;
        r.s1 += q.s1;
      }
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
    unsigned long long *ka = keys_in_smem ? skeys : (keys + o);
    for (int i = tid; i < sz; i += THREADS) {
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
    cta_sync<THREADS>();
    // scratch for the merge passes: the (not yet used) errB area, or the L2-resident error scratch for oversize clusters
    unsigned long long *ktmp = keys_in_smem ? reinterpret_cast<unsigned long long *>(s_errB)
                                            : reinterpret_cast<unsigned long long *>(errs_pool + (size_t)2 * o);
    sort_keys<THREADS, ITEMS>(ka, ktmp, sz);
    if (keys_in_smem) {
      for (int i = tid; i < sz; i += THREADS) keys[o + i] = skeys[i];
    }

    // ---- Phase E: line-fit terms, then SEQUENTIAL prefix sums (six lanes, one per moment) ----
    LineFitPt *lfps_g = lfps_pool + o;
    const int chunk = ALL_SMEM ? NCAP : SCAN_CH;
    if (!ALL_SMEM && tid < 6) s_carry[tid] = 0.0;
    for (int c0 = 0; c0 < sz; c0 += chunk) {
      const int cn = min(chunk, sz - c0);
      for (int ii = tid; ii < cn; ii += THREADS) {
        const unsigned long long k = ka[c0 + ii];
        const int px = (int)(k & 0xffff), py = (int)((k >> 16) & 0xffff);
        const double x = px * .5 + 0.5;
        const double y = py * .5 + 0.5;
        const int ix = (int)x, iy = (int)y;
        double W = 1;
        if (ix > 0 && ix + 1 < g.Wd && iy > 0 && iy + 1 < g.Hd) {
          int grad_x = (int)im[(size_t)iy * Wp + ix + 1] - (int)im[(size_t)iy * Wp + ix - 1];
          int grad_y = (int)im[(size_t)(iy + 1) * Wp + ix] - (int)im[(size_t)(iy - 1) * Wp + ix];
          W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
        }
        const double fx = x, fy = y;
        s_M[ii] = W * fx;
        s_M[MSTRIDE + ii] = W * fy;
        s_M[2 * MSTRIDE + ii] = W * fx * fx;
        s_M[3 * MSTRIDE + ii] = W * fx * fy;
        s_M[4 * MSTRIDE + ii] = W * fy * fy;
        s_M[5 * MSTRIDE + ii] = W;
      }
      cta_sync<THREADS>();
      if (tid < 6) {
        double *base = s_M + tid * MSTRIDE;
        double acc = ALL_SMEM ? 0.0 : s_carry[tid];
        int i = 0;
        for (; i + 8 <= cn; i += 8) {
          double t0 = base[i], t1 = base[i + 1], t2 = base[i + 2], t3 = base[i + 3], t4 = base[i + 4], t5 = base[i + 5],
                 t6 = base[i + 6], t7 = base[i + 7];
          acc += t0;
          base[i] = acc;
          acc += t1;
          base[i + 1] = acc;
          acc += t2;
          base[i + 2] = acc;
          acc += t3;
          base[i + 3] = acc;
          acc += t4;
          base[i + 4] = acc;
          acc += t5;
          base[i + 5] = acc;
          acc += t6;
          base[i + 6] = acc;
          acc += t7;
          base[i + 7] = acc;
        }
        for (; i < cn; i++) {
          acc += base[i];
          base[i] = acc;
        }
        if (!ALL_SMEM) s_carry[tid] = acc;
      }
      cta_sync<THREADS>();
      if (!ALL_SMEM) {
        for (int ii = tid; ii < cn; ii += THREADS) {
          LineFitPt t;
          t.Mx = s_M[ii];
          t.My = s_M[MSTRIDE + ii];
          t.Mxx = s_M[2 * MSTRIDE + ii];
          t.Mxy = s_M[3 * MSTRIDE + ii];
          t.Myy = s_M[4 * MSTRIDE + ii];
          t.W = s_M[5 * MSTRIDE + ii];
          lfps_g[c0 + ii] = t;
        }
        __syncthreads();
      }
    }
    LfAcc<ALL_SMEM> lf;
    lf.m = s_M;
    lf.stride = MSTRIDE;
    lf.g = lfps_g;

    // ---- Phase F/G: per-point line-fit error over a +-ksz window, then 7-tap smoothing (circular) ----
    const int ksz = min(20, sz / 12);
    if (ksz < 2) continue;
    // errA aliases the (now dead) key area when the keys are in shared memory
    double *errA = keys_in_smem ? reinterpret_cast<double *>(skeys) : (errs_pool + (size_t)2 * o);
    double *errB = ALL_SMEM ? s_errB : (errs_pool + (size_t)2 * o + sz);
    // two independent line fits per iteration: their long double-precision division chains overlap
    for (int i = tid; i < sz; i += 2 * THREADS) {
      const int i2 = i + THREADS;
      double e0, e1 = 0;
      int a0 = i - ksz, b0 = i + ksz;
      a0 = a0 < 0 ? a0 + sz : a0;
      b0 = b0 >= sz ? b0 - sz : b0;
      int a1 = i2 - ksz, b1 = i2 + ksz;
      a1 = a1 < 0 ? a1 + sz : a1;
      b1 = b1 >= sz ? b1 - sz : b1;
      fit_line_dev(lf, sz, a0, b0, nullptr, &e0, nullptr);
      if (i2 < sz) fit_line_dev(lf, sz, a1, b1, nullptr, &e1, nullptr);
      errA[i] = e0;
      if (i2 < sz) errA[i2] = e1;
    }
    cta_sync<THREADS>();
    for (int i = tid; i < sz; i += THREADS) {
      double acc = 0;
#pragma unroll
      for (int k = 0; k < 7; k++) {
        int j = i + k - 3;
        j = j < 0 ? j + sz : (j >= sz ? j - sz : j);
        acc += errA[j] * fp.smooth[k];
      }
      errB[i] = acc;
    }
    cta_sync<THREADS>();

    // ---- Phase H: local maxima, compacted in index order (errA area is dead again: reuse it) ----
    double *merr = errA;                                                  // values of the maxima
    uint32_t *midx = reinterpret_cast<uint32_t *>(errA + (sz + 1) / 2);    // their indices
    if (tid == 0) s_run = 0;
    cta_sync<THREADS>();
    for (int i0 = 0; i0 < sz; i0 += THREADS) {
      const int i = i0 + tid;
      bool is_max = false;
      double e = 0;
      if (i < sz) {
        e = errB[i];
        is_max = e > errB[i + 1 == sz ? 0 : i + 1] && e > errB[i == 0 ? sz - 1 : i - 1];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, is_max);
      int woff = 0, tot = __popc(bal);
      if (NW > 1) {
        if (lane == 0) s_scan[wid] = tot;
        __syncthreads();
        tot = 0;
        for (int w = 0; w < NW; w++) {
          if (w < wid) woff += s_scan[w];
          tot += s_scan[w];
        }
      }
      const int run = s_run;
      if (is_max) {
        int pos = run + woff + __popc(bal & ((1u << lane) - 1));
        midx[pos] = (uint32_t)i;
        merr[pos] = e;
      }
      cta_sync<THREADS>();
      if (tid == 0) s_run = run + tot;
      cta_sync<THREADS>();
    }
    const int nmax_all = s_run;
    if (nmax_all < 4) continue;

    // ---- Phase I: keep the max_nmaxima best maxima: threshold = value of descending rank max_nmaxima ----
    if (nmax_all > fp.max_nmaxima) {
      if (nmax_all <= 1024) {
        // rank by counting (unique ranks: ties broken by position), O(n^2 / THREADS)
        for (int i = tid; i < nmax_all; i += THREADS) {
          const double v = merr[i];
          int rank = 0;
          for (int j = 0; j < nmax_all; j++) {
            const double u = merr[j];
            rank += (u > v || (u == v && j < i)) ? 1 : 0;
          }
          if (rank == fp.max_nmaxima) s_thresh = v;
        }
        cta_sync<THREADS>();
      } else {
        for (int round = 0; round <= fp.max_nmaxima; round++) {
          double bv = -CUDART_INF;
          int bi = 0x7fffffff;
          for (int i = tid; i < nmax_all; i += THREADS) {
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
          cta_sync<THREADS>();
          if (tid == 0) {
            for (int w = 1; w < NW; w++) {
              if (s_rv[w] > bv || (s_rv[w] == bv && s_ri[w] < bi)) {
                bv = s_rv[w];
                bi = s_ri[w];
              }
            }
            s_thresh = bv;
            if (bi != 0x7fffffff) merr[bi] = -CUDART_INF;
          }
          cta_sync<THREADS>();
        }
      }
      const double maxima_thresh = s_thresh;
      // ordered compaction of maxima with err > thresh (errB holds the untouched values)
      if (tid == 0) s_run = 0;
      cta_sync<THREADS>();
      for (int i0 = 0; i0 < nmax_all; i0 += THREADS) {
        const int i = i0 + tid;
        bool keep = false;
        uint32_t idx = 0;
        if (i < nmax_all) {
          idx = midx[i];
          keep = !(errB[idx] <= maxima_thresh);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        int woff = 0, tot = __popc(bal);
        if (NW > 1) {
          if (lane == 0) s_scan[wid] = tot;
          __syncthreads();
          tot = 0;
          for (int w = 0; w < NW; w++) {
            if (w < wid) woff += s_scan[w];
            tot += s_scan[w];
          }
        }
        const int run = s_run;
        if (keep) {
          int pos = run + woff + __popc(bal & ((1u << lane) - 1));
          if (pos < MAXM) s_fm[pos] = (int)idx;
        }
        cta_sync<THREADS>();
        if (tid == 0) s_run = run + tot;
        cta_sync<THREADS>();
      }
      if (tid == 0) s_nm = min(s_run, MAXM);
    } else {
      if (tid < nmax_all) s_fm[tid] = (int)midx[tid];
      if (tid == 0) s_nm = nmax_all;
    }
    cta_sync<THREADS>();
    const int nm = s_nm;
    if (nm < 4) continue;

    // ---- Phase J: line fits between every ordered pair of kept maxima ----
    for (int t = tid; t < nm * nm; t += THREADS) {
      int a = t / nm, b = t - a * nm;
      if (a == b) continue;
      double lp[4], e, m;
      fit_line_dev(lf, sz, s_fm[a], s_fm[b], lp, &e, &m);
      p_err[a][b] = e;
      p_mse[a][b] = m;
      p_nx[a][b] = lp[2];
      p_ny[a][b] = lp[3];
    }
    cta_sync<THREADS>();

    // ---- Phase K: best (m0<m1<m2<m3); ties resolved to the lexicographically first, like the serial loops ----
    double best = CUDART_INF;
    uint32_t brank = 0xffffffffu;
    {
      const double max_mse = (double)fp.max_line_fit_mse, max_dot = (double)fp.cos_critical_rad;
      const uchar4 *ctab = combos.c + combos.off[nm];
      const int ncomb = combos.off[nm + 1] - combos.off[nm];
      for (int t = tid; t < ncomb; t += THREADS) {
        const uchar4 c = ctab[t];
        const int m0 = c.x, m1 = c.y, m2 = c.z, m3 = c.w;
        if (p_mse[m0][m1] > max_mse) continue;
        if (p_mse[m1][m2] > max_mse) continue;
        double dot = p_nx[m0][m1] * p_nx[m1][m2] + p_ny[m0][m1] * p_ny[m1][m2];
        if (fabs(dot) > max_dot) continue;
        if (p_mse[m2][m3] > max_mse) continue;
        if (p_mse[m3][m0] > max_mse) continue;
        double err = p_err[m0][m1] + p_err[m1][m2] + p_err[m2][m3] + p_err[m3][m0];
        // the table is in lexicographic order, so t itself is the lexicographic rank
        if (err < best || (err == best && (uint32_t)t < brank)) {
          best = err;
          brank = (uint32_t)t;
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
    if (NW > 1) {
      if (lane == 0) {
        s_rv[wid] = best;
        s_rr[wid] = brank;
      }
      __syncthreads();
    }

    // ---- Phase L: final lines, corners, area / angle gates: lanes 0..3 of warp 0, one line / corner / angle each ----
    if (wid == 0) {
      for (int w = 1; w < NW; w++) {
        if (s_rv[w] < best || (s_rv[w] == best && s_rr[w] < brank)) {
          best = s_rv[w];
          brank = s_rr[w];
        }
      }
      bool ok = brank != 0xffffffffu;
      if (ok && !(best / sz < (double)fp.max_line_fit_mse)) ok = false;
      if (ok) {  // warp-uniform
        const uchar4 c = combos.c[combos.off[nm] + brank];
        const int indices[4] = {s_fm[c.x], s_fm[c.y], s_fm[c.z], s_fm[c.w]};
        const int li = lane & 3;
        double ln[4], mse;
        fit_line_dev(lf, sz, indices[li], indices[(li + 1) & 3], ln, nullptr, &mse);
        ok = __all_sync(0xffffffffu, !(mse > (double)fp.max_line_fit_mse));
        // line of the next edge
        double nn[4];
#pragma unroll
        for (int k = 0; k < 4; k++) nn[k] = __shfl_sync(0xffffffffu, ln[k], (li + 1) & 3);
        const double A00 = ln[3], A01 = -nn[3];
        const double A10 = -ln[2], A11 = nn[2];
        const double B0 = -ln[0] + nn[0];
        const double B1 = -ln[1] + nn[1];
        const double det = A00 * A11 - A10 * A01;
        const double W00 = A11 / det, W01 = -A01 / det;
        const bool det_ok = !(fabs(det) < 0.001);
        const double L0 = W00 * B0 + W01 * B1;
        const float qx = (float)(ln[0] + L0 * A00);
        const float qy = (float)(ln[1] + L0 * A10);
        ok = ok && __all_sync(0xffffffffu, det_ok);
        float qp[4][2];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          qp[k][0] = __shfl_sync(0xffffffffu, qx, k);
          qp[k][1] = __shfl_sync(0xffffffffu, qy, k);
        }
        // area: triangle (0,1,2) on even lanes, (2,3,0) on odd lanes
        double tri;
        {
          const int i0 = (lane & 1) ? 2 : 0, i1 = (lane & 1) ? 3 : 1, i2 = (lane & 1) ? 0 : 2;
          const int ia[3] = {i0, i1, i2}, ib[3] = {i1, i2, i0};
          double length[3];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            double ddx = (double)(qp[ib[k]][0] - qp[ia[k]][0]), ddy = (double)(qp[ib[k]][1] - qp[ia[k]][1]);
            length[k] = sqrt(ddx * ddx + ddy * ddy);
          }
          const double p = (length[0] + length[1] + length[2]) / 2;
          tri = sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        }
        const double tri1 = __shfl_sync(0xffffffffu, tri, 1);
        const double tri0 = __shfl_sync(0xffffffffu, tri, 0);
        double area = 0;
        area += tri0;
        area += tri1;
        if (area < 0.95 * fp.tag_width * fp.tag_width) ok = false;
        // interior angles / winding: corner li
        {
          const double ccr = (double)fp.cos_critical_rad;
          const int i0 = li, i1 = (li + 1) & 3, i2 = (li + 2) & 3;
          double dx1 = (double)(qp[i1][0] - qp[i0][0]);
          double dy1 = (double)(qp[i1][1] - qp[i0][1]);
          double dx2 = (double)(qp[i2][0] - qp[i1][0]);
          double dy2 = (double)(qp[i2][1] - qp[i1][1]);
          double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
          const bool bad = (cos_dtheta > ccr || cos_dtheta < -ccr) || dx1 * dy2 < dy1 * dx2;
          ok = ok && __all_sync(0xffffffffu, !bad);
        }
        if (ok && lane == 0) {
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
}

// clusters -> kQuadBins size bins (index lists); one thread per cluster, warp-aggregated list allocation
__global__ void __launch_bounds__(256) k_bin_clusters(Geo g, const ClusterRec *__restrict__ clusters, uint32_t *__restrict__ bin_idx,
                                                      uint32_t *__restrict__ counters) {
  const uint32_t ncl = min(counters[CNT_CLUSTERS], g.clu_cap);
  for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < ncl; c0 += gridDim.x * blockDim.x) {
    const uint32_t c = c0 + threadIdx.x;
    int bin = -1;
    if (c < ncl) {
      const uint32_t n = clusters[c].count;
      bin = n <= 128 ? 0 : (n <= 256 ? 1 : (n <= 512 ? 2 : (n <= 1024 ? 3 : (n <= 2048 ? 4 : 5))));
    }
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int b = 0; b < kQuadBins; b++) {
      const unsigned m = __ballot_sync(0xffffffffu, bin == b);
      if (m == 0) continue;
      uint32_t base = 0;
      const int leader = __ffs(m) - 1;
      if ((int)lane == leader) base = atomicAdd(&counters[CNT_BIN0 + b], (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (bin == b) bin_idx[(size_t)b * g.clu_cap + base + __popc(m & ((1u << lane) - 1))] = c;
    }
  }
}

template <int THREADS, int NCAP, bool ALL_SMEM, int ITEMS, int SCAN_CH>
static void launch_bin(const Workspace &ws, int bin, int ctas_per_sm, int sms, const ComboTable &ct, cudaStream_t st) {
  const Geo &g = ws.g;
  constexpr size_t smem = (size_t)(2 * NCAP + 6 * ((ALL_SMEM ? NCAP : SCAN_CH) + 1)) * 8;  // keys/errA + errB|sort scratch + moments|staging
  // the opt-in for > 48 KB of dynamic shared memory is a per-device function attribute
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(k_quadfit<THREADS, NCAP, ALL_SMEM, ITEMS, SCAN_CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set[dev] = true;
  }
  k_quadfit<THREADS, NCAP, ALL_SMEM, ITEMS, SCAN_CH><<<sms * ctas_per_sm, THREADS, smem, st>>>(g, ws.fp, ws.clusters, ws.bin_idx, bin, ws.pts, ws.keys,
                                                                                  ws.lfps, ws.errs, ws.dec, ws.quads, ws.counters, ct,
                                                                                  at_Wp(g));
}

int launch_quadfit(const Workspace &ws, int nframes, cudaStream_t s) {
  (void)nframes;
  const Geo &g = ws.g;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ComboTable ct;
  ct.c = reinterpret_cast<const uchar4 *>(ws.combos);
  for (int i = 0; i < 18; i++) ct.off[i] = ws.combo_off[i];
  k_bin_clusters<<<sms * 2, 256, 0, s>>>(g, ws.clusters, ws.bin_idx, ws.counters);
  // the bins are independent: fork onto side streams; large clusters (the long poles) are issued first
  cudaEventRecord(ws.ev_fork, s);
  for (int i = 0; i < 5; i++) cudaStreamWaitEvent(ws.aux[i], ws.ev_fork, 0);
  // tuning knobs (experiments): B200AT_QF_SCALE scales the CTAs per SM of every bin (leave shared memory free for co-running
  // dense kernels); B200AT_QF_GLOBAL=1 keeps the moments of the <=1024-point bins in the L2-resident scratch instead of
  // shared memory (smaller CTAs, more of them per SM)
  static const double qscale = getenv("B200AT_QF_SCALE") ? atof(getenv("B200AT_QF_SCALE")) : 1.0;
  static const bool qglobal = getenv("B200AT_QF_GLOBAL") != nullptr;
  auto sc = [&](int c) { return std::max(1, (int)(c * qscale + 0.5)); };
  launch_bin<256, 4096, false, 16, 512>(ws, 5, sc(2), sms, ct, s);          // n > 2048 (n > 4096: global-memory sort fallback)
  launch_bin<256, 2048, false, 8, 512>(ws, 4, sc(3), sms, ct, ws.aux[0]);   // n <= 2048
  if (!qglobal) {
    launch_bin<128, 1024, true, 8, 512>(ws, 3, sc(3), sms, ct, ws.aux[1]);  // n <= 1024
    launch_bin<64, 512, true, 8, 512>(ws, 2, sc(6), sms, ct, ws.aux[2]);    // n <= 512
    launch_bin<32, 256, true, 8, 512>(ws, 1, sc(10), sms, ct, ws.aux[3]);   // n <= 256
  } else {
    launch_bin<128, 1024, false, 8, 128>(ws, 3, sc(5), sms, ct, ws.aux[1]);
    launch_bin<64, 512, false, 8, 128>(ws, 2, sc(12), sms, ct, ws.aux[2]);
    launch_bin<32, 256, false, 8, 64>(ws, 1, sc(18), sms, ct, ws.aux[3]);
  }
  launch_bin<32, 128, true, 4, 512>(ws, 0, sc(16), sms, ct, ws.aux[4]);     // n <= 128
  for (int i = 0; i < 5; i++) {
    cudaEventRecord(ws.ev_join[i], ws.aux[i]);
    cudaStreamWaitEvent(s, ws.ev_join[i], 0);
  }
  return 7;
}

}  // namespace b200at
