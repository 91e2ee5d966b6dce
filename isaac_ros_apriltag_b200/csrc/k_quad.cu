// k_quad.cu -- fit_quads (a10), the bit-exact form (B200AT_TUNE=qf_exact=1; the default is the windowed pipeline of k_quad2.cu).
// Restates AprilRobotics fit_quad / ptsort / compute_lfps / quad_segment_maxima / fit_line
// (apriltag_quad_thresh.c) with the SAME floating-point types and evaluation order as the
// CPU oracle, so accept/reject decisions and the float corners are bit-identical (built with -fmad=false).
//
// One CTA per boundary-point cluster; clusters are binned by size so that CTA shape and shared-memory footprint fit the
// work, and every per-point array has a compile-time address space (no generic loads):
//   QF_ALL   n <= 128 / 256 (one warp), <= 512 (2 warps), <= 1024 (4 warps): keys, errors AND the six prefix moments in
//            shared memory (64 B per point)
//   QF_KEYS  n <= 2048 / 4096 (8 warps), <= 8192 (16 warps): keys and errors in shared memory (16 B per point), prefix
//            moments in an L2-resident scratch; the serial prefix sum is software-pipelined: warp 0 scans chunk c out of one
//            half of a double-buffered staging ring while the other warps produce the terms of chunk c+1 into the other half
//   QF_GLOBAL n > 8192 (4K-class frames only): every per-point array in global memory
// Per cluster: slope keys (float) -> merge sort of u64 (slope|y|x) keys -> line-fit terms -> SEQUENTIAL double prefix sums
// (a parallel scan would change the roundings; the six moments run as six lanes) -> per-point window error -> 7-tap
// smoothing -> local maxima -> top-(max_nmaxima) by rank counting -> pair table of line fits -> C(n,4) search over a
// precomputed combination table with lexicographic arg-min -> corners, area and angle gates.
// Latency-bound stage (O(boundary points)); HBM bytes are ~12 B per point (read packed point, write sorted key).
#include <algorithm>
#include <cstdlib>

#include "k_quad_common.cuh"

namespace b200at {

enum { QF_ALL = 0, QF_KEYS = 1, QF_GLOBAL = 2 };

template <int NCAP, int MODE, int CH>
struct QfSmem {
  static constexpr int TBL = MAXM * MAXM;  // entries per pair table (mse, nx, ny)
  // region 0: keys (later errA, maxima) | sort scratch (later staging ring / errB); the pair tables alias its start
  static constexpr int R0 = MODE == QF_GLOBAL ? 3 * TBL : (2 * NCAP > 3 * TBL ? 2 * NCAP : 3 * TBL);
  static constexpr int WORDS = R0 + (MODE == QF_ALL ? 6 * (NCAP + 1) : 0);
  static constexpr size_t BYTES = (size_t)WORDS * 8;
  static_assert(MODE != QF_KEYS || 2 * 6 * (CH + 1) <= NCAP, "staging ring must fit the sort scratch");
};

// CPB > 1 (one-warp clusters only): CPB clusters per CTA, one per warp, each with its own shared-memory slice, and the warps
// of a CTA walk through the phases TOGETHER (block barriers at the phase boundaries, rejected clusters idle instead of
// leaving).  The kernel is ~80 KB of straight-line code per cluster; warps that sit in the same phase share the instruction
// lines they fetch, warps in unrelated phases (18 independent one-warp CTAs per SM) thrash the 32 KB instruction cache.
template <int THREADS, int NCAP, int MODE, int ITEMS, int CH, int MINB, int CPB, bool LOOKAHEAD>
__global__ void __launch_bounds__(THREADS * CPB, MINB) k_quadfit(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters,
                                                     const uint32_t *__restrict__ bin_idx, int bin, const uint32_t *__restrict__ pts,
                                                     unsigned long long *__restrict__ keys, LineFitPt *__restrict__ lfps_pool,
                                                     double *__restrict__ errs_pool, const uint8_t *__restrict__ dec,
                                                     QuadRec *__restrict__ quads, uint32_t *__restrict__ counters, ComboTable combos,
                                                     int Wp) {
  constexpr int NW = THREADS / 32;
  constexpr bool SM = MODE != QF_GLOBAL;      // per-point keys / errors in shared memory
  constexpr bool MSM = MODE == QF_ALL;        // prefix moments in shared memory
  constexpr int PPT = SM ? NCAP / THREADS : 1;  // points per thread (register cache between the bbox and the key pass)
  using L = QfSmem<NCAP, MODE, CH>;
  constexpr int MSTRIDE = NCAP + 1;
  static_assert(MODE != QF_KEYS || NW >= 2, "the pipelined scan needs producer warps");
  static_assert(CPB == 1 || THREADS == 32, "several clusters per CTA: one-warp clusters only (warp-level barriers inside a cluster)");
  extern __shared__ unsigned long long dsm_all[];
  const int grp = CPB > 1 ? (int)(threadIdx.x / THREADS) : 0;   // cluster slot of this thread inside the CTA
  unsigned long long *dsm = dsm_all + (size_t)grp * L::WORDS;
  unsigned long long *skeys = dsm;                              // [NCAP]  keys (later: errA, maxima)
  double *s_ring = reinterpret_cast<double *>(dsm + NCAP);      // [NCAP]  sort scratch, staging ring (QF_KEYS), errB
  double *s_M = reinterpret_cast<double *>(dsm + L::R0);        // [6][MSTRIDE] (QF_ALL)
  double *pt_mse = reinterpret_cast<double *>(dsm), *pt_nx = pt_mse + L::TBL, *pt_ny = pt_nx + L::TBL;
  __shared__ BBoxRed s_red_a[CPB][NW];
  __shared__ int s_cluster_a[CPB];
  __shared__ int s_fm_a[CPB][MAXM];
  __shared__ int s_scan_a[CPB][NW];
  __shared__ double s_rv_a[CPB][NW];
  __shared__ int s_ri_a[CPB][NW];
  __shared__ unsigned int s_rr_a[CPB][NW];
  __shared__ double s_thresh_a[CPB];
  BBoxRed *s_red = s_red_a[grp];
  int &s_cluster = s_cluster_a[grp];
  int *s_fm = s_fm_a[grp], *s_scan = s_scan_a[grp], *s_ri = s_ri_a[grp];
  double *s_rv = s_rv_a[grp];
  unsigned int *s_rr = s_rr_a[grp];
  double &s_thresh = s_thresh_a[grp];

  const int tid = CPB > 1 ? (int)(threadIdx.x % THREADS) : (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t nbin = min(counters[CNT_BIN0 + bin], g.clu_cap);
  // CPB == 1: a rejected cluster leaves the iteration (continue).  CPB > 1: it stays, idle, so that every warp of the CTA
  // reaches the phase barriers.
#define QF_DROP()           \
  {                         \
    if (CPB == 1) continue; \
    alive = false;          \
    sz = 0; /* every later per-point loop is empty, the per-cluster tails see no maxima / no combination */ \
  }
#define QF_PHASE() \
  if (CPB > 1) __syncthreads()

  for (;;) {
    cta_sync<THREADS>();
    if (tid == 0) s_cluster = (int)atomicAdd(&counters[CNT_WORK0 + bin], 1u);
    cta_sync<THREADS>();
    const int cw = s_cluster;
    bool alive = (uint32_t)cw < nbin;
    if (CPB > 1) {
      if (!__syncthreads_or(alive ? 1 : 0)) break;  // every cluster slot of the CTA ran out of work
    } else if (!alive) {
      break;
    }
    const ClusterRec cr = alive ? clusters[bin_idx[(size_t)bin * g.clu_cap + cw]] : ClusterRec{0ull, 0u, 0u, 0u, 0u};
    int sz = (int)cr.count;
    const uint32_t o = cr.offset;
    const uint8_t *im = dec + (size_t)cr.frame * g.Hd * Wp;
    unsigned long long *keys_g = keys + o;
    LineFitPt *lfps_g = lfps_pool + o;

    // ---- Phase A: bounding box + integer sums for the border-polarity test ----
    uint32_t pr[PPT];
    BBoxRed r = {1 << 30, -1, 1 << 30, -1, 0, 0, 0};
    if (SM) {
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        const int i = tid + k * THREADS;
        pr[k] = i < sz ? pts[o + i] : 0u;
      }
#pragma unroll
      for (int k = 0; k < PPT; k++)
        if (tid + k * THREADS < sz) bbox_add(r, pr[k]);
    } else {
      for (int i = tid; i < sz; i += THREADS) bbox_add(r, pts[o + i]);
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
    if (alive && (r.xmax - r.xmin) * (r.ymax - r.ymin) < fp.tag_width) QF_DROP();
    const float cx = (float)((r.xmin + r.xmax) * 0.5 + 0.05118);
    const float cy = (float)((r.ymin + r.ymax) * 0.5 + -0.028581);
    // dot = sum (x-cx)*gx + (y-cy)*gy, evaluated exactly on the integer parts (order independent)
    const double dotd = (double)r.s1 - (double)cx * (double)r.sgx - (double)cy * (double)r.sgy;
    const bool reversed = dotd < 0;
    if (alive && !fp.reversed_border && reversed) QF_DROP();
    if (alive && !fp.normal_border && !reversed) QF_DROP();

    QF_PHASE();
    // ---- Phase C: sort keys (slope | y | x) ----
    if (SM) {
#pragma unroll
      for (int k = 0; k < PPT; k++) {
        const int i = tid + k * THREADS;
        if (i < sz) skeys[i] = slope_key(pr[k], cx, cy);
      }
      cta_sync<THREADS>();
      sort_keys<THREADS, ITEMS, LOOKAHEAD>(skeys, reinterpret_cast<unsigned long long *>(s_ring), sz, tid);
    } else {
      for (int i = tid; i < sz; i += THREADS) keys_g[i] = slope_key(pts[o + i], cx, cy);
      cta_sync<THREADS>();
      // scratch for the merge passes: the (not yet used) error area
      sort_keys<THREADS, ITEMS, LOOKAHEAD>(keys_g, reinterpret_cast<unsigned long long *>(errs_pool + (size_t)2 * o), sz, tid);
    }

    QF_PHASE();
    // ---- Phase E: line-fit terms, then SEQUENTIAL prefix sums (six lanes, one per moment) ----
    if (MODE == QF_ALL) {
      // terms of all points straight into the moment arrays (two points per iteration: eight gathers in flight)
      for (int i = tid; i < sz; i += 2 * THREADS) {
        const int i2 = i + THREADS;
        const bool h2 = i2 < sz;
        const unsigned long long k0 = skeys[i], k1 = h2 ? skeys[i2] : 0ull;
        const int ga = grad2_at(im, Wp, g.Wd, g.Hd, k0);
        const int gb = h2 ? grad2_at(im, Wp, g.Wd, g.Hd, k1) : 0;
        keys_g[i] = k0;
        double t[6];
        lfp_terms(k0, ga, t);
#pragma unroll
        for (int m = 0; m < 6; m++) s_M[m * MSTRIDE + i] = t[m];
        if (h2) {
          keys_g[i2] = k1;
          lfp_terms(k1, gb, t);
#pragma unroll
          for (int m = 0; m < 6; m++) s_M[m * MSTRIDE + i2] = t[m];
        }
      }
      cta_sync<THREADS>();
      if (tid < 6) {
        double *base = s_M + tid * MSTRIDE;
        scan_chain(base, sz, 0.0, [&](int i, double v) { base[i] = v; });
      }
      cta_sync<THREADS>();
    } else if (MODE == QF_KEYS) {
      // gradient pass: the slope half of a sorted key is dead, it now carries the squared gradient magnitude
      for (int i = tid; i < sz; i += 4 * THREADS) {
        unsigned long long k[4];
        int g2[4];
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = (i + u * THREADS < sz) ? skeys[i + u * THREADS] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) g2[u] = (i + u * THREADS < sz) ? grad2_at(im, Wp, g.Wd, g.Hd, k[u]) : 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (i + u * THREADS < sz) {
            keys_g[i + u * THREADS] = k[u];
            skeys[i + u * THREADS] = (k[u] & 0xffffffffull) | ((unsigned long long)(uint32_t)g2[u] << 32);
          }
        }
      }
      __syncthreads();
      // software-pipelined scan: warp 0 scans chunk c out of ring half (c & 1) and streams the prefix moments to the
      // L2-resident scratch; the other warps fill half ((c + 1) & 1) with the terms of chunk c + 1
      constexpr int HALF = 6 * (CH + 1);
      auto produce = [&](int c, int t0, int nt) {
        double *b = s_ring + (c & 1) * HALF;
        const int c0 = c * CH, cn = min(CH, sz - c0);
        for (int ii = t0; ii < cn; ii += nt) {
          const unsigned long long k = skeys[c0 + ii];
          double t[6];
          lfp_terms(k, (int)(k >> 32), t);
#pragma unroll
          for (int m = 0; m < 6; m++) b[m * (CH + 1) + ii] = t[m];
        }
      };
      const int nchunks = (sz + CH - 1) / CH;
      produce(0, tid, THREADS);
      __syncthreads();
      double acc = 0.0;
      for (int c = 0; c < nchunks; c++) {
        if (wid == 0) {
          if (lane < 6) {
            const double *b = s_ring + (c & 1) * HALF + lane * (CH + 1);
            double *gp = reinterpret_cast<double *>(lfps_g + (size_t)c * CH) + lane;
            acc = scan_chain(b, min(CH, sz - c * CH), acc, [&](int i, double v) { gp[(size_t)i * 6] = v; });
          }
        } else if (c + 1 < nchunks) {
          produce(c + 1, tid - 32, THREADS - 32);
        }
        __syncthreads();
      }
    } else {
      for (int i = tid; i < sz; i += THREADS) {
        const unsigned long long k = keys_g[i];
        double t[6];
        lfp_terms(k, grad2_at(im, Wp, g.Wd, g.Hd, k), t);
        double *gp = reinterpret_cast<double *>(lfps_g + i);
#pragma unroll
        for (int m = 0; m < 6; m++) gp[m] = t[m];
      }
      __syncthreads();
      if (tid < 6) {
        double *gp = reinterpret_cast<double *>(lfps_g) + tid;
        double acc = 0.0;
        for (int i = 0; i < sz; i++) {
          acc += gp[(size_t)i * 6];
          gp[(size_t)i * 6] = acc;
        }
      }
      __syncthreads();
    }
    LfAcc<MSM> lf;
    lf.m = s_M;
    lf.stride = MSTRIDE;
    lf.g = lfps_g;

    QF_PHASE();
    // ---- Phase F/G: per-point line-fit error over a +-ksz window, then 7-tap smoothing (circular) ----
    const int ksz = min(20, sz / 12);
    if (ksz < 2) QF_DROP();
    // errA takes over the (now dead) key area, errB the sort scratch / staging ring
    double *errA = SM ? reinterpret_cast<double *>(skeys) : (errs_pool + (size_t)2 * o);
    double *errB = SM ? s_ring : (errs_pool + (size_t)2 * o + sz);
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

    QF_PHASE();
    // ---- Phase H: local maxima, compacted in index order (errA area is dead again: reuse it) ----
    double *merr = errA;                                                  // values of the maxima
    uint32_t *midx = reinterpret_cast<uint32_t *>(errA + (sz + 1) / 2);    // their indices
    const int nmax_all = ordered_compact<NW>(
        sz, lane, wid, s_scan,
        [&](int i) {
          const double e = errB[i];
          return e > errB[i + 1 == sz ? 0 : i + 1] && e > errB[i == 0 ? sz - 1 : i - 1];
        },
        [&](int i, int pos) {
          midx[pos] = (uint32_t)i;
          merr[pos] = errB[i];
        });
    if (nmax_all < 4) QF_DROP();
    cta_sync<THREADS>();

    // ---- Phase I: keep the max_nmaxima best maxima: threshold = value of descending rank max_nmaxima ----
    int nm;
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
      const int kept = ordered_compact<NW>(
          nmax_all, lane, wid, s_scan, [&](int i) { return !(errB[midx[i]] <= maxima_thresh); },
          [&](int i, int pos) {
            if (pos < MAXM) s_fm[pos] = (int)midx[i];
          });
      nm = min(kept, MAXM);
    } else {
      if (tid < nmax_all) s_fm[tid] = (int)midx[tid];
      nm = nmax_all;
    }
    cta_sync<THREADS>();
    if (nm < 4) QF_DROP();

    QF_PHASE();
    // ---- Phase J: line fits between every ordered pair of kept maxima (tables over the dead key / error area) ----
    for (int t = tid; t < nm * nm; t += THREADS) {
      int a = t / nm, b = t - a * nm;
      if (a == b) continue;
      double lp[4], m;
      fit_line_dev(lf, sz, s_fm[a], s_fm[b], lp, nullptr, &m);
      pt_mse[a * MAXM + b] = m;
      pt_nx[a * MAXM + b] = lp[2];
      pt_ny[a * MAXM + b] = lp[3];
    }
    cta_sync<THREADS>();

    // ---- Phase K: best (m0<m1<m2<m3); ties resolved to the lexicographically first, like the serial loops ----
    double best = CUDART_INF;
    uint32_t brank = 0xffffffffu;
    {
      const double max_mse = (double)fp.max_line_fit_mse, max_dot = (double)fp.cos_critical_rad;
      const uchar4 *ctab = combos.c + combos.off[nm];
      const int ncomb = combos.off[nm + 1] - combos.off[nm];
      // err of a segment = N * mse (fit_line), N = number of points on the circular index range [i0, i1]
      auto seg_err = [&](int a, int b, double mse) {
        const int i0 = s_fm[a], i1 = s_fm[b];
        const int N = i0 < i1 ? i1 - i0 + 1 : sz - i0 + i1 + 1;
        return N * mse;
      };
      for (int t = tid; t < ncomb; t += THREADS) {
        const uchar4 c = ctab[t];
        const int m0 = c.x, m1 = c.y, m2 = c.z, m3 = c.w;
        const double mse01 = pt_mse[m0 * MAXM + m1];
        if (mse01 > max_mse) continue;
        const double mse12 = pt_mse[m1 * MAXM + m2];
        if (mse12 > max_mse) continue;
        double dot = pt_nx[m0 * MAXM + m1] * pt_nx[m1 * MAXM + m2] + pt_ny[m0 * MAXM + m1] * pt_ny[m1 * MAXM + m2];
        if (fabs(dot) > max_dot) continue;
        const double mse23 = pt_mse[m2 * MAXM + m3];
        if (mse23 > max_mse) continue;
        const double mse30 = pt_mse[m3 * MAXM + m0];
        if (mse30 > max_mse) continue;
        double err = seg_err(m0, m1, mse01) + seg_err(m1, m2, mse12) + seg_err(m2, m3, mse23) + seg_err(m3, m0, mse30);
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

    QF_PHASE();
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
                                                      uint32_t *__restrict__ counters, uint32_t *__restrict__ qinfo) {
  const uint32_t ncl = min(counters[CNT_CLUSTERS], g.clu_cap);
  for (uint32_t c0 = blockIdx.x * blockDim.x; c0 < ncl; c0 += gridDim.x * blockDim.x) {
    const uint32_t c = c0 + threadIdx.x;
    int bin = -1;
    if (c < ncl) {
      const uint32_t n = clusters[c].count, o = clusters[c].offset;
      // ceil(log2(n)) - 7 clamped to [0, kQuadBins - 1]: <=128, <=256, ..., <=8192, larger
      bin = n <= 128 ? 0 : min(kQuadBins - 1, 32 - __clz((int)n - 1) - 7);
      // empty records mark reservations that did not fit the pools (k_cluster_select); anything that does not lie inside
      // the point pool is never handed to the quad fit
      if (n == 0 || n > g.max_cluster_pts || o > g.pts_cap || n > g.pts_cap - o) bin = -1;
      if (qinfo) qinfo[c] = 0u;  // (windowed quad fit) k_qf_sort overwrites it for the clusters that reach the sort
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

#undef QF_DROP
#undef QF_PHASE

template <int THREADS, int NCAP, int MODE, int ITEMS, int CH, int MINB, int CPB, bool LOOKAHEAD>
static void launch_bin_t(const Workspace &ws, int bin, double scale, int sms, const ComboTable &ct, cudaStream_t st) {
  const Geo &g = ws.g;
  constexpr size_t smem = QfSmem<NCAP, MODE, CH>::BYTES * CPB;
  auto kern = k_quadfit<THREADS, NCAP, MODE, ITEMS, CH, MINB, CPB, LOOKAHEAD>;
  // the opt-in for > 48 KB of dynamic shared memory is a per-device function attribute; the persistent grid is sized to
  // the number of CTAs that are resident at once
  static int ctas_per_sm[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev >= 0 && dev < 64 ? dev : 0;
  if (!ctas_per_sm[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS * CPB, smem);
    ctas_per_sm[dev] = std::max(1, n);
  }
  const int grid = std::max(1, (int)(sms * ctas_per_sm[dev] * scale + 0.5));
  kern<<<grid, THREADS * CPB, smem, st>>>(g, ws.fp, ws.clusters, ws.bin_idx, bin, ws.pts, ws.keys, ws.lfps, ws.errs, ws.dec, ws.quads,
                                    ws.counters, ct, at_Wp(g));
}

template <int THREADS, int NCAP, int MODE, int ITEMS, int CH, int MINB, int CPB = 1>
static void launch_bin(const Workspace &ws, int bin, double scale, int sms, const ComboTable &ct, cudaStream_t st) {
  launch_bin_t<THREADS, NCAP, MODE, ITEMS, CH, MINB, CPB, false>(ws, bin, scale, sms, ct, st);
}

void launch_bin_clusters(const Workspace &ws, int sms, cudaStream_t s) {
  k_bin_clusters<<<sms * 2, 256, 0, s>>>(ws.g, ws.clusters, ws.bin_idx, ws.counters, ws.tune.qf_exact ? nullptr : ws.qinfo);
}

int launch_quadfit(const Workspace &ws, int nframes, cudaStream_t s) {
  if (!ws.tune.qf_exact) return launch_quadfit_windowed(ws, nframes, s);
  const Geo &g = ws.g;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ComboTable ct;
  ct.c = reinterpret_cast<const uchar4 *>(ws.combos);
  for (int i = 0; i < 18; i++) ct.off[i] = ws.combo_off[i];
  launch_bin_clusters(ws, sms, s);
  // the bins are independent: fork onto side streams; large clusters (the long poles) are issued first
  cudaEventRecord(ws.ev_fork, s);
  for (int i = 0; i < kQuadAux; i++) cudaStreamWaitEvent(ws.aux[i], ws.ev_fork, 0);
  // (configuration measured best in round 1, profiles/r02_sweep_*.md: 512 / 1024-point bins with the prefix moments in the
  // L2-resident scratch, one-warp bins with several clusters per CTA in phase lockstep)
  const double qscale = 1.0;
  launch_bin<256, 0, QF_GLOBAL, 16, 1, 2>(ws, 7, qscale, sms, ct, s);               // n > 8192
  launch_bin<512, 8192, QF_KEYS, 16, 480, 1>(ws, 6, qscale, sms, ct, ws.aux[0]);    // n <= 8192
  launch_bin<256, 4096, QF_KEYS, 16, 224, 3>(ws, 5, qscale, sms, ct, ws.aux[1]);    // n <= 4096
  launch_bin<256, 2048, QF_KEYS, 8, 160, 4>(ws, 4, qscale, sms, ct, ws.aux[2]);     // n <= 2048
  launch_bin<128, 1024, QF_KEYS, 8, 64, 8>(ws, 3, qscale, sms, ct, ws.aux[3]);      // n <= 1024
  launch_bin<64, 512, QF_KEYS, 8, 32, 16>(ws, 2, qscale, sms, ct, ws.aux[4]);       // n <= 512
  launch_bin<32, 256, QF_ALL, 8, 1, 2, 6>(ws, 1, qscale, sms, ct, ws.aux[5]);       // n <= 256: 6 clusters per CTA
  launch_bin<32, 128, QF_ALL, 4, 1, 2, 8>(ws, 0, qscale, sms, ct, ws.aux[6]);       // n <= 128: 8 clusters per CTA
  for (int i = 0; i < kQuadAux; i++) {
    cudaEventRecord(ws.ev_join[i], ws.aux[i]);
    cudaStreamWaitEvent(s, ws.ev_join[i], 0);
  }
  return 1 + kQuadBins;
}

}  // namespace b200at
