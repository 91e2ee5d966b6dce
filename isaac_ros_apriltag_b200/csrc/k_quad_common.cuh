// k_quad_common.cuh -- device helpers shared by the two quad-fit implementations (k_quad.cu: bit-exact, one CTA per cluster;
// k_quad2.cu: three-kernel pipeline with windowed moments): slope keys, the shared-memory merge sort, line-fit arithmetic
// (AprilRobotics fit_line / compute_lfps, apriltag_quad_thresh.c) with the oracle's types and evaluation order.
#pragma once
#include <math_constants.h>

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
      // 48-byte record, 16-byte aligned: three 128-bit loads
      const double2 *p = reinterpret_cast<const double2 *>(g + i);
      const double2 a = p[0], b = p[1], c = p[2];
      r.Mx = a.x;
      r.My = a.y;
      r.Mxx = b.x;
      r.Mxy = b.y;
      r.Myy = c.x;
      r.W = c.y;
    }
    return r;
  }
};

// fit_line's arithmetic on the six weighted moments of a point range of N points (oracle/apriltag_oracle.cpp fit_line; upstream
// apriltag_quad_thresh.c): covariance, eigen decomposition with float square roots, line = (Ex, Ey, unit normal)
__device__ __forceinline__ void fit_moments_dev(double Mx, double My, double Mxx, double Mxy, double Myy, double W, int N, double *lineparm,
                                                double *err, double *mse) {
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
  fit_moments_dev(Mx, My, Mxx, Mxy, Myy, W, N, lineparm, err, mse);
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
// O(n log^2 n) of a bitonic network, one barrier per pass.  Inlined: the address space of the buffers (shared or global)
// is known at every call site.
// The sorted sequence ends in `a`.
// (Tried and measured slower on B200: register bitonic networks for the per-thread sort and for the merge step -- reading the
// next ITEMS keys of both runs at once and merging min(A[k], B[ITEMS-1-k]) -- instead of the serial merge: +2 % quad-fit time.)
// LOOKAHEAD (qf_sort=1, not measured yet): (a) the serial merge keeps the NEXT key of both runs in registers, so the load that
// refills a side is issued one output ahead of its use instead of sitting on the critical path of every output (ncu: half of
// sort_keys' stall samples in the multi-warp bins are short-scoreboard waits on exactly these dependent loads); (b) merge
// passes whose runs are no longer than a warp's 32 * ITEMS keys only read what the same warp wrote in the pass before, so they
// are separated by __syncwarp() instead of a block barrier (23-30 % of the sort's samples are barrier waits).
template <int THREADS, int ITEMS, bool LOOKAHEAD>
__device__ __forceinline__ void sort_keys(unsigned long long *a, unsigned long long *tmp, int n, int tid) {
  constexpr unsigned long long INF = ~0ull;
  for (int base = tid * ITEMS; base < n; base += THREADS * ITEMS) {
    unsigned long long r[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) r[k] = (base + k < n) ? a[base + k] : INF;
    {
#pragma unroll
      for (int pass = 0; pass < ITEMS; pass++) {
#pragma unroll
        for (int k = (pass & 1); k + 1 < ITEMS; k += 2) {
          unsigned long long x = r[k], y = r[k + 1];
          r[k] = x < y ? x : y;
          r[k + 1] = x < y ? y : x;
        }
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
      if (LOOKAHEAD) {
        unsigned long long na1 = ia + 1 < na ? src[a0 + ia + 1] : INF, nb1 = ib + 1 < nb ? src[b0 + ib + 1] : INF;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
          if (ob + k < b1) {
            const bool take_a = va <= vb;
            dst[ob + k] = take_a ? va : vb;
            if (take_a) {
              ia++;
              va = na1;
              na1 = ia + 1 < na ? src[a0 + ia + 1] : INF;  // needed two outputs from now at the earliest
            } else {
              ib++;
              vb = nb1;
              nb1 = ib + 1 < nb ? src[b0 + ib + 1] : INF;
            }
          }
        }
      } else {
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
    }
    // (the next pass merges runs of 2 * width keys into 4 * width: warp-local as long as 4 * width <= 32 * ITEMS)
    if (LOOKAHEAD && THREADS > 32 && 4 * width <= 32 * ITEMS)
      __syncwarp();
    else
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

__device__ __forceinline__ void bbox_add(BBoxRed &r, uint32_t p) {
  const int x = p & 0x3fff, y = (p >> 14) & 0x3fff;
  const int cxg = (p >> 28) & 3, cyg = (p >> 30) & 3;
  const int gx = cxg == 0 ? 0 : (cxg == 1 ? 255 : -255), gy = cyg == 0 ? 0 : (cyg == 1 ? 255 : -255);
  r.xmin = min(r.xmin, x);
  r.xmax = max(r.xmax, x);
  r.ymin = min(r.ymin, y);
  r.ymax = max(r.ymax, y);
  r.sgx += gx;
  r.sgy += gy;
  r.s1 += (long long)x * gx + (long long)y * gy;
}

// sort key of one boundary point: float slope (monotone angle measure around the bounding-box centre) | y | x
__device__ __forceinline__ unsigned long long slope_key(uint32_t p, float cx, float cy) {
  const int x = p & 0x3fff, y = (p >> 14) & 0x3fff;
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
  const float slope = quadrant + dy / dx;
  return ((unsigned long long)float_orderable(slope) << 32) | ((unsigned long long)y << 16) | (unsigned long long)x;
}

// ---------------------------------------------------------------------------------------------------------------------
// Angular bucket of a sort key (k_quad2.cu, bucket sort): a NON-DECREASING function of the key's float slope into [0, NB), so
// that bucket order never contradicts key order -- the final order is decided by full 64-bit key comparisons inside a bucket
// and is therefore the same total order every other sort produces.  The slope is quadrant base + tan-like ratio r >= 0; the
// map r -> r / 2 (r <= 1), 1 - 1 / (2 r) (r > 1) is a cheap stand-in for atan that spreads a convex outline over the buckets
// within a factor of ~2.  Every step (float subtract, correctly rounded divide, add, multiply by a positive constant,
// truncation) is monotone, so rounding cannot reorder two keys.  nbq = NB / 4.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int key_bucket(unsigned long long key, float nbq, int NB) {
  const uint32_t ob = (uint32_t)(key >> 32);
  const float s = __uint_as_float((ob & 0x80000000u) ? (ob & 0x7fffffffu) : ~ob);  // undo float_orderable
  float qi, Q;
  if (s < 0.0f) {
    qi = 0.0f;
    Q = -65536.0f;
  } else if (s < 65536.0f) {
    qi = 1.0f;
    Q = 0.0f;
  } else if (s < 131072.0f) {
    qi = 2.0f;
    Q = 65536.0f;
  } else {
    qi = 3.0f;
    Q = 131072.0f;
  }
  const float r = s - Q;
  const float t = r <= 1.0f ? 0.5f * r : 1.0f - 0.5f / r;
  const int b = (int)((qi + t) * nbq);
  return max(0, min(b, NB - 1));
}

// ---------------------------------------------------------------------------------------------------------------------
// Register bitonic sort of 32 * E 64-bit keys per warp (k_quad2.cu).  Element i of the warp's chunk lives in lane i / E, register
// i % E, so the low log2(E) index bits are compare-exchanges between registers (one 64-bit compare = ISETP + ISETP.EX, four
// SELs per pair) and the five lane bits are shuffles (two SHFLs, the compare, two SELs per key).  Ascending-only network: the
// first step of every merge level pairs i with i ^ (2^k - 1), the following ones i with i ^ j.
// (Measured dead end: keys packed as positive doubles with fmin / fmax as the compare-exchange -- sm_100a has no DMNMX, the
// pair compiles to DSETP + selects + NaN fix-ups and the network executed 1.7x the instructions of the merge sort it replaced.)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void key_ce(unsigned long long &a, unsigned long long &b) {  // a <= b afterwards
  const bool sw = b < a;
  const unsigned long long t = a;
  a = sw ? b : a;
  b = sw ? t : b;
}

template <int E>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&v)[E], int lane) {
  constexpr int N = 32 * E;
#pragma unroll
  for (int k2 = 2; k2 <= N; k2 <<= 1) {
    // step 1: partner i ^ (k2 - 1)
    if (k2 <= E) {
#pragma unroll
      for (int r = 0; r < E; r++) {
        const int q = r ^ (k2 - 1);
        if (q > r) key_ce(v[r], v[q]);
      }
    } else {
      const int lm = k2 / E - 1;                        // lane bits flipped
      const bool lower = (lane & (k2 / E / 2)) == 0;    // the highest flipped bit of my index is 0: I keep the minimum
      // my register r pairs with the partner's register E - 1 - r: walk the registers from both ends so that no copy of the old
      // values is needed
#pragma unroll
      for (int r = 0; r < E / 2; r++) {
        const unsigned long long pa = __shfl_xor_sync(0xffffffffu, v[E - 1 - r], lm);  // partner of v[r]
        const unsigned long long pb = __shfl_xor_sync(0xffffffffu, v[r], lm);          // partner of v[E - 1 - r]
        v[r] = ((pa < v[r]) == lower) ? pa : v[r];
        v[E - 1 - r] = ((pb < v[E - 1 - r]) == lower) ? pb : v[E - 1 - r];
      }
    }
    // following steps: partner i ^ j
#pragma unroll
    for (int j = k2 >> 2; j >= 1; j >>= 1) {
      if (j < E) {
#pragma unroll
        for (int r = 0; r < E; r++)
          if ((r & j) == 0) key_ce(v[r], v[r | j]);
      } else {
        const int lm = j / E;
        const bool lower = (lane & lm) == 0;
#pragma unroll
        for (int r = 0; r < E; r++) {
          const unsigned long long p = __shfl_xor_sync(0xffffffffu, v[r], lm);
          v[r] = ((p < v[r]) == lower) ? p : v[r];
        }
      }
    }
  }
}

// merge passes of sort_keys from sorted runs of `width0` keys (the last run may be partial); the sorted sequence ends in `a`
template <int THREADS, int ITEMS>
__device__ __forceinline__ void merge_runs(unsigned long long *a, unsigned long long *tmp, int n, int tid, int width0) {
  constexpr unsigned long long INF = ~0ull;
  unsigned long long *src = a, *dst = tmp;
  for (int width = width0; width < n; width <<= 1) {
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

// squared gradient magnitude of the decimated image at the (half-resolution) point of a sorted key; 0 on the image
// border, where the reference uses weight 1 = sqrt(0) + 1
__device__ __forceinline__ int grad2_at(const uint8_t *__restrict__ im, int Wp, int Wd, int Hd, unsigned long long k) {
  const int px = (int)(k & 0xffff), py = (int)((k >> 16) & 0xffff);
  const int ix = (px + 1) >> 1, iy = (py + 1) >> 1;  // == (int)(px * .5 + 0.5) for the non-negative half-pixel coordinates
  int g2 = 0;
  if (ix > 0 && ix + 1 < Wd && iy > 0 && iy + 1 < Hd) {
    const uint8_t *c = im + (size_t)iy * Wp + ix;
    const int grad_x = (int)c[1] - (int)c[-1];
    const int grad_y = (int)c[Wp] - (int)c[-Wp];
    g2 = grad_x * grad_x + grad_y * grad_y;
  }
  return g2;
}

// the six line-fit terms of one point (compute_lfps): W*x, W*y, W*x*x, W*x*y, W*y*y, W with x = px/2 + 0.5
__device__ __forceinline__ void lfp_terms(unsigned long long k, int g2, double t[6]) {
  const int px = (int)(k & 0xffff), py = (int)((k >> 16) & 0xffff);
  const double fx = px * .5 + 0.5;
  const double fy = py * .5 + 0.5;
  const double W = sqrt((double)g2) + 1;
  t[0] = W * fx;
  t[1] = W * fy;
  t[2] = W * fx * fx;
  t[3] = W * fx * fy;
  t[4] = W * fy * fy;
  t[5] = W;
}

// sequential prefix sum of `cn` terms (one moment per calling lane): acc carries across chunks
template <class Store>
__device__ __forceinline__ double scan_chain(const double *t, int cn, double acc, Store store) {
  int i = 0;
  for (; i + 8 <= cn; i += 8) {
    const double t0 = t[i], t1 = t[i + 1], t2 = t[i + 2], t3 = t[i + 3], t4 = t[i + 4], t5 = t[i + 5], t6 = t[i + 6],
                 t7 = t[i + 7];
    acc += t0;
    store(i, acc);
    acc += t1;
    store(i + 1, acc);
    acc += t2;
    store(i + 2, acc);
    acc += t3;
    store(i + 3, acc);
    acc += t4;
    store(i + 4, acc);
    acc += t5;
    store(i + 5, acc);
    acc += t6;
    store(i + 6, acc);
    acc += t7;
    store(i + 7, acc);
  }
  for (; i < cn; i++) {
    acc += t[i];
    store(i, acc);
  }
  return acc;
}

// Ordered stream compaction of the indices i < n with pred(i): emit(i, position).  Warp w owns a contiguous range of
// 32-element strips; one count pass, one exclusive scan over the warps, one write pass (a single pass for one-warp CTAs).
// Returns the number of selected elements (all threads).  The caller synchronises before consuming the emitted data.
template <int NW, class Pred, class Emit>
__device__ __forceinline__ int ordered_compact(int n, int lane, int wid, int *s_scan, Pred pred, Emit emit) {
  const int strips = (n + 31) >> 5;
  const int spw = (strips + NW - 1) / NW;
  const int s0 = wid * spw, s1 = min(strips, s0 + spw);
  int base = 0, total = 0;
  if (NW > 1) {
    int cnt = 0;
    for (int s = s0; s < s1; s++) {
      const int i = s * 32 + lane;
      cnt += __popc(__ballot_sync(0xffffffffu, i < n && pred(i)));
    }
    if (lane == 0) s_scan[wid] = cnt;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const int v = s_scan[w];
      base += w < wid ? v : 0;
      total += v;
    }
  }
  int run = base;
  for (int s = s0; s < s1; s++) {
    const int i = s * 32 + lane;
    const bool p = i < n && pred(i);
    const unsigned bal = __ballot_sync(0xffffffffu, p);
    if (p) emit(i, run + __popc(bal & ((1u << lane) - 1)));
    run += __popc(bal);
  }
  return NW > 1 ? total : run;
}

// combination table for the current nm: all (m0<m1<m2<m3) < nm in lexicographic order, one byte each
struct ComboTable {
  const uchar4 *c;
  int off[18];
};

}  // namespace b200at
