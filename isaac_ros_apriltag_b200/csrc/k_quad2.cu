// k_quad2.cu -- fit_quads (a10), windowed three-kernel pipeline (the default; k_quad.cu is the bit-exact one-CTA-per-cluster form).
// Restates AprilRobotics fit_quad / ptsort / compute_lfps / quad_segment_maxima / fit_line (apriltag_quad_thresh.c; SURVEY
// App. A.5; oracle/apriltag_oracle.cpp:454-764).
//
// What makes fit_quad expensive per boundary point is (1) the sort by angle and (2) one line fit per point over a +-ksz
// window (ksz <= 20), smoothing and local maxima.  Only (1) needs the whole cluster in one place.  (2) is LOCAL: the window
// moments are differences of prefix sums taken 2*ksz+1 points apart, so any prefix base cancels -- a chunk of consecutive
// sorted points plus a circular halo of ksz+5 / ksz+4 points is enough, and what the later stages need from the global prefix
// sums is their value at <= max_nmaxima points per cluster.  Hence:
//   k_qf_sort   per cluster (size bins): bounding box, polarity, slope keys, shared-memory merge sort; writes the sorted points
//               with their gradient magnitude (8 B / point: y, x, |grad|^2) and registers the cluster's chunks as work items.
//               Shared memory: 16 B / point (the exact kernel keeps 64 B / point resident: keys, errors, six prefix moments).
//   k_qf_window flat over (cluster, chunk) work items, one warp each: line-fit terms, chunk-local prefix moments (lane-serial +
//               one warp scan), window error, 7-tap smoothing, local maxima -> per maximum {index, error, chunk-local prefix
//               moments}, per chunk its moment totals.  No dependence between work items: the 8192-point cluster that took
//               one 512-thread CTA for a millisecond is now 47 independent warps.
//   k_qf_tail   per cluster, one warp: the max_nmaxima best maxima, their global prefix moments (chunk totals + local prefix),
//               pair table of line fits, C(n,4) search, corners, area / angle gates (same code shape as the exact kernel).
// Floating point: every formula is the oracle's; what differs is the ASSOCIATION of the prefix sums (lane-serial + tree instead
// of one serial chain) and one reciprocal-multiply instead of five divisions in the per-point window fit.  Effect: window
// errors agree to ~1e-12 relative; the float quad corners are bit-identical except where a double lands within that distance
// of a float rounding boundary or two maxima tie within it (tests state the tolerance and count the differences).
#include <algorithm>
#include <cstdlib>

#include "k_quad_common.cuh"

namespace b200at {

struct MaxRec {  // one local maximum of the smoothed window error, written by k_qf_window
  uint32_t idx, pad;
  double y;
  double P[6];   // inclusive prefix moments at idx relative to the first point of its chunk
};
static_assert(sizeof(MaxRec) == 64, "MaxRec layout");

// chunk c of nch of a cluster of n sorted points: [s, e)
__device__ __forceinline__ void qf_chunk_bounds(int n, int nch, int c, int &s, int &e) {
  s = (int)((long long)c * n / nch);
  e = (int)((long long)(c + 1) * n / nch);
}
// first record of chunk c in the cluster's maxima region (a chunk of len points has at most ceil(len / 2) maxima)
__device__ __forceinline__ int qf_chunk_rec0(int s, int c) { return (s + 1) / 2 + c; }

// ---------------------------------------------------------------------------------------------------------------------
// k_qf_sort
// ---------------------------------------------------------------------------------------------------------------------
// Bounding box + polarity sums of the cluster over all THREADS threads of its worker (reduced value in every thread).
template <int THREADS>
__device__ __forceinline__ BBoxRed bbox_reduce(BBoxRed r, BBoxRed *s_red, int lane, int wid) {
  constexpr int NW = THREADS / 32;
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
      const BBoxRed q = s_red[w];
      r.xmin = min(r.xmin, q.xmin);
      r.xmax = max(r.xmax, q.xmax);
      r.ymin = min(r.ymin, q.ymin);
      r.ymax = max(r.ymax, q.ymax);
      r.sgx += q.sgx;
      r.sgy += q.sgy;
      r.s1 += q.s1;
    }
  }
  return r;
}

// fit_quad's gates before the sort: bounding-box area, border polarity (the dot product is evaluated exactly on integers)
__device__ __forceinline__ bool qf_pregate(const BBoxRed &r, const FitParams &fp, float &cx, float &cy, bool &reversed) {
  cx = (float)((r.xmin + r.xmax) * 0.5 + 0.05118);
  cy = (float)((r.ymin + r.ymax) * 0.5 + -0.028581);
  const double dotd = (double)r.s1 - (double)cx * (double)r.sgx - (double)cy * (double)r.sgy;
  reversed = dotd < 0;
  bool drop = (r.xmax - r.xmin) * (r.ymax - r.ymin) < fp.tag_width;
  drop = drop || (!fp.reversed_border && reversed) || (!fp.normal_border && !reversed);
  return !drop;
}

// the cluster's chunks become work items of k_qf_window (one thread)
__device__ __forceinline__ void qf_register_work(uint32_t ci, uint32_t off, int sz, bool reversed, uint32_t *qinfo, uint32_t *qwbase,
                                                 uint4 *work, uint32_t work_cap, uint32_t *counters) {
  const int nch = qf_nchunks(sz);
  const uint32_t wb = atomicAdd(&counters[CNT_QWORK], (uint32_t)nch);
  if (wb + (uint32_t)nch <= work_cap) {
    for (int c = 0; c < nch; c++) work[wb + c] = make_uint4(off, (uint32_t)sz, (uint32_t)c, ci);
    qwbase[ci] = wb;
    qinfo[ci] = (uint32_t)sz | (reversed ? 0x80000000u : 0u);
  } else {  // (cannot happen with the capacity capi.cu allocates: pts_cap / kQfChunkMax + clu_cap)
    qinfo[ci] = 0u;
    atomicOr(&counters[CNT_STATUS], (uint32_t)ST_QUADS_FULL);
  }
}

// Clusters of at most THREADS * E points.
// MODE 2 (default): ANGULAR BUCKET SORT.  The sort key is an angle around the bounding-box centre, so a monotone map of the key
// (key_bucket) spreads the cluster over NB >= n / 2 buckets of a few points each: one shared-memory histogram pass (the atomic's
// return value is the point's arrival rank in its bucket), one scan of the NB counters, one scatter, and then every point finds
// its final position by counting the smaller keys of its own bucket -- a loop of ~m independent loads and compares (median
// largest bucket of a cluster on the bench frames: 7 points) instead of log2(n) merge passes of chained loads.  Points of
// neighbouring positions sit in the same or adjacent buckets: the loop's loads are broadcasts.  Any monotone bucket map gives the
// same final order (unique 64-bit keys); a degenerate outline (all points at one angle: a bucket of more than Tune::qf_bucket_limit
// points) falls back to the merge sort in global memory.  Shared memory: 10 B per point instead of 16.
// MODE 1: the warp sorts its 32 * E keys in registers (bitonic network, warp_bitonic_sort), no shared memory.
// MODE 0: shared-memory merge sort (ITEMS keys sorted per thread, then merge-path passes).
// WPC > 1 (one-warp clusters): WPC independent cluster workers per CTA, one warp each, no block-wide barrier anywhere.
template <int THREADS, int E, int MODE, int WPC>
__host__ __device__ constexpr size_t qf_sort_smem() {
  return MODE == 1 ? 0 : MODE == 0 ? (size_t)2 * THREADS * E * 8 * WPC : ((size_t)THREADS * E * 8 + (size_t)(THREADS * E / 2 + 2) * 4) * WPC;
}

template <int THREADS, int E, int ITEMS, int MINB, int WPC, int MODE>
__global__ void __launch_bounds__(THREADS *WPC, MINB)
    k_qf_sort(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ bin_idx, int bin,
              const uint32_t *__restrict__ pts, unsigned long long *__restrict__ keys, double *__restrict__ errs_pool,
              const uint8_t *__restrict__ dec, uint32_t *__restrict__ qinfo, uint32_t *__restrict__ qwbase, uint4 *__restrict__ work,
              uint32_t work_cap, uint32_t *__restrict__ counters, int Wp, int bucket_limit) {
  constexpr int NW = THREADS / 32, NCAP = THREADS * E;
  static_assert(WPC == 1 || THREADS == 32, "several workers per CTA: one-warp clusters only");
  static_assert(MODE != 1 || NW == 1, "register sort: one-warp clusters only");
  extern __shared__ unsigned long long dsm_sort[];  // per worker: MODE 0 [2 * NCAP] keys; MODE 2 [NCAP] keys + [NCAP / 2 + 2] counters
  const int grp = WPC > 1 ? (int)(threadIdx.x / THREADS) : 0;
  unsigned long long *skeys = dsm_sort + (size_t)grp * (qf_sort_smem<THREADS, E, MODE == 1 ? 0 : MODE, 1>() / 8), *stmp = skeys + NCAP;
  uint32_t *cnt = reinterpret_cast<uint32_t *>(stmp);  // MODE 2: bucket counters, then bucket start offsets
  __shared__ BBoxRed s_red[NW];
  __shared__ uint32_t s_ws[WPC][NW], s_wm[WPC][NW];
  __shared__ int s_cluster_a[WPC];
  const int tid = WPC > 1 ? (int)(threadIdx.x % THREADS) : (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t nbin = min(counters[CNT_BIN0 + bin], g.clu_cap);
  for (;;) {
    cta_sync<THREADS>();
    if (tid == 0) s_cluster_a[grp] = (int)atomicAdd(&counters[CNT_WORK0 + bin], 1u);
    cta_sync<THREADS>();
    const int cw = s_cluster_a[grp];
    if ((uint32_t)cw >= nbin) break;
    const uint32_t ci = bin_idx[(size_t)bin * g.clu_cap + cw];
    const ClusterRec cr = clusters[ci];
    const int sz = (int)cr.count;
    const uint32_t o = cr.offset;
    const uint8_t *im = dec + (size_t)cr.frame * g.Hd * Wp;
    unsigned long long *keys_g = keys + o;
    // warp w takes points [w * 32 E, (w + 1) * 32 E): register r of lane l holds point w * 32 E + r * 32 + l (coalesced loads; the
    // assignment of unsorted points to registers is free)
    const int wbase = wid * 32 * E;
    uint32_t pr[E];
    BBoxRed r = {1 << 30, -1, 1 << 30, -1, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int i = wbase + k * 32 + lane;
      pr[k] = i < sz ? pts[o + i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < E; k++)
      if (wbase + k * 32 + lane < sz) bbox_add(r, pr[k]);
    r = bbox_reduce<THREADS>(r, s_red, lane, wid);
    float cx, cy;
    bool reversed;
    if (!qf_pregate(r, fp, cx, cy, reversed)) {  // (uniform over the cluster's threads)
      if (tid == 0) qinfo[ci] = 0u;
      continue;
    }
    if (MODE == 1) {
      unsigned long long v[E];
#pragma unroll
      for (int k = 0; k < E; k++) v[k] = (k * 32 + lane < sz) ? slope_key(pr[k], cx, cy) : ~0ull;
      warp_bitonic_sort<E>(v, lane);
      // element lane * E + k of the sorted sequence is now v[k]: sorted points out, straight from the registers; the slope half of
      // a key is dead, it now carries the squared gradient magnitude of the decimated image at the point (compute_lfps' weight is
      // sqrt of it, + 1): E gathers in flight per lane
      int g2[E];
#pragma unroll
      for (int k = 0; k < E; k++) g2[k] = (lane * E + k < sz) ? grad2_at(im, Wp, g.Wd, g.Hd, v[k]) : 0;
#pragma unroll
      for (int k = 0; k < E; k++)
        if (lane * E + k < sz) keys_g[lane * E + k] = (v[k] & 0xffffffffull) | ((unsigned long long)(uint32_t)g2[k] << 32);
    } else if (MODE == 2) {
      int NB = 32;
      while (NB * 2 < sz) NB <<= 1;  // power of two, n / 2 <= NB <= NCAP / 2 (NCAP >= 64); measured: NB >= n is slower (4.01 vs 3.87 ms)
      const float nbq = (float)(NB / 4);
      for (int i = tid; i < NB; i += THREADS) cnt[i] = 0u;
      unsigned long long v[E];
#pragma unroll
      for (int k = 0; k < E; k++) v[k] = slope_key(pr[k], cx, cy);
      cta_sync<THREADS>();
      // histogram: the atomic's return value is the point's arrival rank inside its bucket
      uint32_t ba[E];
#pragma unroll
      for (int k = 0; k < E; k++) {
        ba[k] = 0u;
        if (wbase + k * 32 + lane < sz) {
          const uint32_t b = (uint32_t)key_bucket(v[k], nbq, NB);
          ba[k] = b | (atomicAdd(&cnt[b], 1u) << 16);
        }
      }
      cta_sync<THREADS>();
      // exclusive scan of the counters in place (thread t owns cpt consecutive counters) + the largest bucket
      const int cpt = NB >= THREADS ? NB / THREADS : 1;
      const int c0 = tid * cpt;
      uint32_t sum = 0, mx = 0;
      if (c0 < NB)
        for (int j = 0; j < cpt; j++) {
          const uint32_t c = cnt[c0 + j];
          sum += c;
          mx = max(mx, c);
        }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
#pragma unroll
      for (int of = 16; of > 0; of >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, of));
      uint32_t excl = incl - sum;
      if (NW > 1) {
        if (lane == 31) s_ws[grp][wid] = incl;
        if (lane == 0) s_wm[grp][wid] = mx;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < NW; w++) {
          excl += w < wid ? s_ws[grp][w] : 0u;
          mx = max(mx, s_wm[grp][w]);
        }
      }
      if (mx > (uint32_t)bucket_limit) {  // (uniform) degenerate outline: merge sort in global memory, scratch = the cluster's slice of errs
#pragma unroll
        for (int k = 0; k < E; k++)
          if (wbase + k * 32 + lane < sz) keys_g[wbase + k * 32 + lane] = v[k];
        cta_sync<THREADS>();
        sort_keys<THREADS, ITEMS, false>(keys_g, reinterpret_cast<unsigned long long *>(errs_pool + (size_t)2 * o), sz, tid);
        for (int i = tid; i < sz; i += THREADS) {
          const unsigned long long k = keys_g[i];
          keys_g[i] = (k & 0xffffffffull) | ((unsigned long long)(uint32_t)grad2_at(im, Wp, g.Wd, g.Hd, k) << 32);
        }
      } else {
        if (c0 < NB)
          for (int j = 0; j < cpt; j++) {
            const uint32_t c = cnt[c0 + j];
            cnt[c0 + j] = excl;
            excl += c;
          }
        if (tid == 0) cnt[NB] = (uint32_t)sz;
        cta_sync<THREADS>();
#pragma unroll
        for (int k = 0; k < E; k++)
          if (wbase + k * 32 + lane < sz) skeys[cnt[ba[k] & 0xffffu] + (ba[k] >> 16)] = v[k];
        cta_sync<THREADS>();
        // final position = bucket start + number of smaller keys in the bucket; sorted points out with their squared gradient
        for (int i = tid; i < sz; i += THREADS) {
          const unsigned long long key = skeys[i];
          const int g2 = grad2_at(im, Wp, g.Wd, g.Hd, key);  // (the four gathers are in flight during the rank loop)
          const int b = key_bucket(key, nbq, NB);
          const int s = (int)cnt[b], e = (int)cnt[b + 1];
          int rank = 0;
          for (int j = s; j < e; j++) rank += skeys[j] < key ? 1 : 0;
          keys_g[s + rank] = (key & 0xffffffffull) | ((unsigned long long)(uint32_t)g2 << 32);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < E; k++) {
        const int i = wbase + k * 32 + lane;
        if (i < sz) skeys[i] = slope_key(pr[k], cx, cy);
      }
      cta_sync<THREADS>();
      sort_keys<THREADS, ITEMS, false>(skeys, stmp, sz, tid);
      for (int i = tid; i < sz; i += 4 * THREADS) {
        unsigned long long k[4];
        int g2[4];
#pragma unroll
        for (int u = 0; u < 4; u++) k[u] = (i + u * THREADS < sz) ? skeys[i + u * THREADS] : 0ull;
#pragma unroll
        for (int u = 0; u < 4; u++) g2[u] = (i + u * THREADS < sz) ? grad2_at(im, Wp, g.Wd, g.Hd, k[u]) : 0;
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (i + u * THREADS < sz) keys_g[i + u * THREADS] = (k[u] & 0xffffffffull) | ((unsigned long long)(uint32_t)g2[u] << 32);
      }
    }
    if (tid == 0) qf_register_work(ci, o, sz, reversed, qinfo, qwbase, work, work_cap, counters);
  }
}

// Clusters of more than 8192 points (4K-class frames only): keys and merge scratch in global memory.
__global__ void __launch_bounds__(256, 2)
    k_qf_sort_global(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ bin_idx, int bin,
                     const uint32_t *__restrict__ pts, unsigned long long *__restrict__ keys, double *__restrict__ errs_pool,
                     const uint8_t *__restrict__ dec, uint32_t *__restrict__ qinfo, uint32_t *__restrict__ qwbase, uint4 *__restrict__ work,
                     uint32_t work_cap, uint32_t *__restrict__ counters, int Wp) {
  constexpr int THREADS = 256;
  __shared__ BBoxRed s_red[THREADS / 32];
  __shared__ int s_cluster;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t nbin = min(counters[CNT_BIN0 + bin], g.clu_cap);
  for (;;) {
    __syncthreads();
    if (tid == 0) s_cluster = (int)atomicAdd(&counters[CNT_WORK0 + bin], 1u);
    __syncthreads();
    const int cw = s_cluster;
    if ((uint32_t)cw >= nbin) break;
    const uint32_t ci = bin_idx[(size_t)bin * g.clu_cap + cw];
    const ClusterRec cr = clusters[ci];
    const int sz = (int)cr.count;
    const uint32_t o = cr.offset;
    const uint8_t *im = dec + (size_t)cr.frame * g.Hd * Wp;
    unsigned long long *keys_g = keys + o;
    BBoxRed r = {1 << 30, -1, 1 << 30, -1, 0, 0, 0};
    for (int i = tid; i < sz; i += THREADS) bbox_add(r, pts[o + i]);
    r = bbox_reduce<THREADS>(r, s_red, lane, wid);
    float cx, cy;
    bool reversed;
    if (!qf_pregate(r, fp, cx, cy, reversed)) {
      if (tid == 0) qinfo[ci] = 0u;
      continue;
    }
    for (int i = tid; i < sz; i += THREADS) keys_g[i] = slope_key(pts[o + i], cx, cy);
    __syncthreads();
    sort_keys<THREADS, 16, false>(keys_g, reinterpret_cast<unsigned long long *>(errs_pool + (size_t)2 * o), sz, tid);
    for (int i = tid; i < sz; i += THREADS) {
      const unsigned long long k = keys_g[i];
      keys_g[i] = (k & 0xffffffffull) | ((unsigned long long)(uint32_t)grad2_at(im, Wp, g.Wd, g.Hd, k) << 32);
    }
    if (tid == 0) qf_register_work(ci, o, sz, reversed, qinfo, qwbase, work, work_cap, counters);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_qf_window
// ---------------------------------------------------------------------------------------------------------------------
constexpr int QW_SLOTS = 32 * kQfSlotsPerLane;

template <int QW_WARPS>
__global__ void __launch_bounds__(32 * QW_WARPS)
    k_qf_window(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ qinfo,
                const uint4 *__restrict__ work, uint32_t work_cap, const unsigned long long *__restrict__ keys,
                LineFitPt *__restrict__ lfps_pool, double *__restrict__ wtot, uint32_t *__restrict__ wnmax,
                uint32_t *__restrict__ counters) {
  constexpr int PPL = kQfSlotsPerLane, SLOTS = QW_SLOTS;
  extern __shared__ double dsm_win[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *P = dsm_win + (size_t)wid * (8 * SLOTS);  // [6][SLOTS] prefix moments (SoA)
  double *E = P + 6 * SLOTS;                        // [SLOTS] staged keys, then window errors
  double *Y = E + SLOTS;                            // [SLOTS] smoothed errors
  const uint32_t nwork = min(counters[CNT_QWORK], work_cap);
  const double f0 = (double)fp.smooth[0], f1 = (double)fp.smooth[1], f2 = (double)fp.smooth[2], f3 = (double)fp.smooth[3],
               f4 = (double)fp.smooth[4], f5 = (double)fp.smooth[5], f6 = (double)fp.smooth[6];
  // One work item ahead: while an item is being processed, the next one's queue ticket, metadata and keys are already in
  // flight (the dependent chain atomic -> work[] -> cluster record -> keys is ~4 memory round trips; without the prefetch it
  // was 35 % of the kernel's stall samples).
  struct Item {
    uint32_t w, off;
    int n, c, s, len, ksz;
    bool valid;
  };
  unsigned long long kq[PPL];
  auto fetch = [&](Item &it) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&counters[CNT_Q2], 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    it.w = w;
    it.valid = w < nwork;
    if (!it.valid) return;
    const uint4 wk = work[w];  // self-contained record: (point offset, point count, chunk, cluster)
    it.off = wk.x;
    it.n = (int)wk.y;
    it.c = (int)wk.z;
    int e;
    qf_chunk_bounds(it.n, qf_nchunks(it.n), it.c, it.s, e);
    it.len = e - it.s;
    it.ksz = min(20, it.n / 12);
    const int HLn = it.ksz + 5, Ln = it.len + 2 * it.ksz + 9;
    const unsigned long long *kg = keys + it.off;
#pragma unroll
    for (int q = 0; q < PPL; q++) {
      const int j = lane + 32 * q;
      int gi = it.s - HLn + j;
      gi = gi < 0 ? gi + it.n : gi;
      gi = gi >= it.n ? gi - it.n : gi;
      kq[q] = j < Ln ? kg[gi] : 0ull;
    }
  };
  Item cur;
  fetch(cur);
  while (cur.valid) {
    const uint32_t w = cur.w, off = cur.off;
    const int n = cur.n, c = cur.c, s = cur.s, len = cur.len, ksz = cur.ksz;
    (void)n;
    const int HL = ksz + 5;            // halo: ksz + 5 points before the chunk, ksz + 4 after (circular)
    const int L = len + 2 * ksz + 9;   // <= SLOTS by the choice of kQfChunkMax
    unsigned long long *Ek = reinterpret_cast<unsigned long long *>(E);
#pragma unroll
    for (int q = 0; q < PPL; q++) {
      const int j = lane + 32 * q;
      if (j < L) Ek[j] = kq[q];
    }
    __syncwarp();
    Item nxt;
    fetch(nxt);
    // line-fit terms and prefix moments: a lane owns PPL consecutive slots (serial chain), one warp scan joins the lanes
    double acc[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < PPL; q++) {
      const int sl = lane * PPL + q;
      if (sl < L) {
        const unsigned long long k = Ek[sl];
        double t[6];
        lfp_terms(k, (int)(k >> 32), t);
#pragma unroll
        for (int m = 0; m < 6; m++) {
          acc[m] += t[m];
          P[m * SLOTS + sl] = acc[m];
        }
      }
    }
    double lo[6];
#pragma unroll
    for (int m = 0; m < 6; m++) {
      double v = acc[m];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += u;
      }
      lo[m] = __shfl_up_sync(0xffffffffu, v, 1);  // sum of the lanes before mine
    }
    if (lane > 0) {
#pragma unroll
      for (int q = 0; q < PPL; q++) {
        const int sl = lane * PPL + q;
        if (sl < L) {
#pragma unroll
          for (int m = 0; m < 6; m++) P[m * SLOTS + sl] += lo[m];
        }
      }
    }
    __syncwarp();
    // window error of slot j: line fit over slots [j - ksz, j + ksz] (fit_line(i - ksz, i + ksz): N = 2 ksz + 1 points)
    const double Nd = (double)(2 * ksz + 1);
    for (int j = ksz + 1 + lane; j < ksz + len + 9; j += 32) {
      const int a = j - ksz - 1, b = j + ksz;
      const double Mx = P[b] - P[a];
      const double My = P[SLOTS + b] - P[SLOTS + a];
      const double Mxx = P[2 * SLOTS + b] - P[2 * SLOTS + a];
      const double Mxy = P[3 * SLOTS + b] - P[3 * SLOTS + a];
      const double Myy = P[4 * SLOTS + b] - P[4 * SLOTS + a];
      const double W = P[5 * SLOTS + b] - P[5 * SLOTS + a];
      const double rw = __drcp_rn(W);
      const double Ex = Mx * rw, Ey = My * rw;
      const double Cxx = Mxx * rw - Ex * Ex;
      const double Cxy = Mxy * rw - Ex * Ey;
      const double Cyy = Myy * rw - Ey * Ey;
      const double disc = (double)sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
      E[j] = Nd * (0.5 * (Cxx + Cyy - disc));
    }
    __syncwarp();
    // 7-tap smoothing (the oracle's accumulation order)
    for (int j = HL - 1 + lane; j < HL + len + 1; j += 32) {
      double y = 0;
      y += E[j - 3] * f0;
      y += E[j - 2] * f1;
      y += E[j - 1] * f2;
      y += E[j] * f3;
      y += E[j + 1] * f4;
      y += E[j + 2] * f5;
      y += E[j + 3] * f6;
      Y[j] = y;
    }
    __syncwarp();
    // local maxima of the chunk's own points, in index order, into the chunk's slice of the cluster's maxima region
    MaxRec *mr = reinterpret_cast<MaxRec *>(lfps_pool + off) + qf_chunk_rec0(s, c);
    int run = 0;
    for (int j0 = HL; j0 < HL + len; j0 += 32) {
      const int j = j0 + lane;
      bool is = false;
      double y = 0;
      if (j < HL + len) {
        y = Y[j];
        is = y > Y[j + 1] && y > Y[j - 1];
      }
      const unsigned bal = __ballot_sync(0xffffffffu, is);
      if (is) {
        MaxRec rec;
        rec.idx = (uint32_t)(s + j - HL);
        rec.pad = 0;
        rec.y = y;
#pragma unroll
        for (int m = 0; m < 6; m++) rec.P[m] = P[m * SLOTS + j] - P[m * SLOTS + HL - 1];
        mr[run + __popc(bal & ((1u << lane) - 1u))] = rec;
      }
      run += __popc(bal);
    }
    if (lane < 6) wtot[(size_t)w * 6 + lane] = P[lane * SLOTS + HL + len - 1] - P[lane * SLOTS + HL - 1];
    if (lane == 0) wnmax[w] = (uint32_t)run;
    __syncwarp();
    cur = nxt;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_qf_tail
// ---------------------------------------------------------------------------------------------------------------------
constexpr int QT_WARPS = 4;
constexpr int QT_LIN = 256;  // maxima of a cluster held in shared memory for the top-k selection (more: the global scratch)

__global__ void __launch_bounds__(32 * QT_WARPS, 8)
    k_qf_tail(Geo g, FitParams fp, const ClusterRec *__restrict__ clusters, const uint32_t *__restrict__ qinfo,
              const uint32_t *__restrict__ qwbase, const unsigned long long *__restrict__ keys, const LineFitPt *__restrict__ lfps_pool,
              double *__restrict__ errs_pool, const double *__restrict__ wtot, const uint32_t *__restrict__ wnmax,
              QuadRec *__restrict__ quads, uint32_t *__restrict__ counters, ComboTable combos) {
  constexpr int TBL = MAXM * MAXM;
  __shared__ double s_G_a[QT_WARPS][MAXM][6];    // global inclusive prefix moments at the kept maxima
  __shared__ double s_H_a[QT_WARPS][MAXM][6];    // ... at the point before each kept maximum (0 for point 0)
  __shared__ double s_T_a[QT_WARPS][6];          // moments of the whole cluster
  __shared__ double s_tab_a[QT_WARPS][3 * TBL];  // pair tables: mse, nx, ny
  __shared__ double s_val_a[QT_WARPS][QT_LIN];
  __shared__ int s_fm_a[QT_WARPS][MAXM];
  __shared__ int s_kp_a[QT_WARPS][MAXM];         // ordinal (position in index order) of each kept maximum
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double(*s_G)[6] = s_G_a[wid];
  double(*s_H)[6] = s_H_a[wid];
  double *s_T = s_T_a[wid];
  double *pt_mse = s_tab_a[wid], *pt_nx = pt_mse + TBL, *pt_ny = pt_nx + TBL;
  double *s_val = s_val_a[wid];
  int *s_fm = s_fm_a[wid], *s_kp = s_kp_a[wid];
  const uint32_t ncl = min(counters[CNT_CLUSTERS], g.clu_cap);
  // Dynamic queue (cluster costs differ by orders of magnitude), software-pipelined: the ticket of the cluster after the next is
  // in flight while the next cluster's metadata loads and the current one is processed (the chain ticket -> qinfo -> record was
  // exposed once per cluster: long-scoreboard stalls were 31 % of the kernel's samples).
  auto take_raw = [&]() -> uint32_t { return lane == 0 ? atomicAdd(&counters[CNT_Q3], 1u) : 0u; };
  uint32_t ci_next = __shfl_sync(0xffffffffu, take_raw(), 0);
  uint32_t info_n = 0u, wb_n = 0u;
  ClusterRec cr_n = {0ull, 0u, 0u, 0u, 0u};
  if (ci_next < ncl) {
    info_n = qinfo[ci_next];
    cr_n = clusters[ci_next];
    wb_n = qwbase[ci_next];
  }
  uint32_t t_raw = take_raw();
  for (;;) {
    __syncwarp();
    const uint32_t ci = ci_next;
    if (ci >= ncl) break;
    const uint32_t info = info_n;
    const ClusterRec cr = cr_n;
    const uint32_t wb = wb_n;
    ci_next = __shfl_sync(0xffffffffu, t_raw, 0);  // (issued one cluster ago)
    if (ci_next < ncl) {
      info_n = qinfo[ci_next];
      cr_n = clusters[ci_next];
      wb_n = qwbase[ci_next];
    }
    t_raw = take_raw();
    if (info == 0u) continue;
    const int sz = (int)(info & 0x7fffffffu);
    const bool reversed = (info >> 31) != 0u;
    const int nch = qf_nchunks(sz);
    const MaxRec *mbase = reinterpret_cast<const MaxRec *>(lfps_pool + cr.offset);
    int m = 0;
    for (int c = 0; c < nch; c++) m += (int)wnmax[wb + c];
    if (m < 4) continue;
    // record of the maximum with ordinal p (index order: chunk by chunk)
    auto rec_of = [&](int p, int &chunk) -> const MaxRec * {
      int c = 0, base = 0;
      for (;;) {
        const int cnt = (int)wnmax[wb + c];
        if (p < base + cnt || c == nch - 1) break;
        base += cnt;
        c++;
      }
      int s, e;
      qf_chunk_bounds(sz, nch, c, s, e);
      chunk = c;
      return mbase + qf_chunk_rec0(s, c) + (p - base);
    };
    // ---- keep the max_nmaxima best: threshold = value of descending rank max_nmaxima, keep err > threshold ----
    int nm = 0;
    if (m <= 32) {
      int chunk = 0;
      double v = -CUDART_INF;
      if (lane < m) v = rec_of(lane, chunk)->y;
      bool keep = lane < m;
      if (m > fp.max_nmaxima) {
        int rank = 0;
        for (int j = 0; j < m; j++) {
          const double u = __shfl_sync(0xffffffffu, v, j);
          rank += (u > v || (u == v && j < lane)) ? 1 : 0;
        }
        const unsigned tb = __ballot_sync(0xffffffffu, lane < m && rank == fp.max_nmaxima);
        const double thresh = __shfl_sync(0xffffffffu, v, __ffs(tb) - 1);
        keep = keep && !(v <= thresh);
      }
      const unsigned kb = __ballot_sync(0xffffffffu, keep);
      nm = min(__popc(kb), MAXM);
      if (keep) {
        const int pos = __popc(kb & ((1u << lane) - 1u));
        if (pos < MAXM) s_kp[pos] = lane;
      }
    } else {
      // values in index order into a linear array (shared memory, or the cluster's slice of the global scratch)
      double *val = m <= QT_LIN ? s_val : errs_pool + (size_t)2 * cr.offset;
      {
        int base = 0;
        for (int c = 0; c < nch; c++) {
          const int cnt = (int)wnmax[wb + c];
          int s, e;
          qf_chunk_bounds(sz, nch, c, s, e);
          const MaxRec *rc = mbase + qf_chunk_rec0(s, c);
          for (int p = lane; p < cnt; p += 32) val[base + p] = rc[p].y;
          base += cnt;
        }
      }
      __syncwarp();
      // element of rank r in the order (value descending, position ascending), r = 0 .. max_nmaxima: one pass per rank
      double pv = CUDART_INF;
      int pp = -1;
      for (int round = 0; round <= fp.max_nmaxima; round++) {
        double bv = -CUDART_INF;
        int bi = 0x7fffffff;
        for (int i = lane; i < m; i += 32) {
          const double v = val[i];
          const bool after = v < pv || (v == pv && i > pp);  // not yet taken
          if (after && (v > bv || (v == bv && i < bi))) {
            bv = v;
            bi = i;
          }
        }
        for (int of = 16; of > 0; of >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, of);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, of);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        pv = bv;
        pp = bi;
      }
      const double thresh = pv;  // value of rank max_nmaxima (m > 32 > max_nmaxima here)
      int run = 0;
      for (int i0 = 0; i0 < m; i0 += 32) {
        const int i = i0 + lane;
        const bool keep = i < m && !(val[i] <= thresh);
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int pos = run + __popc(kb & ((1u << lane) - 1u));
          if (pos < MAXM) s_kp[pos] = i;
        }
        run += __popc(kb);
      }
      nm = min(run, MAXM);
    }
    __syncwarp();
    if (nm < 4) continue;
    // ---- global prefix moments at the kept maxima: totals of the chunks before + chunk-local prefix ----
    if (lane < nm) {
      int chunk = 0;
      const MaxRec *rc = rec_of(s_kp[lane], chunk);
      const uint32_t idx = rc->idx;
      double G[6];
#pragma unroll
      for (int q = 0; q < 6; q++) G[q] = 0;
      for (int c = 0; c < chunk; c++) {
#pragma unroll
        for (int q = 0; q < 6; q++) G[q] += wtot[(size_t)(wb + c) * 6 + q];
      }
      const unsigned long long k = keys[cr.offset + idx];
      double t[6];
      lfp_terms(k, (int)(k >> 32), t);
#pragma unroll
      for (int q = 0; q < 6; q++) {
        G[q] += rc->P[q];
        s_G[lane][q] = G[q];
        s_H[lane][q] = idx == 0 ? 0.0 : G[q] - t[q];
      }
      s_fm[lane] = (int)idx;
    }
    if (lane < 6) {
      double T = 0;
      for (int c = 0; c < nch; c++) T += wtot[(size_t)(wb + c) * 6 + lane];
      s_T[lane] = T;
    }
    __syncwarp();
    // fit_line between kept maxima a -> b (oracle fit_line: prefix difference, wrapping when i0 > i1).  FAST (the pair table,
    // which only SELECTS the four corners): reciprocal-multiply instead of the seven divisions; the four final lines, which
    // give the corners, use the oracle's divisions.
    auto fit_pair = [&](int a, int b, double *lineparm, double *mse, bool fast) {
      const int i0 = s_fm[a], i1 = s_fm[b];
      double M[6];
      int N;
      if (i0 < i1) {
        N = i1 - i0 + 1;
#pragma unroll
        for (int q = 0; q < 6; q++) M[q] = s_G[b][q];
        if (i0 > 0) {
#pragma unroll
          for (int q = 0; q < 6; q++) M[q] -= s_H[a][q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 6; q++) {
          M[q] = s_T[q] - s_H[a][q];
          M[q] += s_G[b][q];
        }
        N = sz - i0 + i1 + 1;
      }
      if (!fast) {
        fit_moments_dev(M[0], M[1], M[2], M[3], M[4], M[5], N, lineparm, nullptr, mse);
        return;
      }
      const double rw = __drcp_rn(M[5]);
      const double Ex = M[0] * rw, Ey = M[1] * rw;
      const double Cxx = M[2] * rw - Ex * Ex, Cxy = M[3] * rw - Ex * Ey, Cyy = M[4] * rw - Ey * Ey;
      const double disc = (double)sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy));
      *mse = 0.5 * (Cxx + Cyy - disc);
      const double eig = 0.5 * (Cxx + Cyy + disc);
      const double nx1 = Cxx - eig, ny1 = Cxy, M1 = nx1 * nx1 + ny1 * ny1;
      const double nx2 = Cxy, ny2 = Cyy - eig, M2 = nx2 * nx2 + ny2 * ny2;
      const bool first = M1 > M2;
      const double nx = first ? nx1 : nx2, ny = first ? ny1 : ny2;
      const double length = (double)sqrtf((float)(first ? M1 : M2));
      const double rl = fabs(length) < 1e-12 ? 0.0 : __drcp_rn(length);
      lineparm[2] = nx * rl;
      lineparm[3] = ny * rl;
    };
    // ---- pair table ----
    for (int t = lane; t < nm * nm; t += 32) {
      const int a = t / nm, b = t - a * nm;
      if (a == b) continue;
      double lp[4], ms;
      fit_pair(a, b, lp, &ms, true);
      pt_mse[a * MAXM + b] = ms;
      pt_nx[a * MAXM + b] = lp[2];
      pt_ny[a * MAXM + b] = lp[3];
    }
    __syncwarp();
    // ---- best (m0<m1<m2<m3); ties resolved to the lexicographically first, like the serial loops ----
    double best = CUDART_INF;
    uint32_t brank = 0xffffffffu;
    {
      const double max_mse = (double)fp.max_line_fit_mse, max_dot = (double)fp.cos_critical_rad;
      const uchar4 *ctab = combos.c + combos.off[nm];
      const int ncomb = combos.off[nm + 1] - combos.off[nm];
      auto seg_err = [&](int a, int b, double mse) {
        const int i0 = s_fm[a], i1 = s_fm[b];
        const int N = i0 < i1 ? i1 - i0 + 1 : sz - i0 + i1 + 1;
        return N * mse;
      };
      for (int t = lane; t < ncomb; t += 32) {
        const uchar4 c = ctab[t];
        const int m0 = c.x, m1 = c.y, m2 = c.z, m3 = c.w;
        const double mse01 = pt_mse[m0 * MAXM + m1];
        if (mse01 > max_mse) continue;
        const double mse12 = pt_mse[m1 * MAXM + m2];
        if (mse12 > max_mse) continue;
        const double dot = pt_nx[m0 * MAXM + m1] * pt_nx[m1 * MAXM + m2] + pt_ny[m0 * MAXM + m1] * pt_ny[m1 * MAXM + m2];
        if (fabs(dot) > max_dot) continue;
        const double mse23 = pt_mse[m2 * MAXM + m3];
        if (mse23 > max_mse) continue;
        const double mse30 = pt_mse[m3 * MAXM + m0];
        if (mse30 > max_mse) continue;
        const double err = seg_err(m0, m1, mse01) + seg_err(m1, m2, mse12) + seg_err(m2, m3, mse23) + seg_err(m3, m0, mse30);
        if (err < best || (err == best && (uint32_t)t < brank)) {
          best = err;
          brank = (uint32_t)t;
        }
      }
    }
    for (int of = 16; of > 0; of >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, of);
      const uint32_t orank = __shfl_xor_sync(0xffffffffu, brank, of);
      if (ov < best || (ov == best && orank < brank)) {
        best = ov;
        brank = orank;
      }
    }
    // ---- final lines, corners, area / angle gates: lanes 0..3, one line / corner / angle each ----
    bool ok = brank != 0xffffffffu;
    if (ok && !(best / sz < (double)fp.max_line_fit_mse)) ok = false;
    if (!ok) continue;  // warp-uniform
    {
      const uchar4 c = combos.c[combos.off[nm] + brank];
      const int kept[4] = {c.x, c.y, c.z, c.w};
      const int li = lane & 3;
      double ln[4], mse;
      fit_pair(kept[li], kept[(li + 1) & 3], ln, &mse, false);
      ok = __all_sync(0xffffffffu, !(mse > (double)fp.max_line_fit_mse));
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
          const double ddx = (double)(qp[ib[k]][0] - qp[ia[k]][0]), ddy = (double)(qp[ib[k]][1] - qp[ia[k]][1]);
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
      {
        const double ccr = (double)fp.cos_critical_rad;
        const int i0 = li, i1 = (li + 1) & 3, i2 = (li + 2) & 3;
        const double dx1 = (double)(qp[i1][0] - qp[i0][0]);
        const double dy1 = (double)(qp[i1][1] - qp[i0][1]);
        const double dx2 = (double)(qp[i2][0] - qp[i1][0]);
        const double dy2 = (double)(qp[i2][1] - qp[i1][1]);
        const double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
        const bool bad = (cos_dtheta > ccr || cos_dtheta < -ccr) || dx1 * dy2 < dy1 * dx2;
        ok = ok && __all_sync(0xffffffffu, !bad);
      }
      if (ok && lane == 0) {
        const uint32_t qi = atomicAdd(&counters[CNT_QUADS], 1u);
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

// ---------------------------------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------------------------------
static int device_index() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < 64 ? dev : 0;
}

template <int THREADS, int E, int ITEMS, int MINB, int WPC, int MODE>
static void launch_sort_bin(const Workspace &ws, int bin, int sms, cudaStream_t st) {
  const Geo &g = ws.g;
  constexpr size_t smem = qf_sort_smem<THREADS, E, MODE, WPC>();
  auto kern = k_qf_sort<THREADS, E, ITEMS, MINB, WPC, MODE>;
  static int ctas_per_sm[64] = {};
  const int dev = device_index();
  if (!ctas_per_sm[dev]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, THREADS * WPC, smem);
    ctas_per_sm[dev] = std::max(1, n);
  }
  kern<<<sms * ctas_per_sm[dev], THREADS * WPC, smem, st>>>(g, ws.fp, ws.clusters, ws.bin_idx, bin, ws.pts, ws.keys, ws.errs, ws.dec, ws.qinfo,
                                                           ws.qwbase, ws.qwork, ws.qwork_cap, ws.counters, at_Wp(g), ws.tune.qf_bucket_limit);
}

// warps per CTA of k_qf_window: 14.3 KB of shared memory per warp.  Measured: 4 per CTA (3 CTAs = 12 warps per SM) 1.585 ms,
// 3 per CTA (5 CTAs = 15 warps per SM) 1.629 ms -- the kernel is not occupancy bound.
template <int QW_WARPS>
static void launch_window(const Workspace &ws, int sms, cudaStream_t s) {
  constexpr size_t smem = (size_t)QW_WARPS * 8 * QW_SLOTS * sizeof(double);
  static int ctas_per_sm[64] = {};
  const int di = device_index();
  if (!ctas_per_sm[di]) {
    cudaFuncSetAttribute(k_qf_window<QW_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_qf_window<QW_WARPS>, 32 * QW_WARPS, smem);
    ctas_per_sm[di] = std::max(1, n);
  }
  k_qf_window<QW_WARPS><<<sms * ctas_per_sm[di], 32 * QW_WARPS, smem, s>>>(ws.g, ws.fp, ws.clusters, ws.qinfo, ws.qwork, ws.qwork_cap, ws.keys, ws.lfps,
                                                                            ws.qwtot, ws.qwnmax, ws.counters);
}

int launch_quadfit_windowed(const Workspace &ws, int nframes, cudaStream_t s) {
  (void)nframes;
  const Geo &g = ws.g;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  ComboTable ct;
  ct.c = reinterpret_cast<const uchar4 *>(ws.combos);
  for (int i = 0; i < 18; i++) ct.off[i] = ws.combo_off[i];
  launch_bin_clusters(ws, sms, s);
  // the size bins are independent: fork onto side streams, large clusters (the long poles) first
  cudaEventRecord(ws.ev_fork, s);
  for (int i = 0; i < kQuadAux; i++) cudaStreamWaitEvent(ws.aux[i], ws.ev_fork, 0);
  k_qf_sort_global<<<sms * 2, 256, 0, s>>>(g, ws.fp, ws.clusters, ws.bin_idx, 7, ws.pts, ws.keys, ws.errs, ws.dec, ws.qinfo, ws.qwbase, ws.qwork,
                                          ws.qwork_cap, ws.counters, at_Wp(g));                 // n > 8192 (4K-class frames)
  launch_sort_bin<512, 16, 16, 2, 1, 2>(ws, 6, sms, ws.aux[0]);  // n <= 8192
  launch_sort_bin<256, 16, 16, 4, 1, 2>(ws, 5, sms, ws.aux[1]);  // n <= 4096
  launch_sort_bin<256, 8, 8, 4, 1, 2>(ws, 4, sms, ws.aux[2]);    // n <= 2048
  launch_sort_bin<128, 8, 8, 8, 1, 2>(ws, 3, sms, ws.aux[3]);    // n <= 1024
  launch_sort_bin<64, 8, 8, 16, 1, 2>(ws, 2, sms, ws.aux[4]);    // n <= 512
  launch_sort_bin<32, 8, 8, 4, 8, 2>(ws, 1, sms, ws.aux[5]);     // n <= 256: one warp per cluster, 8 workers per CTA
  launch_sort_bin<32, 4, 4, 4, 8, 1>(ws, 0, sms, ws.aux[6]);     // n <= 128: register bitonic network (measured: 0.218 vs 0.226 ms)
  for (int i = 0; i < kQuadAux; i++) {
    cudaEventRecord(ws.ev_join[i], ws.aux[i]);
    cudaStreamWaitEvent(s, ws.ev_join[i], 0);
  }
  launch_window<4>(ws, sms, s);
  {
    static int ctas_per_sm[64] = {};
    const int di = device_index();
    if (!ctas_per_sm[di]) {
      int n = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_qf_tail, 32 * QT_WARPS, 0);
      ctas_per_sm[di] = std::max(1, n);
    }
    k_qf_tail<<<sms * ctas_per_sm[di], 32 * QT_WARPS, 0, s>>>(g, ws.fp, ws.clusters, ws.qinfo, ws.qwbase, ws.keys, ws.lfps, ws.errs, ws.qwtot,
                                                              ws.qwnmax, ws.quads, ws.counters, ct);
  }
  return 3 + kQuadBins;
}

}  // namespace b200at
