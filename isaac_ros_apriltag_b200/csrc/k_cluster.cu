// k_cluster.cu -- gradient_clusters (a9): boundary points between a black and a white component, grouped by
// the component pair.  Restates AprilRobotics do_gradient_clusters (v3.4.x, incl. `connected_last`;
// SURVEY App. A.4).  The CPU hash-map-of-zarrays becomes a count / select / emit triple over a per-frame
// open-addressing table, so only clusters that can pass fit_quad's size gates (24 <= n <= 2(2w+2h)) are ever
// materialised:
//   k_cluster_count  : every boundary point -> (key) -> hash insert + count   (warp-aggregated atomics)
//   k_cluster_select : per frame, scan the table: kept clusters get a contiguous segment in the point pool
//   k_cluster_emit   : every boundary point of a kept cluster -> its segment   (order inside a segment is
//                      irrelevant: the quad-fit sort key (slope, y, x) is a total order)
// Algorithmic bytes per frame: read Pd (thr2) + 4*Pd (labels) per pass, + 4 B per emitted point.
#include "detector.h"

namespace b200at {

// A boundary point exists between pixel (x, y) and its neighbour at (1,0), (0,1), (-1,1), (1,1) iff v0 + v1 == 255 -- thr2 already
// folds both size gates (components < 25 px read as 127).  `connected_last`: the (-1,1) probe is skipped when the previous pixel's
// (1,1) probe produced a point; at x == 1 there is no previous pixel.

// cheap 64->32 bit mix for the (rep_hi, rep_lo) keys
__device__ __forceinline__ uint32_t hash_key2(unsigned long long k) {
  uint32_t h = (uint32_t)(k >> 32) * 0x9E3779B1u ^ (uint32_t)k * 0x85EBCA6Bu;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  return h;
}

// Four pixels per thread (x = 4t .. 4t+3 of row y): the byte image is read as aligned words (three per row instead of seven
// byte loads per pixel), the labels as one uint4 + scalars per row, and the points of a warp (up to 16 per thread) are compacted
// into shared memory with one exclusive scan of the per-thread counts per half.  The order of the points inside a cluster's segment is
// irrelevant (sorted later).
// Count pass: issue bound (ncu: 79 % of peak issue): per chunk of 32 compacted points one __match_any_sync groups equal keys, the
// group leaders insert / find the key and add the group size with a fire-and-forget atomic.
// Emit pass: was latency bound -- per chunk a dependent chain key load -> compare -> atomicAdd with return -> store, 47 % of
// its stall samples on exactly that chain (profiles/r03_ncu_full_irregular_batch32.csv).  Now the chain is software-pipelined
// over kPipe chunks: every lane first loads the table entry of its own key's home slot for all chunks (no leader-only
// divergence in front of the loads), then the leaders issue all cursor atomics, and only then the results are consumed.
// The point word is assembled per pixel (the gradient-sign codes of the four probes are constants of the pixel's polarity).
// (Measured and dropped, profiles/r03_variants.md: deferring the emit pass's stores by one round; recording (slot, point)
// pairs in the count pass and scattering them instead of a second pass; grouping the points of a whole row by key in a
// shared-memory table so that one thread per distinct key touches the global table.)
constexpr int kPipe = 2;
template <bool EMIT>
__global__ void __launch_bounds__(256, EMIT ? 6 : 8) k_cluster_pass4(Geo g, const uint8_t *__restrict__ thr2, const uint32_t *__restrict__ lab,
                                                                     unsigned long long *__restrict__ hkey, uint32_t *__restrict__ hcnt,
                                                                     const uint32_t *__restrict__ hoff, uint32_t *__restrict__ hcur,
                                                                     uint32_t *__restrict__ pts, uint32_t *__restrict__ counters, int Wp) {
  // the four pixels of a thread are processed as two halves of two: 256 compacted points per warp at most, 24 KB of shared
  // memory per CTA (measured: all four at once costs more in occupancy than the second scan costs in instructions)
  __shared__ unsigned long long s_key[8][256];
  __shared__ uint32_t s_pt[EMIT ? 8 : 1][EMIT ? 256 : 1];
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y + 1;  // 1 .. Hd-2
  const int fr = blockIdx.z;
  const size_t fo = (size_t)fr * g.Hd * Wp;
  const uint8_t *r0 = thr2 + fo + (size_t)y * Wp, *r1 = r0 + Wp;
  const uint32_t *l0 = lab + fo + (size_t)y * Wp, *l1 = l0 + Wp;
  unsigned long long *hk = hkey + (size_t)fr * g.hcap;
  const size_t ho = (size_t)fr * g.hcap;
  const uint32_t hmask = g.hcap - 1;
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool any = x4 <= g.Wd - 2;  // at least one of my pixels can be interior
  // 6-pixel windows (columns x4-1 .. x4+4) of the two byte rows and of the label rows
  uint32_t a0 = 0x7f7f7f7fu, a1 = 0x7f7f7f7fu, a2 = 0x7f7f7f7fu, b0 = 0x7f7f7f7fu, b1 = 0x7f7f7f7fu, b2 = 0x7f7f7f7fu;
  uint4 la = make_uint4(0, 0, 0, 0), lb = make_uint4(0, 0, 0, 0);
  uint32_t la4 = 0, lbm = 0, lb4 = 0;
  if (any) {
    a1 = *reinterpret_cast<const uint32_t *>(r0 + x4);
    b1 = *reinterpret_cast<const uint32_t *>(r1 + x4);
    la = *reinterpret_cast<const uint4 *>(l0 + x4);
    lb = *reinterpret_cast<const uint4 *>(l1 + x4);
    if (x4 > 0) {
      a0 = *reinterpret_cast<const uint32_t *>(r0 + x4 - 4);
      b0 = *reinterpret_cast<const uint32_t *>(r1 + x4 - 4);
      lbm = l1[x4 - 1];
    }
    if (x4 + 4 < Wp) {
      a2 = *reinterpret_cast<const uint32_t *>(r0 + x4 + 4);
      b2 = *reinterpret_cast<const uint32_t *>(r1 + x4 + 4);
      la4 = l0[x4 + 4];
      lb4 = l1[x4 + 4];
    }
  }
  int va[6], vb[6];
  va[0] = (int)(a0 >> 24);
  vb[0] = (int)(b0 >> 24);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    va[1 + j] = (int)((a1 >> (8 * j)) & 0xffu);
    vb[1 + j] = (int)((b1 >> (8 * j)) & 0xffu);
  }
  va[5] = (int)(a2 & 0xffu);
  vb[5] = (int)(b2 & 0xffu);
  const uint32_t lrow0[5] = {la.x, la.y, la.z, la.w, la4};            // labels of row y,   columns x4 .. x4+4
  const uint32_t lrow1[6] = {lbm, lb.x, lb.y, lb.z, lb.w, lb4};       // labels of row y+1, columns x4-1 .. x4+4
  uint2 *skey2 = reinterpret_cast<uint2 *>(s_key[wid]);
#pragma unroll
  for (int half = 0; half < 2; half++) {
    // my points of pixels 2 * half, 2 * half + 1: bit (4 * jj + k) of `mask` = probe k of pixel jj produced a point
    unsigned mask = 0;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * half + jj;
      const int x = x4 + j;
      const bool in = any && x >= 1 && x <= g.Wd - 2;
      const int v0 = va[1 + j], vl = va[j], vr = va[2 + j], dl = vb[j], dc = vb[1 + j], dr = vb[2 + j];
      const bool prev_conn = (x > 1) && (vl + dc == 255);
      if (in && (v0 + vr == 255)) mask |= 1u << (4 * jj + 0);
      if (in && (v0 + dc == 255)) mask |= 1u << (4 * jj + 1);
      if (in && !prev_conn && (v0 + dl == 255)) mask |= 1u << (4 * jj + 2);
      if (in && (v0 + dr == 255)) mask |= 1u << (4 * jj + 3);
    }
    const int cnt = __popc(mask);
    // exclusive scan of the per-thread counts
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int pos = incl - cnt;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * half + jj;
      const uint32_t rep0 = lrow0[j];
      const uint32_t rep1[4] = {lrow0[j + 1], lrow1[j + 1], lrow1[j], lrow1[j + 2]};
      // point word (2x+dx) | (2y+dy) << 14 | cx << 28 | cy << 30 with cx / cy = sign code (0: zero, 1: positive, 2: negative) of
      // the gradient dx * (v1 - v0), dy * (v1 - v0); v1 - v0 = +255 when v0 == 0, -255 when v0 == 255 (v0 + v1 == 255)
      const uint32_t A = va[1 + j] == 0 ? 1u : 2u, Bc = 3u - A;
      const uint32_t low = (uint32_t)(2 * (x4 + j)) | ((uint32_t)(2 * y) << 14);
      const uint32_t ptw[4] = {(low + 1u) | (A << 28), (low + (1u << 14)) | (A << 30), (low - 1u + (1u << 14)) | (Bc << 28) | (A << 30),
                               (low + 1u + (1u << 14)) | (A << 28) | (A << 30)};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if ((mask >> (4 * jj + k)) & 1u) {
          const uint32_t q1 = rep1[k];
          skey2[pos] = make_uint2(min(rep0, q1), max(rep0, q1));  // key = larger representative << 32 | smaller
          if (EMIT) s_pt[wid][pos] = ptw[k];
          pos++;
        }
      }
    }
    __syncwarp();
    if (!EMIT) {
      for (int c0 = 0; c0 < total; c0 += 32) {
        const int j = c0 + (int)lane;
        const bool has = j < total;
        const unsigned act = __ballot_sync(0xffffffffu, has);
        if (!has) continue;
        const unsigned long long key = s_key[wid][j];
        const unsigned peers = __match_any_sync(act, key);
        const int leader = __ffs(peers) - 1;
        const int n = __popc(peers);
        uint32_t slot = hash_key2(key) & hmask;
        if ((int)lane == leader) {
          uint32_t found = 0xffffffffu;
          for (uint32_t probe = 0; probe < g.hcap; probe++) {
            unsigned long long cur = hk[slot];
            if (cur == key) {
              found = slot;
              break;
            }
            if (cur == 0ULL) {
              unsigned long long old = atomicCAS(&hk[slot], 0ULL, key);
              if (old == 0ULL || old == key) {
                found = slot;
                break;
              }
            }
            slot = (slot + 1) & hmask;
          }
          if (found == 0xffffffffu)
            atomicOr(&counters[CNT_STATUS], (uint32_t)ST_HASH_FULL);
          else
            atomicAdd(&hcnt[ho + found], (uint32_t)n);
        }
      }
    } else {
      for (int c0 = 0; c0 < total; c0 += 32 * kPipe) {
        unsigned long long key[kPipe], cur[kPipe];
        uint32_t slot[kPipe], off[kPipe], base[kPipe];
        unsigned peers[kPipe];
        // (A) the table entry at the home slot of every lane's own key, all chunks in flight
#pragma unroll
        for (int u = 0; u < kPipe; u++) {
          const int j = c0 + 32 * u + (int)lane;
          key[u] = j < total ? s_key[wid][j] : 0ULL;
          slot[u] = hash_key2(key[u]) & hmask;
          cur[u] = 0ULL;
          off[u] = 0xffffffffu;
          if (j < total) {
            cur[u] = hk[slot[u]];
            off[u] = hoff[ho + slot[u]];
          }
        }
        // (B) collisions probe on (rare); equal keys are grouped and the group leader advances the cluster's cursor
#pragma unroll
        for (int u = 0; u < kPipe; u++) {
          base[u] = 0xffffffffu;
          peers[u] = 0u;
          if (c0 + 32 * u >= total) continue;  // (uniform)
          const bool has = c0 + 32 * u + (int)lane < total;
          const unsigned act = __ballot_sync(0xffffffffu, has);
          if (!has) continue;
          for (uint32_t probe = 0; cur[u] != key[u] && cur[u] != 0ULL && probe < g.hcap; probe++) {
            slot[u] = (slot[u] + 1) & hmask;
            cur[u] = hk[slot[u]];
            off[u] = hoff[ho + slot[u]];
          }
          peers[u] = __match_any_sync(act, key[u]);
          const int leader = __ffs(peers[u]) - 1;
          // (the atomic's return value is not touched before (C): the next chunk's work issues behind it)
          if ((int)lane == leader && cur[u] == key[u] && off[u] != 0xffffffffu)
            base[u] = atomicAdd(&hcur[ho + slot[u]], (uint32_t)__popc(peers[u]));
          else
            off[u] = 0xffffffffu;
        }
        // (C) the points go to their cluster's segment
#pragma unroll
        for (int u = 0; u < kPipe; u++) {
          if (peers[u] == 0u) continue;
          const int leader = __ffs(peers[u]) - 1;
          const uint32_t o = __shfl_sync(peers[u], off[u], leader), bs = __shfl_sync(peers[u], base[u], leader);
          if (o != 0xffffffffu) pts[o + bs + __popc(peers[u] & ((1u << lane) - 1))] = s_pt[wid][c0 + 32 * u + (int)lane];
        }
      }
    }
    __syncwarp();  // the buffer is reused by the second half
  }
}

// One CTA per (frame, table segment of kSelSeg slots).  A thread owns kSelPT consecutive slots and keeps their counts in
// registers: one block scan gives the segment's totals (-> one reservation in the global cluster / point pools) and every
// slot's position inside the reservation (clusters of a segment appear in table-slot order).  The table is read once.
constexpr int kSelPT = 16, kSelSeg = 1024 * kSelPT;
__global__ void __launch_bounds__(1024) k_cluster_select(Geo g, const unsigned long long *__restrict__ hkey,
                                                         const uint32_t *__restrict__ hcnt, uint32_t *__restrict__ hoff,
                                                         uint32_t *__restrict__ hcur, ClusterRec *__restrict__ clusters,
                                                         uint32_t *__restrict__ counters, int segs) {
  const int fr = blockIdx.x / segs, seg = blockIdx.x % segs;
  const uint32_t seg_len = g.hcap / (uint32_t)segs;  // <= kSelSeg; a multiple of 4 (hcap is a power of two >= 4 * segs)
  const size_t ho = (size_t)fr * g.hcap + (size_t)seg * seg_len;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ uint32_t s_c[32], s_p[32];
  __shared__ uint32_t s_base_c, s_base_p, s_ok;
  const uint32_t i0 = (uint32_t)tid * kSelPT;
  uint32_t c[kSelPT];
#pragma unroll
  for (int q = 0; q < kSelPT / 4; q++) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i0 + 4 * q < seg_len) v = *reinterpret_cast<const uint4 *>(hcnt + ho + i0 + 4 * q);
    c[4 * q] = v.x;
    c[4 * q + 1] = v.y;
    c[4 * q + 2] = v.z;
    c[4 * q + 3] = v.w;
  }
  uint32_t nc = 0, np = 0;
#pragma unroll
  for (int k = 0; k < kSelPT; k++) {
    const bool keep = c[k] >= 24u && c[k] <= g.max_cluster_pts;
    nc += keep ? 1u : 0u;
    np += keep ? c[k] : 0u;
  }
  // block scan of (clusters, points) per thread
  uint32_t ic = nc, ip = np;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t tc = __shfl_up_sync(0xffffffffu, ic, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
    if (lane >= o) {
      ic += tc;
      ip += tp;
    }
  }
  if (lane == 31) {
    s_c[wid] = ic;
    s_p[wid] = ip;
  }
  __syncthreads();
  if (wid == 0) {
    const uint32_t vc = s_c[lane], vp = s_p[lane];
    uint32_t jc = vc, jp = vp;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t tc = __shfl_up_sync(0xffffffffu, jc, o), tp = __shfl_up_sync(0xffffffffu, jp, o);
      if (lane >= o) {
        jc += tc;
        jp += tp;
      }
    }
    s_c[lane] = jc - vc;  // exclusive warp offsets
    s_p[lane] = jp - vp;
    if (lane == 31) {  // totals of the segment: one reservation
      const uint32_t tc = jc, tp = jp;
      const uint32_t bc = atomicAdd(&counters[CNT_CLUSTERS], tc);
      const uint32_t bp = atomicAdd(&counters[CNT_POINTS], tp);
      uint32_t ok = 1;
      if (bc + tc > g.clu_cap) {
        atomicOr(&counters[CNT_STATUS], (uint32_t)ST_CLUSTERS_FULL);
        ok = 0;
      }
      if (bp + tp > g.pts_cap) {
        atomicOr(&counters[CNT_STATUS], (uint32_t)ST_POINTS_FULL);
        ok = 0;
      }
      s_base_c = bc;
      s_base_p = bp;
      s_ok = ok;
    }
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  uint32_t ec = s_base_c + s_c[wid] + ic - nc;  // position of my first kept slot inside the pools
  uint32_t ep = s_base_p + s_p[wid] + ip - np;
  uint32_t off[kSelPT];
#pragma unroll
  for (int k = 0; k < kSelPT; k++) {
    const bool would_keep = c[k] >= 24u && c[k] <= g.max_cluster_pts;
    off[k] = 0xffffffffu;
    if (would_keep) {
      const uint32_t i = i0 + k;
      if (ok) {
        ClusterRec r;
        r.key = hkey[ho + i];
        r.offset = ep;
        r.count = c[k];
        r.frame = (uint32_t)fr;
        r.pad = 0;
        clusters[ec] = r;
        off[k] = ep;
      } else if (ec < g.clu_cap) {
        // A reservation that did not fit keeps its range of the cluster list (the counter stays advanced): its records are
        // written EMPTY (count 0), so that a consumer that walks [0, min(CNT_CLUSTERS, clu_cap)) never reads a stale record.
        clusters[ec] = ClusterRec{hkey[ho + i], 0u, 0u, (uint32_t)fr, 0u};
      }
      ec++;
      ep += c[k];
    }
  }
#pragma unroll
  for (int q = 0; q < kSelPT / 4; q++) {
    if (i0 + 4 * q < seg_len) {
      *reinterpret_cast<uint4 *>(hoff + ho + i0 + 4 * q) = make_uint4(off[4 * q], off[4 * q + 1], off[4 * q + 2], off[4 * q + 3]);
      *reinterpret_cast<uint4 *>(hcur + ho + i0 + 4 * q) = make_uint4(0, 0, 0, 0);
    }
  }
}

int launch_cluster(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  const int Wp = at_Wp(g);
  if (g.Wd < 3 || g.Hd < 3) return 0;
  cudaMemsetAsync(ws.hkey, 0, (size_t)nframes * g.hcap * sizeof(unsigned long long), s);
  cudaMemsetAsync(ws.hcnt, 0, (size_t)nframes * g.hcap * sizeof(uint32_t), s);
  const int segs = g.hcap > (uint32_t)kSelSeg ? (int)(g.hcap / kSelSeg) : 1;  // hcap is a power of two
  dim3 g4(((g.Wd + 3) / 4 + 255) / 256, g.Hd - 2, nframes);
  k_cluster_pass4<false><<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  k_cluster_select<<<nframes * segs, 1024, 0, s>>>(g, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.clusters, ws.counters, segs);
  k_cluster_pass4<true><<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  return 5;
}

}  // namespace b200at
