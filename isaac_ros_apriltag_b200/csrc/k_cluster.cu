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

// The four probes of pixel (x, y).  thr2 already folds both size gates (components < 25 px read as 127), so a
// point exists iff v0 + v1 == 255.  `connected_last`: the (-1,1) probe is skipped when the previous pixel's
// (1,1) probe produced a point; at x == 1 there is no previous pixel.
struct Probes {
  bool p[4];
};
__device__ __forceinline__ Probes eval_probes(const uint8_t *img, int Wp, int x, int y) {
  const uint8_t *r0 = img + (size_t)y * Wp, *r1 = r0 + Wp;
  int v0 = r0[x];
  Probes pr;
  int vl = r0[x - 1], vr = r0[x + 1], dl = r1[x - 1], dc = r1[x], dr = r1[x + 1];
  bool prev_conn = (x > 1) && (vl + dc == 255);
  pr.p[0] = (v0 + vr == 255);
  pr.p[1] = (v0 + dc == 255);
  pr.p[2] = !prev_conn && (v0 + dl == 255);
  pr.p[3] = (v0 + dr == 255);
  return pr;
}

// cheap 64->32 bit mix for the (rep_hi, rep_lo) keys
__device__ __forceinline__ uint32_t hash_key2(unsigned long long k) {
  uint32_t h = (uint32_t)(k >> 32) * 0x9E3779B1u ^ (uint32_t)k * 0x85EBCA6Bu;
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 13;
  return h;
}

// One thread per pixel evaluates the four probes; the (on average < 1 per pixel) resulting points are COMPACTED per
// warp through shared memory so the expensive part (match / table probe / atomic) runs on dense lanes: one round per
// 32 points instead of four rounds per 32 pixels.
// EAGER: the label loads do not wait for the probe results and the segment offset is loaded together with the table key, so a
// warp's critical path has three dependent memory round trips (pixels + labels, table slot, atomic) instead of five.
template <bool EMIT, bool EAGER>
__global__ void __launch_bounds__(256) k_cluster_pass(Geo g, const uint8_t *__restrict__ thr2, const uint32_t *__restrict__ lab,
                                                      unsigned long long *__restrict__ hkey, uint32_t *__restrict__ hcnt,
                                                      const uint32_t *__restrict__ hoff, uint32_t *__restrict__ hcur,
                                                      uint32_t *__restrict__ pts, uint32_t *__restrict__ counters, int Wp) {
  __shared__ unsigned long long s_key[8][128];
  __shared__ uint32_t s_pt[8][128];
  const int x = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int y = blockIdx.y + 1;
  const int fr = blockIdx.z;
  const size_t fo = (size_t)fr * g.Hd * Wp;
  const uint8_t *img = thr2 + fo;
  const uint32_t *labf = lab + fo;
  unsigned long long *hk = hkey + (size_t)fr * g.hcap;
  const size_t ho = (size_t)fr * g.hcap;
  const uint32_t hmask = g.hcap - 1;
  const bool in = (x <= g.Wd - 2) && (y <= g.Hd - 2);
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  Probes pr = {{false, false, false, false}};
  uint32_t rep0 = 0;
  int v0 = 127;
  const int dxs[4] = {1, 0, -1, 1};
  const int dys[4] = {0, 1, 1, 1};
  uint32_t rep1[4];
  if (EAGER) {
    if (in) {
      // (x, y) is interior: all five label addresses are inside the frame
      rep0 = labf[(size_t)y * Wp + x];
#pragma unroll
      for (int k = 0; k < 4; k++) rep1[k] = labf[(size_t)(y + dys[k]) * Wp + x + dxs[k]];
      pr = eval_probes(img, Wp, x, y);
      v0 = img[(size_t)y * Wp + x];
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) rep1[k] = 0u;
    }
  } else {
    if (in) {
      pr = eval_probes(img, Wp, x, y);
      v0 = img[(size_t)y * Wp + x];
      if (pr.p[0] || pr.p[1] || pr.p[2] || pr.p[3]) rep0 = labf[(size_t)y * Wp + x];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) rep1[k] = pr.p[k] ? labf[(size_t)(y + dys[k]) * Wp + x + dxs[k]] : 0u;
  }
  // compaction: probe-major order inside the warp
  int total = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const unsigned m = __ballot_sync(0xffffffffu, pr.p[k]);
    if (pr.p[k]) {
      const int pos = total + __popc(m & ((1u << lane) - 1));
      const uint32_t r1 = rep1[k];
      s_key[wid][pos] = rep0 < r1 ? (((unsigned long long)r1 << 32) | rep0) : (((unsigned long long)rep0 << 32) | r1);
      if (EMIT) {
        const int dx = dxs[k], dy = dys[k];
        const int d = 255 - 2 * v0;  // v1 - v0 with v0 + v1 == 255
        const int gx = dx * d, gy = dy * d;
        const uint32_t cx = gx == 0 ? 0u : (gx > 0 ? 1u : 2u), cy = gy == 0 ? 0u : (gy > 0 ? 1u : 2u);
        // packed point: x (14 bits) | y (14 bits) | gx code (2) | gy code (2); code 0 = 0, 1 = +255, 2 = -255
        s_pt[wid][pos] = (uint32_t)(2 * x + dx) | ((uint32_t)(2 * y + dy) << 14) | (cx << 28) | (cy << 30);
      }
    }
    total += __popc(m);
  }
  __syncwarp();
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
    if (!EMIT) {
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
    } else {
      uint32_t base = 0xffffffffu;
      if ((int)lane == leader) {
        for (uint32_t probe = 0; probe < g.hcap; probe++) {
          unsigned long long cur = hk[slot];
          uint32_t off = 0;
          if (EAGER) off = hoff[ho + slot];  // issued together with the key load
          if (cur == key) {
            if (!EAGER) off = hoff[ho + slot];
            if (off != 0xffffffffu) base = off + atomicAdd(&hcur[ho + slot], (uint32_t)n);
            break;
          }
          if (cur == 0ULL) break;
          slot = (slot + 1) & hmask;
        }
      }
      base = __shfl_sync(peers, base, leader);
      if (base != 0xffffffffu) pts[base + __popc(peers & ((1u << lane) - 1))] = s_pt[wid][j];
    }
  }
}

// Four pixels per thread (x = 4t .. 4t+3 of row y): the byte image is read as aligned words (three per row instead of seven
// byte loads per pixel), the labels as one uint4 + scalars per row, and the points of a warp are compacted with ONE exclusive
// scan of per-thread counts per half instead of one ballot per (pixel, probe).  Same table protocol as k_cluster_pass; the
// kernels are instruction-issue bound (ncu: 66 % of peak issue at 82 % warps active), this variant executes about half the
// instructions per pixel.  The order of the points inside a cluster's segment differs, which is irrelevant (sorted later).
// DEFER (emit pass, cluster_eager=3): ncu puts 35 % of the emit pass's stall samples on the wait for the segment-cursor atomic's
// return value.  Here a round's points are stored one round LATER: the leader issues the atomic, the warp goes on to the next
// round's match / probe, and only then picks up the previous round's base -- the atomic has had a whole round to come back.
// RECORD (cluster_eager=4, not measured yet): the count pass also writes every point, with the table slot it was counted in,
// to a per-row record list; the emit pass -- which repeats the whole count pass (pixels, labels, probes, match, table probe)
// only to learn where each point goes -- is replaced by k_cluster_scatter, which reads the 8-byte records.
struct RecArgs {
  uint2 *rec;         // [frame][row][cap]: .x = table slot of the point's cluster, .y = packed point
  uint32_t *cnt;      // [frame][row]
  int cap;
};

template <bool EMIT, bool DEFER, bool RECORD>
__device__ __forceinline__ void cluster_pass4_body(const Geo &g, const uint8_t *__restrict__ thr2, const uint32_t *__restrict__ lab,
                                                   unsigned long long *__restrict__ hkey, uint32_t *__restrict__ hcnt,
                                                   const uint32_t *__restrict__ hoff, uint32_t *__restrict__ hcur,
                                                   uint32_t *__restrict__ pts, uint32_t *__restrict__ counters, int Wp, RecArgs ra) {
  __shared__ unsigned long long s_key[8][256];
  __shared__ uint32_t s_pt[8][256];
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
  const int dxs[4] = {1, 0, -1, 1};
  const int dys[4] = {0, 1, 1, 1};
  // DEFER: the round whose points have not been stored yet
  bool p_has = false;
  unsigned p_peers = 0;
  int p_leader = 0;
  uint32_t p_base = 0xffffffffu, p_pt = 0;
  auto flush_pending = [&]() {
    if (p_has) {
      const uint32_t b = __shfl_sync(p_peers, p_base, p_leader);
      if (b != 0xffffffffu) pts[b + __popc(p_peers & ((1u << lane) - 1))] = p_pt;
    }
    p_has = false;
  };
#pragma unroll
  for (int half = 0; half < 2; half++) {
    // my points of pixels 2*half, 2*half+1: bit (2 * jj + k) of `mask` = probe k of pixel jj produced a point
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
    const int base0 = incl - cnt;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      const int j = 2 * half + jj;
      const int x = x4 + j;
      const int v0 = va[1 + j];
      const uint32_t rep0 = lrow0[j];
      const uint32_t rep1[4] = {lrow0[j + 1], lrow1[j + 1], lrow1[j], lrow1[j + 2]};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int bit = 4 * jj + k;
        if ((mask >> bit) & 1u) {
          const int pos = base0 + __popc(mask & ((1u << bit) - 1u));
          const uint32_t q1 = rep1[k];
          s_key[wid][pos] = rep0 < q1 ? (((unsigned long long)q1 << 32) | rep0) : (((unsigned long long)rep0 << 32) | q1);
          if (EMIT || RECORD) {
            const int dx = dxs[k], dy = dys[k];
            const int d = 255 - 2 * v0;  // v1 - v0 with v0 + v1 == 255
            const int gx = dx * d, gy = dy * d;
            const uint32_t cx = gx == 0 ? 0u : (gx > 0 ? 1u : 2u), cy = gy == 0 ? 0u : (gy > 0 ? 1u : 2u);
            s_pt[wid][pos] = (uint32_t)(2 * x + dx) | ((uint32_t)(2 * y + dy) << 14) | (cx << 28) | (cy << 30);
          }
        }
      }
    }
    __syncwarp();
    if (RECORD) {
      // room in the row's record list for this warp's points: the atomic is issued now and picked up after the first table probe
      uint32_t rb = 0;
      uint2 *rrow = ra.rec + ((size_t)fr * g.Hd + y) * ra.cap;
      if (lane == 0 && total > 0) rb = atomicAdd(&ra.cnt[(size_t)fr * g.Hd + y], (uint32_t)total);
      bool rb_ready = false;
      for (int c0 = 0; c0 < total; c0 += 32) {  // (every lane walks every round: the pick-up below is a full-warp shuffle)
        const int j = c0 + (int)lane;
        const bool has = j < total;
        const unsigned act = __ballot_sync(0xffffffffu, has);
        uint32_t found = 0xffffffffu, pt = 0;
        if (has) {
          const unsigned long long key = s_key[wid][j];
          pt = s_pt[wid][j];
          const unsigned peers = __match_any_sync(act, key);
          const int leader = __ffs(peers) - 1;
          if ((int)lane == leader) {
            uint32_t slot = hash_key2(key) & hmask;
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
              atomicAdd(&hcnt[ho + found], (uint32_t)__popc(peers));
          }
          found = __shfl_sync(peers, found, leader);
        }
        if (!rb_ready) {
          rb = __shfl_sync(0xffffffffu, rb, 0);
          rb_ready = true;
        }
        if (has) {
          if (rb + (uint32_t)j < (uint32_t)ra.cap)
            rrow[rb + j] = make_uint2(found, pt);
          else
            atomicOr(&counters[CNT_STATUS], (uint32_t)ST_POINTS_FULL);
        }
      }
    } else if (EMIT && DEFER) {
      // every lane walks every round (no early `continue`): the deferred store of the previous round is one warp-uniform
      // program point, reached together by all lanes of that round's groups
      for (int c0 = 0; c0 < total; c0 += 32) {
        const int j = c0 + (int)lane;
        const bool has = j < total;
        const unsigned act = __ballot_sync(0xffffffffu, has);
        unsigned peers = 0;
        int leader = 0;
        uint32_t base = 0xffffffffu, pt = 0;
        if (has) {
          const unsigned long long key = s_key[wid][j];
          pt = s_pt[wid][j];
          peers = __match_any_sync(act, key);
          leader = __ffs(peers) - 1;
          if ((int)lane == leader) {
            const int n = __popc(peers);
            uint32_t slot = hash_key2(key) & hmask;
            for (uint32_t probe = 0; probe < g.hcap; probe++) {
              const unsigned long long cur = hk[slot];
              const uint32_t off = hoff[ho + slot];
              if (cur == key) {
                if (off != 0xffffffffu) base = off + atomicAdd(&hcur[ho + slot], (uint32_t)n);  // consumed one round later
                break;
              }
              if (cur == 0ULL) break;
              slot = (slot + 1) & hmask;
            }
          }
        }
        flush_pending();
        p_has = has;
        p_peers = peers;
        p_leader = leader;
        p_base = base;
        p_pt = pt;
      }
    } else
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
      if (!EMIT) {
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
      } else {
        uint32_t base = 0xffffffffu;
        if ((int)lane == leader) {
          for (uint32_t probe = 0; probe < g.hcap; probe++) {
            const unsigned long long cur = hk[slot];
            const uint32_t off = hoff[ho + slot];  // issued together with the key load
            if (cur == key) {
              if (off != 0xffffffffu) base = off + atomicAdd(&hcur[ho + slot], (uint32_t)n);
              break;
            }
            if (cur == 0ULL) break;
            slot = (slot + 1) & hmask;
          }
        }
        base = __shfl_sync(peers, base, leader);
        if (base != 0xffffffffu) pts[base + __popc(peers & ((1u << lane) - 1))] = s_pt[wid][j];
      }
    }
    __syncwarp();  // the buffer is reused by the second half
  }
  if (EMIT && DEFER) flush_pending();
}

template <bool EMIT, bool DEFER = false>
__global__ void __launch_bounds__(256) k_cluster_pass4(Geo g, const uint8_t *__restrict__ thr2, const uint32_t *__restrict__ lab,
                                                       unsigned long long *__restrict__ hkey, uint32_t *__restrict__ hcnt,
                                                       const uint32_t *__restrict__ hoff, uint32_t *__restrict__ hcur,
                                                       uint32_t *__restrict__ pts, uint32_t *__restrict__ counters, int Wp) {
  cluster_pass4_body<EMIT, DEFER, false>(g, thr2, lab, hkey, hcnt, hoff, hcur, pts, counters, Wp, RecArgs{nullptr, nullptr, 0});
}

__global__ void __launch_bounds__(256) k_cluster_record(Geo g, const uint8_t *__restrict__ thr2, const uint32_t *__restrict__ lab,
                                                        unsigned long long *__restrict__ hkey, uint32_t *__restrict__ hcnt,
                                                        uint32_t *__restrict__ counters, int Wp, RecArgs ra) {
  cluster_pass4_body<false, false, true>(g, thr2, lab, hkey, hcnt, nullptr, nullptr, nullptr, counters, Wp, ra);
}

// one CTA per (row, frame): the row's records -> their clusters' segments (same aggregation as the emit pass: lanes that hold
// points of one cluster share one atomic on its cursor)
__global__ void __launch_bounds__(256) k_cluster_scatter(Geo g, RecArgs ra, const uint32_t *__restrict__ hoff, uint32_t *__restrict__ hcur,
                                                         uint32_t *__restrict__ pts) {
  const int y = blockIdx.x + 1, fr = blockIdx.y;
  const size_t ho = (size_t)fr * g.hcap;
  const uint32_t n = min(ra.cnt[(size_t)fr * g.Hd + y], (uint32_t)ra.cap);
  const uint2 *rrow = ra.rec + ((size_t)fr * g.Hd + y) * ra.cap;
  const unsigned lane = threadIdx.x & 31;
  for (uint32_t i0 = 0; i0 < n; i0 += blockDim.x) {
    const uint32_t i = i0 + threadIdx.x;
    uint32_t off = 0xffffffffu, slot = 0, pt = 0;
    if (i < n) {
      const uint2 r = rrow[i];
      slot = r.x;
      pt = r.y;
      if (slot != 0xffffffffu) off = hoff[ho + slot];
    }
    const bool has = off != 0xffffffffu;
    const unsigned act = __ballot_sync(0xffffffffu, has);
    if (!has) continue;
    const unsigned peers = __match_any_sync(act, slot);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&hcur[ho + slot], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    pts[off + base + __popc(peers & ((1u << lane) - 1))] = pt;
  }
}

// One CTA per (frame, table segment).  Pass 1 totals -> one reservation in the global cluster / point pools,
// pass 2 ordered allocation inside the reservation (clusters of a segment appear in table-slot order).
__global__ void __launch_bounds__(1024) k_cluster_select(Geo g, const unsigned long long *__restrict__ hkey,
                                                         const uint32_t *__restrict__ hcnt, uint32_t *__restrict__ hoff,
                                                         uint32_t *__restrict__ hcur, ClusterRec *__restrict__ clusters,
                                                         uint32_t *__restrict__ counters, int segs) {
  const int fr = blockIdx.x / segs, seg = blockIdx.x % segs;
  const uint32_t seg_len = g.hcap / (uint32_t)segs;
  const size_t ho = (size_t)fr * g.hcap + (size_t)seg * seg_len;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  __shared__ uint32_t s_c[32], s_p[32];
  __shared__ uint32_t s_base_c, s_base_p, s_run_c, s_run_p, s_ok;
  // pass 1
  uint32_t nc = 0, np = 0;
  for (uint32_t i = tid; i < seg_len; i += 1024) {
    uint32_t c = hcnt[ho + i];
    if (c >= 24u && c <= g.max_cluster_pts) {
      nc++;
      np += c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    nc += __shfl_xor_sync(0xffffffffu, nc, o);
    np += __shfl_xor_sync(0xffffffffu, np, o);
  }
  if (lane == 0) {
    s_c[wid] = nc;
    s_p[wid] = np;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t tc = 0, tp = 0;
    for (int w = 0; w < 32; w++) {
      tc += s_c[w];
      tp += s_p[w];
    }
    uint32_t bc = atomicAdd(&counters[CNT_CLUSTERS], tc);
    uint32_t bp = atomicAdd(&counters[CNT_POINTS], tp);
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
    s_run_c = 0;
    s_run_p = 0;
    s_ok = ok;
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  // pass 2: chunked block scan in slot order
  for (uint32_t i0 = 0; i0 < seg_len; i0 += 1024) {
    const uint32_t i = i0 + tid;
    uint32_t c = (i < seg_len) ? hcnt[ho + i] : 0;
    const bool would_keep = c >= 24u && c <= g.max_cluster_pts;
    const bool keep = ok && would_keep;
    uint32_t fc = would_keep ? 1u : 0u, fp = would_keep ? c : 0u;  // (positions are those of the reservation, kept or not)
    // inclusive warp scan
    uint32_t ic = fc, ip = fp;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t tc = __shfl_up_sync(0xffffffffu, ic, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
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
      uint32_t vc = s_c[lane], vp = s_p[lane];
      uint32_t jc = vc, jp = vp;
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t tc = __shfl_up_sync(0xffffffffu, jc, o), tp = __shfl_up_sync(0xffffffffu, jp, o);
        if (lane >= o) {
          jc += tc;
          jp += tp;
        }
      }
      s_c[lane] = jc - vc;  // exclusive warp offsets
      s_p[lane] = jp - vp;
    }
    __syncthreads();
    const uint32_t run_c = s_run_c, run_p = s_run_p;
    const uint32_t ec = run_c + s_c[wid] + ic - fc;  // exclusive index of this slot's cluster
    const uint32_t ep = run_p + s_p[wid] + ip - fp;
    if (i < seg_len) {
      if (keep) {
        ClusterRec r;
        r.key = hkey[ho + i];
        r.offset = s_base_p + ep;
        r.count = c;
        r.frame = (uint32_t)fr;
        r.pad = 0;
        clusters[s_base_c + ec] = r;
        hoff[ho + i] = s_base_p + ep;
      } else {
        // A reservation that did not fit keeps its range of the cluster list (the counter stays advanced): its records are
        // written EMPTY (count 0), so that a consumer that walks [0, min(CNT_CLUSTERS, clu_cap)) never reads a stale record.
        if (!ok && would_keep && s_base_c + ec < g.clu_cap) clusters[s_base_c + ec] = ClusterRec{hkey[ho + i], 0u, 0u, (uint32_t)fr, 0u};
        hoff[ho + i] = 0xffffffffu;
      }
      hcur[ho + i] = 0;
    }
    __syncthreads();
    if (tid == 1023) {
      s_run_c = ec + fc;
      s_run_p = ep + fp;
    }
    __syncthreads();
  }
}

int launch_cluster(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  const int Wp = at_Wp(g);
  if (g.Wd < 3 || g.Hd < 3) return 0;
  cudaMemsetAsync(ws.hkey, 0, (size_t)nframes * g.hcap * sizeof(unsigned long long), s);
  cudaMemsetAsync(ws.hcnt, 0, (size_t)nframes * g.hcap * sizeof(uint32_t), s);
  dim3 gp((g.Wd - 2 + 255) / 256, g.Hd - 2, nframes);
  const int segs = g.hcap >= 65536 ? 4 : 1;  // hcap is a power of two
  if (ws.tune.cluster_eager == 4 && ws.rec) {
    dim3 g4(((g.Wd + 3) / 4 + 255) / 256, g.Hd - 2, nframes);
    RecArgs ra{ws.rec, ws.rec_cnt, ws.rec_cap};
    cudaMemsetAsync(ws.rec_cnt, 0, (size_t)nframes * g.Hd * sizeof(uint32_t), s);
    k_cluster_record<<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.counters, Wp, ra);
    k_cluster_select<<<nframes * segs, 1024, 0, s>>>(g, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.clusters, ws.counters, segs);
    k_cluster_scatter<<<dim3(g.Hd - 2, nframes), 256, 0, s>>>(g, ra, ws.hoff, ws.hcur, ws.pts);
    return 6;
  }
  if (ws.tune.cluster_eager >= 2) {
    dim3 g4(((g.Wd + 3) / 4 + 255) / 256, g.Hd - 2, nframes);
    k_cluster_pass4<false><<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
    k_cluster_select<<<nframes * segs, 1024, 0, s>>>(g, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.clusters, ws.counters, segs);
    if (ws.tune.cluster_eager == 3)
      k_cluster_pass4<true, true><<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
    else
      k_cluster_pass4<true><<<g4, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
    return 5;
  }
  if (ws.tune.cluster_eager)
    k_cluster_pass<false, true><<<gp, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  else
    k_cluster_pass<false, false><<<gp, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  k_cluster_select<<<nframes * segs, 1024, 0, s>>>(g, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.clusters, ws.counters, segs);
  if (ws.tune.cluster_eager)
    k_cluster_pass<true, true><<<gp, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  else
    k_cluster_pass<true, false><<<gp, 256, 0, s>>>(g, ws.thr2, ws.lab, ws.hkey, ws.hcnt, ws.hoff, ws.hcur, ws.pts, ws.counters, Wp);
  return 5;
}

}  // namespace b200at
