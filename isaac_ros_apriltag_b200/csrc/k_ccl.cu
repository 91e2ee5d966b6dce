// k_ccl.cu -- connected components of the threshold image (a8): white 8-connected, black 4-connected, 127
// ignored, with AprilRobotics' exact link set (do_unionfind_first_line / do_unionfind_line2, SURVEY App. A.3,
// including the link elisions, which DO change the partition at the first/last image columns).
//
// Design: label-equivalence union-find with atomicMin (representative = minimum pixel index, which is also
// the oracle's canonical label):
//   1. k_ccl_tile_sweep : ONE WARP per 32x32 tile (staged incl. halo by a TMA bulk tensor copy), rows top to bottom: runs by ballot,
//                         a run that touches the component above inherits its label, shared-memory unions only where a run
//                         touches two labels; pixel counts per local root
//   2. k_ccl_border     : only links that cross tile borders are merged in global memory (lists written by the tile kernel)
//   3. k_ccl_roots      : the former tile roots (the frame's root list, ~5 % of the pixels) chase to their global root and hand
//                         over their count
//   4. k_ccl_rootflag   : the size gate (< 25 px) is decided per tile root and stored in bit 31 of its label
//   5. k_ccl_flatmark   : one gather per pixel lab0[lab0[p]] -> final labels (lab) and thr2 (thr with small components -> 127)
// Algorithmic bytes per frame: read Pd (thr) + write 4*Pd (labels) = 5*Pd (SURVEY 8d contract figure).
#include "detector.h"

namespace b200at {

constexpr int TW = 32, TH = 32;  // CCL tile: one warp per row

struct Nb {
  bool L, U, UL, UR;
};

// Links AprilRobotics makes for "current" pixel (x, y) with value v (see file header).
__device__ __forceinline__ Nb ccl_links(int v, int vL, int vU, int vUL, int vUR, int x, int y, int Wd) {
  Nb n = {false, false, false, false};
  if (v == 127 || x < 1 || x > Wd - 2) return n;
  n.L = (vL == v);
  if (y >= 1) {
    n.U = (vU == v) && (x == 1 || !((vL == vUL) && (vUL == vU)));
    if (v == 255) {
      n.UL = (vUL == v) && (x == 1 || !(vL == vUL || vU == vUL));
      n.UR = (vUR == v) && !(vU == vUR);
    }
  }
  return n;
}

// find with path SPLITTING: every node on the walked path is re-pointed to its grandparent, so the pointer chases of
// the other lanes / later unions get shorter quickly (the walks of a warp diverge: the longest chain sets the pace).
// Safe under concurrent atomicMin unions: unions only ever modify ROOT entries, a non-root never becomes a root again,
// and every value written here is an ancestor of the node it is written to.
__device__ __forceinline__ uint32_t find_s(volatile uint32_t *L, uint32_t a) {
  uint32_t p = L[a];
  while (p != a) {
    const uint32_t gp = L[p];
    if (gp != p) L[a] = gp;
    a = p;
    p = gp;
  }
  return a;
}
__device__ __forceinline__ void unite_s(uint32_t *L, uint32_t a, uint32_t b) {
  bool done;
  do {
    a = find_s(L, a);
    b = find_s(L, b);
    if (a < b) {
      uint32_t old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      uint32_t old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}
__device__ __forceinline__ uint32_t find_g(uint32_t *L, uint32_t a) {
  uint32_t p = __ldcg(&L[a]);
  while (p != a) {
    const uint32_t gp = __ldcg(&L[p]);
    if (gp != p) __stcg(&L[a], gp);  // path splitting (see find_s)
    a = p;
    p = gp;
  }
  return a;
}
__device__ __forceinline__ void unite_g(uint32_t *L, uint32_t a, uint32_t b) {
  bool done;
  do {
    a = find_g(L, a);
    b = find_g(L, b);
    if (a < b) {
      uint32_t old = atomicMin(&L[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      uint32_t old = atomicMin(&L[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

constexpr int TPITCH = 64;  // shared-memory row pitch of the staged tile = TMA box width: x0-16 .. x0+47
constexpr int TOFF = 16;    // column of pixel x0 (TMA needs the box start 16-byte aligned: x0-16; the halo x0-1 is column 15)

// ---------------------------------------------------------------------------------------------------------------------
// Row-sweep variant of the tile kernel: ONE WARP per 32x32 tile (lane = column), rows top to bottom.  A run of a row that
// touches the component(s) above INHERITS a label (segmented min over the run, full-mask shuffles); shared-memory unions are
// only needed where a run touches two different labels.  On thresholded noise (0.5 vertical / diagonal links per pixel, run
// length 2.3) that replaces ~490 pointer-chasing unions per tile by a few dozen.  Per-run pixel counts go to the run's
// label and are moved to the final local roots at the end.  Same outputs as k_ccl_tile: lab = global index of the local
// root (minimum pixel index of the component inside the tile), csize = pixel count at local roots, 0 elsewhere.
// CTA = 4 warps = 4 horizontally adjacent tiles; each tile is staged (with halo) by its own TMA box.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int SWEEP_TILES = 4;
constexpr uint32_t kSmallFlag = 0x80000000u;  // bit 31 of a tile root's working label: its component has fewer than 25 pixels
constexpr int kReqCap = 192;  // cross-border links of a tile: <= 3 * 32 (row 0) + 32 + 31 + 31 (columns 0 / 31)
constexpr int TBYTES = TPITCH * (TH + 1);              // one staged tile + halo row
constexpr int TSLOT = (TBYTES + 127) & ~127;           // 128-byte aligned slots

template <bool USE_TMA>
__global__ void __launch_bounds__(32 * SWEEP_TILES) k_ccl_tile_sweep(Geo g, const uint8_t *__restrict__ thr, uint32_t *__restrict__ lab,
                                                                      uint32_t *__restrict__ csize, uint32_t *__restrict__ roots,
                                                                      uint32_t *__restrict__ nroots, uint2 *__restrict__ reqs,
                                                                      uint32_t *__restrict__ reqcnt, int Wp,
                                                                      const __grid_constant__ CUtensorMap tmap) {
  __shared__ __align__(128) uint8_t s_t[SWEEP_TILES][TSLOT];
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ uint32_t s_L[SWEEP_TILES][TH * TW];
  const int fr = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int x0 = (blockIdx.x * SWEEP_TILES + wid) * TW, y0 = blockIdx.y * TH;
  const uint8_t *img = thr + (size_t)fr * g.Hd * Wp;
  const int ntiles = min(SWEEP_TILES, (g.Wd - blockIdx.x * SWEEP_TILES * TW + TW - 1) / TW);  // tiles of this CTA inside the image
#ifndef B200AT_EMU
  if (USE_TMA) {
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((uint32_t)(TBYTES * ntiles)) : "memory");
      for (int w = 0; w < ntiles; w++) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&s_t[w][0]);
        const int bx = (blockIdx.x * SWEEP_TILES + w) * TW - TOFF;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(bx), "r"(y0 - 1), "r"(fr + g.tma_frame0), "r"(mb)
            : "memory");
      }
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_TMA_SWEEP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@!p bra WAIT_TMA_SWEEP;\n"
        "}\n" ::"r"(mb)
        : "memory");
  } else
#endif
  {
    // plain staging: (TH+1) x (TW+2) bytes per tile; out-of-image = 127 (never links)
    for (int w = 0; w < ntiles; w++) {
      const int wx0 = (blockIdx.x * SWEEP_TILES + w) * TW;
      for (int i = tid; i < (TH + 1) * (TW + 2); i += 32 * SWEEP_TILES) {
        const int r = i / (TW + 2), c = i % (TW + 2);
        const int y = y0 - 1 + r, x = wx0 - 1 + c;
        uint8_t v = 127;
        if (y >= 0 && y < g.Hd && x >= 0 && x < g.Wd) v = img[(size_t)y * Wp + x];
        s_t[w][r * TPITCH + c + TOFF - 1] = v;
      }
    }
    __syncthreads();
  }
  if (x0 >= g.Wd) return;  // (no block-wide barrier below this point)
  const uint8_t *t = s_t[wid];
  uint32_t *L = s_L[wid];
  const uint32_t NONE = 0xffffffffu;
  const int x = x0 + lane;
  const int rows = min(TH, g.Hd - y0);
  uint32_t plab = NONE;  // label of the pixel above (row ly - 1) in this lane's column
  unsigned dead = 0u;    // bit ly: my pixel of row ly is a 127-pixel (never connected, never counted) or a padding column
  for (int ly = 0; ly < rows; ly++) {
    const int y = y0 + ly;
    const uint8_t *tr = t + (ly + 1) * TPITCH + lane + TOFF, *tu = tr - TPITCH;
    Nb n = {false, false, false, false};
    if (x < g.Wd) n = ccl_links(tr[0], tr[-1], tu[0], tu[-1], tu[1], x, y, g.Wd);
    const int vcur = (int)tr[0];
    const unsigned ml = __ballot_sync(0xffffffffu, n.L && lane > 0);  // bit x: x is linked to x-1 inside the tile
    const unsigned upto = (2u << lane) - 1u;                          // lanes 0..lane (lane 31: all ones)
    const int rs = 31 - __clz(~ml & upto);                            // first lane of my run (bit 0 of ~ml is always set)
    const unsigned above = ~ml & ~upto;                               // run starts to the right of my lane
    const unsigned run_mask = (above ? ((1u << (__ffs(above) - 1)) - 1u) : 0xffffffffu) & ~((1u << rs) - 1u);
    const uint32_t pl_l = __shfl_up_sync(0xffffffffu, plab, 1), pl_r = __shfl_down_sync(0xffffffffu, plab, 1);
    uint32_t c1 = NONE, c2 = NONE, c3 = NONE;
    if (ly > 0) {
      if (n.U) c1 = plab;
      if (n.UL && lane > 0) c2 = pl_l;
      if (n.UR && lane < TW - 1) c3 = pl_r;
    }
    const int i = ly * TW + lane;
    // the run inherits the label of its LEFTMOST upward link (one ballot + one shuffle; any label of the component will do: the
    // unions below make a run's labels equivalent and the root of a component is its smallest label = its first pixel)
    const uint32_t mine = min(c1, min(c2, c3));
    const unsigned linked = __ballot_sync(0xffffffffu, mine != NONE) & run_mask;
    uint32_t rl = __shfl_sync(0xffffffffu, mine, linked ? __ffs(linked) - 1 : lane);
    if (!linked) rl = (uint32_t)(ly * TW + rs);  // no link upwards anywhere in the run: new label = its first pixel
    L[i] = rl;
    __syncwarp();
    // ONE union site, executed as often as the busiest lane needs it
    bool n1 = c1 != NONE && c1 != rl, n2 = c2 != NONE && c2 != rl && c2 != c1, n3 = c3 != NONE && c3 != rl && c3 != c1 && c3 != c2;
    while (__any_sync(0xffffffffu, n1 || n2 || n3)) {
      if (n1 || n2 || n3) {
        const uint32_t pick = n1 ? c1 : (n2 ? c2 : c3);
        if (n1)
          n1 = false;
        else if (n2)
          n2 = false;
        else
          n3 = false;
        unite_s(L, pick, rl);
      }
    }
    dead |= (x >= g.Wd || vcur == 127) ? (1u << ly) : 0u;
    plab = rl;
  }
  __syncwarp();
  // The tile's cross-border links (a few dozen, at most kReqCap) as a list of (pixel, neighbour) pairs: k_ccl_border no longer
  // re-reads the threshold image around every border pixel, it only walks the lists.  The links of row 0 (lane <-> column) and of
  // columns 0 / 31 (lane <-> row) are re-evaluated here from the staged tile: ~100 instructions per tile instead of ~10 per row.
  {
    const int tiles_x = (g.Wd + TW - 1) / TW, tile = blockIdx.y * tiles_x + blockIdx.x * SWEEP_TILES + wid;
    const size_t tl = (size_t)fr * tiles_x * gridDim.y + tile;
    uint2 *rq = reqs + tl * kReqCap;
    uint32_t nreq = 0;
    const unsigned lt = (1u << lane) - 1u;
    auto emit = [&](bool flag, uint32_t a, uint32_t b) {
      const unsigned bal = __ballot_sync(0xffffffffu, flag);
      if (flag) rq[nreq + __popc(bal & lt)] = make_uint2(a, b);
      nreq += __popc(bal);
    };
    auto links_at = [&](int lx, int ly) -> Nb {
      Nb n = {false, false, false, false};
      const uint8_t *tr = t + (ly + 1) * TPITCH + lx + TOFF, *tu = tr - TPITCH;
      if (ly < rows && x0 + lx < g.Wd) n = ccl_links(tr[0], tr[-1], tu[0], tu[-1], tu[1], x0 + lx, y0 + ly, g.Wd);
      return n;
    };
    const Nb n0 = links_at(lane, 0);
    const uint32_t me0 = (uint32_t)(y0 * Wp + x);
    emit(n0.U, me0, me0 - Wp);
    emit(n0.UL, me0, me0 - Wp - 1);
    emit(n0.UR, me0, me0 - Wp + 1);
    const Nb nl = links_at(0, lane), nr = links_at(TW - 1, lane);
    const uint32_t meL = (uint32_t)((y0 + lane) * Wp + x0), meR = meL + TW - 1;
    emit(nl.L, meL, meL - 1);
    emit(nl.UL && lane > 0, meL, meL - Wp - 1);  // (row 0 is in the first list)
    emit(nr.UR && lane > 0, meR, meR - Wp + 1);
    if (lane == 0) reqcnt[tl] = nreq;
  }
  __syncwarp();
  // The staged tile is dead from here on: its buffer becomes the pixel counters of the tile's labels (1024 x 16 bit; a tile has
  // 1024 pixels), which keeps the kernel at 6.2 KB of shared memory per tile (36 warps per SM instead of 20).
  uint32_t *cnt = reinterpret_cast<uint32_t *>(s_t[wid]);
  for (int k = lane; k < TH * TW / 2; k += 32) cnt[k] = 0u;
  __syncwarp();
  uint32_t *labf = lab + (size_t)fr * g.Hd * Wp;
  uint32_t *szf = csize + (size_t)fr * g.Hd * Wp;
  // flatten inside the tile and count the pixels of every local root (equal roots of a row are grouped by __match_any_sync:
  // one shared-memory atomic per root and row).  The rows in which this lane's column holds a local root are remembered as a
  // bit mask.
  unsigned rootbits = 0u;
  for (int ly = 0; ly < rows; ly++) {
    const int i = ly * TW + lane;
    const uint32_t r = find_s(L, i);
    if (x < g.Wd) labf[(size_t)(y0 + ly) * Wp + x] = (uint32_t)((y0 + r / TW) * Wp + (x0 + r % TW));
    const bool live = ((dead >> ly) & 1u) == 0u;
    const unsigned act = __ballot_sync(0xffffffffu, live);
    const unsigned bmask = (ly == 0 || ly == rows - 1) ? 0xffffffffu : 0x80000001u;  // lanes of this row on the tile border
    if (live) {
      const unsigned peers = __match_any_sync(act, r);
      if (lane == __ffs(peers) - 1) {
        atomicAdd(&cnt[r >> 1], (uint32_t)__popc(peers) << (16 * (r & 1u)));
        // bit 15 of a counter: the component touches the tile border (only those can merge with another tile)
        if (peers & bmask) atomicOr(&cnt[r >> 1], 0x8000u << (16 * (r & 1u)));
      }
      rootbits |= (r == (uint32_t)i) ? (1u << ly) : 0u;
    }
  }
  __syncwarp();
  // Local roots whose component does not touch the tile border are FINAL: global root = themselves, size = their count; the size
  // gate goes straight into bit 31 of their label (what k_ccl_rootflag does for the others).  Only border-touching roots go to the
  // frame's root list (on thresholded noise most components are specks inside a tile: the list shrinks several times).
  {
    unsigned listbits = 0u, rb = rootbits;
    while (rb) {
      const int ly = __ffs(rb) - 1;
      rb &= rb - 1u;
      const int i = ly * TW + lane;
      const uint32_t c16 = (cnt[i >> 1] >> (16 * (i & 1))) & 0xffffu;
      if (c16 & 0x8000u) {
        listbits |= 1u << ly;
      } else {
        const size_t gi = (size_t)(y0 + ly) * Wp + x;
        szf[gi] = c16;
        if (c16 < 25u) labf[gi] = (uint32_t)gi | kSmallFlag;
      }
    }
    rootbits = listbits;
  }
  // The tile's local roots (~5 % of the pixels) go to the frame's root list, their pixel counts to the size image: the later
  // kernels never scan the size image (it is written at roots only).  One reservation per tile, one loop trip per root of the
  // busiest column.
  const int myroots = __popc(rootbits);
  int incl = myroots;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += u;
  }
  uint32_t base = 0;
  if (lane == 31 && incl > 0) base = atomicAdd(&nroots[fr], (uint32_t)incl);
  base = __shfl_sync(0xffffffffu, base, 31) + (uint32_t)(incl - myroots);
  uint32_t *rootf = roots + (size_t)fr * g.Hd * Wp;
  while (rootbits) {
    const int ly = __ffs(rootbits) - 1;
    rootbits &= rootbits - 1u;
    const size_t gi = (size_t)(y0 + ly) * Wp + x;
    const int i = ly * TW + lane;
    szf[gi] = (cnt[i >> 1] >> (16 * (i & 1))) & 0x7fffu;
    rootf[base++] = (uint32_t)gi;
  }
}

// Cross-tile merges in global memory: one warp per tile walks the tile's list of (pixel, neighbour) links written by the tile kernel.
__global__ void __launch_bounds__(128) k_ccl_border(Geo g, uint32_t *__restrict__ lab, const uint2 *__restrict__ reqs,
                                                    const uint32_t *__restrict__ reqcnt, int ntiles, int Wp) {
  const int fr = blockIdx.y;
  const int lane = threadIdx.x & 31, tile = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (tile >= ntiles) return;
  const size_t tl = (size_t)fr * ntiles + tile;
  const uint32_t n = min(reqcnt[tl], (uint32_t)kReqCap);
  const uint2 *rq = reqs + tl * kReqCap;
  uint32_t *labf = lab + (size_t)fr * g.Hd * Wp;
  for (uint32_t i = lane; i < n; i += 32) {
    const uint2 r = rq[i];
    unite_g(labf, r.x, r.y);
  }
}

// Flatten + size gate in two phases.  After the tile and border kernels every pixel points at a (former) tile root, and only
// those carry a count.  Phase A (k_ccl_roots): the former tile roots -- the frame's root list written by the tile kernel, ~5 % of
// the pixels -- chase to their global root, point at it directly and hand over their count.  Phase B (k_ccl_flatmark): every
// pixel needs exactly ONE gather, lab[lab[p]], and the size gate (thr2 = thr with pixels of components < 25 px forced to 127:
// folds the size gates of gradient_clusters into one byte image) is applied in the same pass: lab and thr are read once.
constexpr int ROOT_CTAS = 48;  // per frame, grid-stride over the root list
__global__ void __launch_bounds__(256) k_ccl_roots(Geo g, uint32_t *__restrict__ lab, uint32_t *__restrict__ csize,
                                                   const uint32_t *__restrict__ roots, const uint32_t *__restrict__ nroots, int Wp) {
  const int fr = blockIdx.y;
  const size_t fo = (size_t)fr * g.Hd * Wp;
  uint32_t *L = lab + fo;
  const uint32_t n = nroots[fr];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t me = roots[fo + i];
    uint32_t a = me, p = __ldcg(&L[a]);
    while (p != a) {
      a = p;
      p = __ldcg(&L[a]);
    }
    if (a != me) {
      __stcg(&L[me], a);
      atomicAdd(&csize[fo + a], csize[fo + me]);  // (only global roots receive counts: csize[me] is final)
    }
  }
}

// Phase A': the size gate is decided once per (former) tile root and travels in bit 31 of its label, so that phase B needs no
// second gather into the size image.
__global__ void __launch_bounds__(256) k_ccl_rootflag(Geo g, uint32_t *__restrict__ lab, const uint32_t *__restrict__ csize,
                                                      const uint32_t *__restrict__ roots, const uint32_t *__restrict__ nroots, int Wp) {
  const int fr = blockIdx.y;
  const size_t fo = (size_t)fr * g.Hd * Wp;
  uint32_t *L = lab + fo;
  const uint32_t n = nroots[fr];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t me = roots[fo + i];
    const uint32_t a = L[me];  // the global root (k_ccl_roots); only this thread writes L[me]
    if (csize[fo + a] < 25u) L[me] = a | kSmallFlag;
  }
}

__global__ void __launch_bounds__(256) k_ccl_flatmark(Geo g, const uint8_t *__restrict__ thr, const uint32_t *__restrict__ lab0,
                                                      uint32_t *__restrict__ lab, uint8_t *__restrict__ thr2, int Wp) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  const int fr = blockIdx.z;
  if (x4 >= g.Wd) return;
  const size_t fo = (size_t)fr * g.Hd * Wp;
  const uint32_t *L = lab0 + fo;  // (read only here: the flags of the tile roots must survive until every pixel has seen them)
  const uint32_t me0 = (uint32_t)(y * Wp + x4);
  const uchar4 tv = *reinterpret_cast<const uchar4 *>(thr + fo + me0);
  const uint4 pv = *reinterpret_cast<const uint4 *>(L + me0);
  uint8_t v[4] = {tv.x, tv.y, tv.z, tv.w};
  const uint32_t p[4] = {pv.x, pv.y, pv.z, pv.w};
  uint32_t r[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    // AprilRobotics never connects 127 pixels (singletons); padding columns are ignored
    const bool live = v[k] != 127 && x4 + k < g.Wd;
    r[k] = live ? L[p[k] & ~kSmallFlag] : me0 + k;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (r[k] & kSmallFlag) v[k] = 127;
    r[k] &= ~kSmallFlag;
  }
  *reinterpret_cast<uint4 *>(lab + fo + me0) = make_uint4(r[0], r[1], r[2], r[3]);
  *reinterpret_cast<uchar4 *>(thr2 + fo + me0) = make_uchar4(v[0], v[1], v[2], v[3]);
}

int launch_ccl(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  const int Wp = at_Wp(g);
  dim3 gt((g.Wd + TW - 1) / TW, (g.Hd + TH - 1) / TH, nframes);
  dim3 gs((gt.x + SWEEP_TILES - 1) / SWEEP_TILES, gt.y, gt.z);
  cudaMemsetAsync(ws.nroots, 0, sizeof(uint32_t) * nframes, s);
  // Tune::ccl_tma: 1 = tiles staged in shared memory by TMA, 0 = staged by plain loads (also the path without the driver entry point)
  if (ws.tune.ccl_tma && ws.use_tma)
    k_ccl_tile_sweep<true><<<gs, 32 * SWEEP_TILES, 0, s>>>(g, ws.thr, ws.lab0, ws.csize, ws.roots, ws.nroots, ws.ccl_req, ws.ccl_reqcnt, Wp, ws.thr_tmap);
  else
    k_ccl_tile_sweep<false><<<gs, 32 * SWEEP_TILES, 0, s>>>(g, ws.thr, ws.lab0, ws.csize, ws.roots, ws.nroots, ws.ccl_req, ws.ccl_reqcnt, Wp, ws.thr_tmap);
  const int ntl = (int)(gt.x * gt.y);
  k_ccl_border<<<dim3((ntl + 3) / 4, nframes), 128, 0, s>>>(g, ws.lab0, ws.ccl_req, ws.ccl_reqcnt, ntl, Wp);
  k_ccl_roots<<<dim3(ROOT_CTAS, nframes), 256, 0, s>>>(g, ws.lab0, ws.csize, ws.roots, ws.nroots, Wp);
  dim3 gp(((g.Wd + 3) / 4 + 255) / 256, g.Hd, nframes);
  k_ccl_rootflag<<<dim3(ROOT_CTAS, nframes), 256, 0, s>>>(g, ws.lab0, ws.csize, ws.roots, ws.nroots, Wp);
  k_ccl_flatmark<<<gp, 256, 0, s>>>(g, ws.thr, ws.lab0, ws.lab, ws.thr2, Wp);
  return 6;
}

}  // namespace b200at
