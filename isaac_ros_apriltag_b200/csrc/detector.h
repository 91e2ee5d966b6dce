// detector.h -- internal layout of the B200 AprilTag detector workspace and kernel entry points.
//
// HBM layout (all buffers batch-major, one slab per frame so a CTA never straddles frames):
//   dec   u8  [B][Hd][Wd]   decimated (+blurred) gray        a5/a6 of SURVEY.md section 8a
//   tmin/tmax u8 [B][th][tw] per-tile min / max (pre-dilation) a7
//   thr   u8  [B][Hd][Wd]   {0,127,255}                       a7
//   lab   u32 [B][Hd][Wd]   union-find parent -> min-index representative   a8
//   csize u32 [B][Hd][Wd]   component size, stored at the representative    a8
//   thr2  u8  [B][Hd][Wd]   thr with pixels of components < 25 px forced to 127 (folds the size gate of a9)
//   hkey/hcnt/hoff/hcur     per-frame open-addressing table keyed by (rep_hi<<32|rep_lo)   a9
//   pts   u32 pool          packed boundary points of kept clusters, cluster-contiguous    a9
//   keys  u64 pool          (slope|y|x) sort keys, sorted per cluster                      a10
//   lfps  6xf64 pool        sequential weighted prefix moments                             a10
//   quads / cands / out     fixed-capacity record lists                                    a10-a15
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200_apriltags.h"

namespace b200at {

constexpr int kMaxFamilies = B200AT_MAX_FAMILIES;
constexpr int kMaxBits = 52;
constexpr int kMaxNMaxima = 12;  // upper bound of the max_nmaxima option (pair tables live in shared memory)

struct DevFamily {
  int nbits, ncodes, width_at_border, total_width, reversed_border, index;
  int8_t bit_x[kMaxBits], bit_y[kMaxBits];  // relative to the border; negative / >= width_at_border = outside it
  const unsigned long long *codes;  // device pointer
};

struct FrameDesc {
  const uint8_t *ptr;
  unsigned long long pitch;
};

struct LineFitPt {
  double Mx, My, Mxx, Mxy, Myy, W;
};

struct ClusterRec {  // == b200AprilTagsClusterRec_t
  unsigned long long key;
  uint32_t offset, count, frame, pad;
};

struct QuadRec {  // == b200AprilTagsQuadRec_t
  unsigned long long key;
  float p[4][2];
  uint32_t frame;
  uint32_t reversed_border;
};

struct Cand {  // decoded candidate before reconcile
  unsigned long long key;
  int32_t family, id, hamming;
  float decision_margin;
  double H[9];
  double c[2];
  double p[4][2];
};

// counters[] slots
enum { CNT_CLUSTERS = 0, CNT_POINTS = 1, CNT_QUADS = 2, CNT_STATUS = 3, CNT_CANDS = 4, CNT_DETS = 5,
       CNT_FETCHED = 6 /* 16-byte chunks fetched on demand from the caller's host frames (sparse host path) */, CNT_WORK_DECODE = 7,
       CNT_BIN0 = 8 /* ..15: clusters per size bin */, CNT_WORK0 = 16 /* ..23: quad-fit work queues */,
       CNT_QWORK = 24 /* (cluster, chunk) work items registered by k_qf_sort */, CNT_Q2 = 25, CNT_Q3 = 26 /* work queues of k_qf_window / k_qf_tail */,
       CNT_N = 32 };
constexpr int kQuadBins = 8;
// windowed quad fit (k_quad2.cu): a warp of k_qf_window holds kQfSlotsPerLane * 32 sorted points (chunk + circular halo of at most
// 2 * 20 + 9 points) in shared memory; a cluster of n points is cut into qf_nchunks(n) balanced chunks
constexpr int kQfSlotsPerLane = 7;   // odd: the lane-blocked 8-byte accesses are bank-conflict free
constexpr int kQfChunkMax = 32 * kQfSlotsPerLane - 49;
__host__ __device__ inline int qf_nchunks(int n) { return (n + kQfChunkMax - 1) / kQfChunkMax; }
constexpr int kQuadAux = kQuadBins - 1;  // side streams: the quad-fit bins run concurrently
constexpr int kMaxChunks = 4;  // a batch is processed as up to 4 frame chunks on two phase-shifted streams
enum { ST_HASH_FULL = 1, ST_POINTS_FULL = 2, ST_CLUSTERS_FULL = 4, ST_QUADS_FULL = 8, ST_CANDS_FULL = 16, ST_OUT_TRUNC = 32 };

struct Geo {
  int W, H;      // input frame
  int Wd, Hd;    // decimated
  int tw, th;    // full tiles
  int f;         // integer decimation factor
  int ts;        // tile size
  int enc;       // B200AT_ENC_*
  int bpp;       // bytes per input pixel
  int min_wb_diff;
  int fast_align;  // every frame pointer/pitch of the current batch is 16-byte aligned -> vector load path
  uint32_t hcap;       // hash slots per frame (power of two)
  uint32_t pts_cap;    // point pool capacity (whole batch)
  uint32_t clu_cap;    // cluster list capacity (whole batch)
  uint32_t quad_cap;   // quad list capacity (whole batch)
  uint32_t cand_cap;   // candidates per frame
  uint32_t max_tags;   // outputs per frame
  uint32_t max_cluster_pts;  // 2*(2*Wd+2*Hd)
  int tma_frame0;      // frame offset of this workspace view inside the whole-batch TMA tensor
  // sparse host path (capi.cu, b200AprilTagsDetectBatchHost): only every row_step-th source row is staged by DMA; the other
  // rows are fetched on demand, in segments of (1 << seg_shift) pixels, where a quad needs them (k_decode.cu)
  int row_step;        // 0 = every row of the frame is present
  int seg_shift;       // log2 of the segment width in pixels (<= 64 segments per row)
};

struct FitParams {
  int tag_width;         // min_tag_width after decimation, >= 3
  int normal_border, reversed_border;
  int max_nmaxima;
  float max_line_fit_mse;
  float cos_critical_rad;
  float smooth[7];       // exp(-j*j/2) taps, j=-3..3, as float (host libm, same as the oracle)
  float quad_decimate;
  int refine_edges;
  double decode_sharpening;
  int max_hamming;
  int nfam;
  float fx, fy, cx, cy;
  float tagsize;
};

// Per-handle knobs, read once from the environment (B200AT_TUNE="key=value,key=value"; see capi.cu).  Every variant that did not
// win its measurement was deleted (profiles/r03_variants.md); what is left are two real trade-offs.
struct Tune {
  int ccl_tma;    // 1 (default) = CCL tiles staged in shared memory by TMA; 0 = staged by plain loads (the pipeline is then TMA-free)
  int qf_exact;   // 1 = k_quad.cu (one CTA per cluster, serial prefix sums: float corners bit-identical to the CPU oracle);
                  // 0 (default) = k_quad2.cu (sort / windowed moments / tail: same formulas, prefix sums associated differently)
  int qf_bucket_limit;  // bucket sort of the quad fit: a cluster with a bucket larger than this is sorted by the global-memory merge sort
                        // instead (default 1024; small values route every cluster through the fallback: used by the tests)
};

struct Workspace {
  Geo g;
  FitParams fp;
  Tune tune;
  DevFamily fams[kMaxFamilies];
  FrameDesc *frames;
  uint8_t *dec, *dec_tmp, *tmin, *tmax, *thr, *thr2;
  uint8_t *tth, *tlow;       // [B][ceil(Hd/4)][Wp/4] per-tile threshold value / low-contrast flag (tile size 4)
  uint32_t *lab, *csize;     // final labels (global representative per pixel); pixel counts at representatives
  uint32_t *lab0;            // CCL working labels: tile-local roots, then tile roots -> global roots (+ size flag in bit 31)
  uint2 *ccl_req;            // [B][tiles][192] cross-tile links (pixel, neighbour) found by the CCL tile kernel
  uint32_t *ccl_reqcnt;      // [B][tiles]
  uint32_t *roots, *nroots;  // [B][Hd][Wp] pixel indices of the CCL tile roots of a frame (a list: the first nroots[frame] entries), [B]
  unsigned long long *hkey;
  uint32_t *hcnt, *hoff, *hcur;
  ClusterRec *clusters;
  uint32_t *pts;
  unsigned long long *keys;
  LineFitPt *lfps;
  double *errs;  // 2 x pts_cap
  QuadRec *quads, *quads_refined;
  Cand *cands;
  uint32_t *cand_count;  // [B]
  b200AprilTagsDetection_t *out;
  uint32_t *out_count;   // [B]
  uint32_t *counters;    // [CNT_N]
  uint32_t *bin_idx;     // [kQuadBins][clu_cap] cluster indices per size bin
  // sparse host path: per-row segment bitmaps (bit s of word [frame][y] = pixels [s << seg_shift, (s+1) << seg_shift) of row y),
  // the host-mapped source frames the segments are fetched from, and the per-quad homographies handed from k_refine to k_decode_bits
  unsigned long long *need1, *need2;  // [B][H] rows needed by refine_edges / by the decode samples
  FrameDesc *src_frames;       // [B] device-accessible addresses of the caller's (pinned) host frames
  double *quad_H;              // [quad_cap][10]: H[0..8], valid flag
  // windowed quad fit (k_quad2.cu)
  uint32_t *qinfo;             // [clu_cap] 0 = rejected before the sort, else point count | reversed_border << 31
  uint32_t *qwbase;            // [clu_cap] first work item of the cluster (its chunks are consecutive)
  uint4 *qwork;                // [qwork_cap] self-contained work items of k_qf_window: (point offset, point count, chunk, cluster)
  uint32_t qwork_cap;
  double *qwtot;               // [qwork_cap][6] moment totals of the chunk
  uint32_t *qwnmax;            // [qwork_cap] local maxima found in the chunk
  const unsigned char *combos;  // per nm (4..kMaxNMaxima): all m0<m1<m2<m3 < nm in lexicographic order, uchar4 each
  int combo_off[18];     // combos for nm start at combo_off[nm], count combo_off[nm+1]-combo_off[nm]
  CUtensorMap thr_tmap;  // TMA descriptor of thr as a (Wp, Hd, B) u8 tensor, box 64 x 33 x 1 (CCL tile + halo)
  int use_tma;
  cudaStream_t aux[kQuadAux];
  cudaEvent_t ev_fork, ev_join[kQuadAux];
  // rectify / resize pre-stage: per-pixel source coordinates (shared by all frames: L2-resident), gray output, its frame table
  const float2 *rect_map;   // [H][W] or nullptr
  uint8_t *rect_img;        // [B][H][rect_pitch]
  FrameDesc *rect_frames;   // [B] -> rect_img
  int rect_pitch, rect_src_w, rect_src_h;
  uint8_t blur_k[32];
  int blur_ksz;
  int blur_sharpen;
};

__host__ __device__ inline int at_Wp(const Geo &g) { return (g.Wd + 15) & ~15; }  // internal row pitch
__host__ __device__ inline int at_twp(const Geo &g) { return g.tw > 0 ? g.tw : 1; }    // tile-array pitch

// ---- launchers (each returns the number of kernel launches it issued) ----
int launch_rectify(const Workspace &ws, int nframes, cudaStream_t s);     // raw frames (ws.frames) -> ws.rect_img
int launch_preprocess(const Workspace &ws, int nframes, cudaStream_t s);
int launch_threshold(const Workspace &ws, int nframes, cudaStream_t s);
int launch_ccl(const Workspace &ws, int nframes, cudaStream_t s);
int launch_cluster(const Workspace &ws, int nframes, cudaStream_t s);
int launch_quadfit(const Workspace &ws, int nframes, cudaStream_t s);           // dispatches on Tune::qf_exact
int launch_quadfit_windowed(const Workspace &ws, int nframes, cudaStream_t s);  // k_quad2.cu
void launch_bin_clusters(const Workspace &ws, int sms, cudaStream_t s);         // k_quad.cu (shared by both)
int launch_decode(const Workspace &ws, int nframes, cudaStream_t s);
int launch_sparse_fetch1(const Workspace &ws, int nframes, cudaStream_t s);  // sparse host path: mark + fetch for refine_edges
int launch_sparse_back(const Workspace &ws, int nframes, cudaStream_t s);    // ... refine, second fetch, decode
int launch_finalize(const Workspace &ws, int nframes, cudaStream_t s);

}  // namespace b200at
