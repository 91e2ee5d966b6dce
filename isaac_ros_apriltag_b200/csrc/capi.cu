// capi.cu -- the C ABI (include/b200_apriltags.h): workspace management, stage sequencing, host marshalling.
// Part 1 entry points are the three symbols the reference node binds
// (/root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:450-452, 491-493, 556).
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "detector.h"
#include "tag_families_data.inc"

using namespace b200at;

// cuAprilTagsID_t as the reference node (compiled against CUDA's 8-byte aligned float2) lays it out; every in-repo caller
// built WITHOUT CUDA headers must agree (include/b200_apriltags.h stand-in, capi.py TagID, apriltag_node_core.cpp)
static_assert(sizeof(cuAprilTagsID_t) == 88 && alignof(cuAprilTagsID_t) == 8, "cuAprilTagsID_t layout");
static_assert(offsetof(cuAprilTagsID_t, id) == 32 && offsetof(cuAprilTagsID_t, orientation) == 36 && offsetof(cuAprilTagsID_t, translation) == 72,
              "cuAprilTagsID_t layout");

namespace {

struct HostFamily {
  int nbits, ncodes, width_at_border, total_width, reversed_border;
  const unsigned char *bit_x, *bit_y;  // (built-in tables: all coordinates inside the border)
  const unsigned long long *codes;
};
// caller-supplied families (b200AprilTagsRegisterFamily), process wide
struct CustomFamily {
  bool used = false;
  int nbits = 0, ncodes = 0, width_at_border = 0, total_width = 0, reversed_border = 0;
  std::vector<int8_t> bit_x, bit_y;
  std::vector<unsigned long long> codes;
};
CustomFamily g_custom_families[B200AT_MAX_FAMILIES - B200AT_NUM_FAMILIES];
std::mutex g_custom_mutex;
const HostFamily kHostFamilies[B200AT_NUM_FAMILIES] = {
    {tag36h11_nbits, tag36h11_ncodes, tag36h11_width_at_border, tag36h11_total_width, 0, tag36h11_bit_x, tag36h11_bit_y, tag36h11_codes},
    {tag25h9_nbits, tag25h9_ncodes, tag25h9_width_at_border, tag25h9_total_width, 0, tag25h9_bit_x, tag25h9_bit_y, tag25h9_codes},
    {tag16h5_nbits, tag16h5_ncodes, tag16h5_width_at_border, tag16h5_total_width, 0, tag16h5_bit_x, tag16h5_bit_y, tag16h5_codes},
    {tag36h10_nbits, tag36h10_ncodes, tag36h10_width_at_border, tag36h10_total_width, 0, tag36h10_bit_x, tag36h10_bit_y, tag36h10_codes},
};

#define CK(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) {                                                                                  \
      fprintf(stderr, "[b200apriltags] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
              cudaGetErrorString(e_));                                                                        \
      return (e_ == cudaErrorMemoryAllocation) ? B200AT_ERR_NOMEM : B200AT_ERR_CUDA;                          \
    }                                                                                                         \
  } while (0)

// B200AT_TUNE="ccl_tma=0,qf_exact=1,qf_bucket_limit=4": the per-handle knobs (detector.h, struct Tune).  Unknown keys are reported and ignored.
Tune parse_tune() {
  Tune t;
  t.ccl_tma = 1;
  t.qf_exact = 0;
  t.qf_bucket_limit = 1024;
  const char *e = getenv("B200AT_TUNE");
  if (!e) return t;
  std::string str(e);
  size_t pos = 0;
  while (pos < str.size()) {
    size_t end = str.find(',', pos);
    if (end == std::string::npos) end = str.size();
    const std::string kv = str.substr(pos, end - pos);
    pos = end + 1;
    const size_t eq = kv.find('=');
    if (eq == std::string::npos) continue;
    const std::string k = kv.substr(0, eq);
    const int v = atoi(kv.c_str() + eq + 1);
    if (k == "ccl_tma") t.ccl_tma = v;
    else if (k == "qf_exact") t.qf_exact = v;
    else if (k == "qf_bucket_limit") t.qf_bucket_limit = v > 0 ? v : 1;
    else fprintf(stderr, "[b200apriltags] B200AT_TUNE: unknown key '%s'\n", k.c_str());
  }
  return t;
}

uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

// one host call in flight: its pinned tables and result buffers
struct HostCall {
  bool active = false, failed = false, sparse = false;
  uint32_t n = 0, nsub = 0;
  uint64_t seq = 0, dma_bytes = 0;
  int launches = 0;
  FrameDesc *frames_tab = nullptr, *src_tab = nullptr;   // [cap_frames] staged frames / device-mapped caller frames
  b200AprilTagsDetection_t *out = nullptr;               // [cap_frames][max_tags]
  uint32_t *out_count = nullptr, *counters = nullptr;    // [cap_frames], [cap_subs][kMaxChunks * CNT_N]
  size_t cap_frames = 0, cap_subs = 0;
  cudaEvent_t done = nullptr;
};
struct HostPendingBack {
  bool valid = false, last = false;
  int call = 0;
  uint64_t k = 0;
  uint32_t sub = 0, start = 0, len = 0;
};

struct cuAprilTagsHandle_st {
  Workspace ws;
  b200AprilTagsOptions_t opt;
  int device = 0;
  uint32_t max_batch = 1;
  std::vector<void *> dev_allocs;
  unsigned long long *dev_codes[B200AT_MAX_FAMILIES] = {};
  // Device-pointer entry points: up to TWO batches may be queued (b200AprilTagsEnqueueBatch twice before the first
  // b200AprilTagsCollectBatch, same stream), so that the host side of a call -- frame table, graph launch, copying the results
  // out -- overlaps the previous batch's kernels instead of leaving the GPU idle between calls.  Each slot has its own pinned
  // frame table / result buffers, its own cached CUDA graph (the graph bakes those host addresses in) and a completion event.
  struct DevSlot {
    FrameDesc *h_frames = nullptr;
    b200AprilTagsDetection_t *h_out = nullptr;
    uint32_t *h_out_count = nullptr;
    uint32_t *h_counters = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_n = 0;
    int graph_fast = -1, graph_enc = -1;
    cudaStream_t graph_stream = nullptr;
    cudaEvent_t done = nullptr;
    uint32_t n = 0;
    int launches = 0;
  } dslot[2];
  uint64_t dev_enq = 0, dev_col = 0;  // batches queued / collected (slot = counter & 1)
  // results of the batch collected last (point into its slot): cuAprilTagsDetect and b200AprilTagsReadBuffer read them
  b200AprilTagsDetection_t *h_out = nullptr;
  uint32_t *h_out_count = nullptr;
  uint32_t *h_counters = nullptr;
  // host entry points: staging slots, streams and events of the persistent sub-batch pipeline (see host_enqueue)
  uint8_t *d_stage = nullptr;
  size_t stage_pitch = 0;
  uint32_t stage_sub = 0;    // frames per staging slot
  uint32_t stage_slots = 0;
  cudaStream_t own_stream = nullptr;   // compute stream of the host path
  cudaStream_t copy_stream = nullptr;
  cudaStream_t fetch_stream = nullptr, tail_stream = nullptr;  // (high priority: few, latency-bound CTAs)
  cudaEvent_t ev_copied[3] = {nullptr, nullptr, nullptr}, ev_consumed[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_front[2] = {nullptr, nullptr}, ev_fetched[2] = {nullptr, nullptr}, ev_decoded[2] = {nullptr, nullptr},
              ev_tail[2] = {nullptr, nullptr}, ev_backdone[2] = {nullptr, nullptr};
  bool sparse_bufs = false;            // need1 / need2 / src_frames / quad_H allocated
  uint32_t sparse_skip = 0;            // calls left on the full copy after a sparse call that fetched nearly every row anyway
  float2 *rect_map_buf = nullptr;      // rectify / resize pre-stage: the map's device buffer (ws.rect_map points at it while enabled)
  HostCall calls[2];                   // up to two host calls in flight
  HostPendingBack pending_back;        // BACK of the last sub-batch queued so far (queued after the next FRONT, or at collect)
  uint64_t host_k = 0, host_seq = 0;   // global sub-batch / call counters
  bool host_pipe = false;              // shape of what is in flight
  uint32_t host_S = 0;
  // state of the batch in flight
  cudaStream_t cur_stream = nullptr;
  uint32_t cur_n = 0;
  bool in_flight = false;
  uint32_t last_status = 0;
  uint64_t last_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int launches = 0;
  bool timing = false;
  cudaEvent_t ev[B200AT_NUM_STAGES + 1] = {};
  float stage_ms[B200AT_NUM_STAGES] = {};
  float tag_dim = 0;
  // the second workspace view of the host path has its own side streams / events for the quad-fit fork / join
  cudaStream_t lane_aux[kQuadAux] = {};
  cudaEvent_t lane_fork = nullptr, lane_join[kQuadAux] = {};
  cudaEvent_t ev_in = nullptr;  // orders the handle's own stream after the caller's legacy default stream (resolve_sync_stream)
  // CUDA graph of one whole batch (all stage launches, fork/join of the quad-fit streams, D2H): replayed when the same
  // (n, stream, alignment class, encoding) comes again, e.g. the one-frame-at-a-time node path
  uint32_t plain_calls = 0;
  bool use_graph = true;
};

namespace {

template <typename T>
int dev_alloc(cuAprilTagsHandle_st *h, T **p, size_t count) {
  void *q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = 16;
  CK(cudaMalloc(&q, bytes));
  h->dev_allocs.push_back(q);
  *p = reinterpret_cast<T *>(q);
  return 0;
}

void destroy_handle(cuAprilTagsHandle_st *h) {
  if (!h) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->fetch_stream) cudaStreamSynchronize(h->fetch_stream);
  if (h->tail_stream) cudaStreamSynchronize(h->tail_stream);
  for (auto &d : h->dslot) {
    if (d.graph_exec) cudaGraphExecDestroy(d.graph_exec);
    if (d.done) cudaEventDestroy(d.done);
  }
  for (int i = 0; i < kQuadAux; i++) {
    if (h->lane_aux[i]) cudaStreamDestroy(h->lane_aux[i]);
    if (h->lane_join[i]) cudaEventDestroy(h->lane_join[i]);
  }
  if (h->lane_fork) cudaEventDestroy(h->lane_fork);
  if (h->ev_in) cudaEventDestroy(h->ev_in);
  for (void *p : h->dev_allocs) cudaFree(p);
  for (auto &d : h->dslot) {
    if (d.h_frames) cudaFreeHost(d.h_frames);
    if (d.h_out) cudaFreeHost(d.h_out);
    if (d.h_out_count) cudaFreeHost(d.h_out_count);
    if (d.h_counters) cudaFreeHost(d.h_counters);
  }
  for (auto &c : h->calls) {
    if (c.frames_tab) cudaFreeHost(c.frames_tab);
    if (c.src_tab) cudaFreeHost(c.src_tab);
    if (c.out) cudaFreeHost(c.out);
    if (c.out_count) cudaFreeHost(c.out_count);
    if (c.counters) cudaFreeHost(c.counters);
    if (c.done) cudaEventDestroy(c.done);
  }
  for (int i = 0; i < 3; i++) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
  }
  for (int i = 0; i < 2; i++) {
    if (h->ev_front[i]) cudaEventDestroy(h->ev_front[i]);
    if (h->ev_fetched[i]) cudaEventDestroy(h->ev_fetched[i]);
    if (h->ev_decoded[i]) cudaEventDestroy(h->ev_decoded[i]);
    if (h->ev_tail[i]) cudaEventDestroy(h->ev_tail[i]);
    if (h->ev_backdone[i]) cudaEventDestroy(h->ev_backdone[i]);
  }
  if (h->fetch_stream) cudaStreamDestroy(h->fetch_stream);
  if (h->tail_stream) cudaStreamDestroy(h->tail_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < kQuadAux; i++) {
    if (h->ws.aux[i]) cudaStreamDestroy(h->ws.aux[i]);
    if (h->ws.ev_join[i]) cudaEventDestroy(h->ws.ev_join[i]);
  }
  if (h->ws.ev_fork) cudaEventDestroy(h->ws.ev_fork);
  for (auto &e : h->ev)
    if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (prev >= 0) cudaSetDevice(prev);
  delete h;
}

int bpp_of(int enc) {
  switch (enc) {
    case B200AT_ENC_MONO8: return 1;
    case B200AT_ENC_RGB8:
    case B200AT_ENC_BGR8: return 3;
    case B200AT_ENC_RGBA8:
    case B200AT_ENC_BGRA8: return 4;
    default: return 0;
  }
}

void to_id_struct(const b200AprilTagsDetection_t &d, cuAprilTagsID_t *o) {
  // message corner order = reverse of AprilRobotics p[0..3] (verified on the POL golden values)
  for (int k = 0; k < 4; k++) {
    o->corners[k].x = (float)d.p[3 - k][0];
    o->corners[k].y = (float)d.p[3 - k][1];
  }
  o->id = (uint16_t)d.id;
  o->hamming_error = (uint8_t)d.hamming;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o->orientation[c * 3 + r] = (float)d.R[r * 3 + c];  // column major
  for (int k = 0; k < 3; k++) o->translation[k] = (float)d.t[k];
}

}  // namespace

extern "C" {

const char *b200AprilTagsVersion(void) { return "b200apriltags 0.1.0 (sm_100a)"; }

void b200AprilTagsDefaultOptions(b200AprilTagsOptions_t *o) {
  memset(o, 0, sizeof(*o));
  o->struct_size = sizeof(*o);
  o->family_mask = 1u << B200AT_FAM_36H11;
  o->max_batch = 1;
  o->max_tags = 64;
  o->tile_size = 4;
  o->quad_decimate = 2.0f;
  o->quad_sigma = 0.0f;
  o->refine_edges = 1;
  o->decode_sharpening = 0.25;
  o->min_white_black_diff = 5;
  o->max_nmaxima = 10;
  o->critical_rad = (float)(10 * M_PI / 180);
  o->max_line_fit_mse = 10.0f;
  o->max_hamming = 2;
  o->input_encoding = B200AT_ENC_BGR8;
  o->device = -1;
}

int b200AprilTagsRegisterFamily(int32_t slot, const b200AprilTagsFamilyDesc_t *d) {
  if (slot < B200AT_NUM_FAMILIES || slot >= B200AT_MAX_FAMILIES || !d || d->struct_size != sizeof(*d)) return B200AT_ERR_INVALID_ARG;
  if (d->nbits < 1 || d->nbits > (uint32_t)kMaxBits || d->ncodes < 1 || !d->bit_x || !d->bit_y || !d->codes) return B200AT_ERR_INVALID_ARG;
  if (d->width_at_border < 1 || d->total_width < d->width_at_border || d->total_width > 12 || ((d->total_width - d->width_at_border) & 1)) return B200AT_ERR_INVALID_ARG;
  const int lo = -(int)(d->total_width - d->width_at_border) / 2, hi = lo + (int)d->total_width - 1;
  for (uint32_t k = 0; k < d->nbits; k++)
    if (d->bit_x[k] < lo || d->bit_x[k] > hi || d->bit_y[k] < lo || d->bit_y[k] > hi) return B200AT_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(g_custom_mutex);
  CustomFamily &cf = g_custom_families[slot - B200AT_NUM_FAMILIES];
  cf.nbits = (int)d->nbits;
  cf.ncodes = (int)d->ncodes;
  cf.width_at_border = (int)d->width_at_border;
  cf.total_width = (int)d->total_width;
  cf.reversed_border = d->reversed_border ? 1 : 0;
  cf.bit_x.assign(d->bit_x, d->bit_x + d->nbits);
  cf.bit_y.assign(d->bit_y, d->bit_y + d->nbits);
  cf.codes.assign(d->codes, d->codes + d->ncodes);
  cf.used = true;
  return B200AT_OK;
}

int b200AprilTagsCreate(cuAprilTagsHandle *out, uint32_t W, uint32_t H, const cuAprilTagsCameraIntrinsics_t *cam, float tag_dim,
                        const b200AprilTagsOptions_t *opt_in) {
  if (!out) return B200AT_ERR_INVALID_ARG;
  *out = nullptr;
  b200AprilTagsOptions_t opt;
  b200AprilTagsDefaultOptions(&opt);
  if (opt_in) {
    if (opt_in->struct_size != sizeof(opt)) return B200AT_ERR_INVALID_ARG;
    opt = *opt_in;
  }
  if (W == 0 || H == 0 || W > 16382 || H > 16382) return B200AT_ERR_INVALID_ARG;
  if (opt.max_batch == 0 || opt.max_tags == 0 || opt.tile_size == 0) return B200AT_ERR_INVALID_ARG;
  if (opt.family_mask == 0 || (opt.family_mask >> B200AT_MAX_FAMILIES) != 0) return B200AT_ERR_UNSUPPORTED;
  {
    std::lock_guard<std::mutex> lk(g_custom_mutex);
    for (int i = B200AT_NUM_FAMILIES; i < B200AT_MAX_FAMILIES; i++)
      if ((opt.family_mask & (1u << i)) && !g_custom_families[i - B200AT_NUM_FAMILIES].used) return B200AT_ERR_UNSUPPORTED;
  }
  if (bpp_of(opt.input_encoding) == 0) return B200AT_ERR_INVALID_ARG;
  // integer decimation factors, plus AprilRobotics' 3->2 "1.5" special case
  float qd = opt.quad_decimate;
  const bool dec15 = (qd == 1.5f);
  if (!dec15 && (!(qd >= 1.0f) || qd != floorf(qd) || qd > 8.0f)) return B200AT_ERR_UNSUPPORTED;
  if (opt.max_nmaxima < 4 || opt.max_nmaxima > kMaxNMaxima) return B200AT_ERR_UNSUPPORTED;
  if (opt.max_hamming < 0 || opt.max_hamming > 3) return B200AT_ERR_UNSUPPORTED;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "[b200apriltags] no CUDA device: this library has no CPU fallback\n");
    return B200AT_ERR_NO_DEVICE;
  }
  cuAprilTagsHandle_st *h = new (std::nothrow) cuAprilTagsHandle_st();
  if (!h) return B200AT_ERR_NOMEM;
  int caller_device = -1;
  cudaGetDevice(&caller_device);
  struct RestoreDevice {  // the caller's current device is left as it was found
    int d;
    ~RestoreDevice() {
      if (d >= 0) cudaSetDevice(d);
    }
  } restore_device{caller_device};
  if (opt.device >= 0) {
    if (cudaSetDevice(opt.device) != cudaSuccess) {
      delete h;
      return B200AT_ERR_INVALID_ARG;
    }
  }
  cudaGetDevice(&h->device);
  h->opt = opt;
  h->max_batch = opt.max_batch;
  h->tag_dim = tag_dim;
  Workspace &ws = h->ws;
  memset(&ws, 0, sizeof(ws));
  Geo &g = ws.g;
  const int f = dec15 ? 0 : (int)qd;  // Geo::f == 0 marks the 1.5 mode
  g.W = (int)W;
  g.H = (int)H;
  g.f = f;
  if (dec15) {
    g.Wd = (int)W / 3 * 2;
    g.Hd = (int)H / 3 * 2;
  } else {
    g.Wd = f > 1 ? 1 + ((int)W - 1) / f : (int)W;
    g.Hd = f > 1 ? 1 + ((int)H - 1) / f : (int)H;
  }
  g.ts = (int)opt.tile_size;
  g.tw = g.Wd / g.ts;
  g.th = g.Hd / g.ts;
  g.enc = opt.input_encoding;
  g.bpp = bpp_of(opt.input_encoding);
  g.min_wb_diff = opt.min_white_black_diff;
  g.fast_align = 0;
  if (g.Wd > 8191 || g.Hd > 8191 || g.tw < 1 || g.th < 1) {
    delete h;
    return B200AT_ERR_UNSUPPORTED;
  }
  const uint32_t B = opt.max_batch;
  const size_t Wp = (size_t)at_Wp(g), Pp = Wp * g.Hd, Pd = (size_t)g.Wd * g.Hd;
  // one slot per (black component, white component) pair with both >= 25 px: measured ~1 k per 1080p frame even on the
  // all-texture bench workload; Pd/8 slots leave room for ~40 k pairs (a 5x5-cell checkerboard) before ST_HASH_FULL
  g.hcap = opt.hash_slots_per_frame ? next_pow2(std::max<uint32_t>(opt.hash_slots_per_frame, 64u)) : next_pow2((uint32_t)std::max<size_t>(Pd / 8, 4096));
  const size_t ppf = opt.points_per_frame ? opt.points_per_frame : Pd;
  const size_t cpf = opt.clusters_per_frame ? opt.clusters_per_frame : std::max<size_t>(Pd / 64, 1024);
  const size_t qpf = opt.quads_per_frame ? opt.quads_per_frame : 1024;
  if (ppf * B > 0xfffffff0ull || cpf * B > 0xfffffff0ull) {
    delete h;
    return B200AT_ERR_INVALID_ARG;
  }
  g.pts_cap = (uint32_t)(ppf * B);
  g.clu_cap = (uint32_t)(cpf * B);
  g.quad_cap = (uint32_t)(qpf * B);
  g.cand_cap = std::min<uint32_t>(1024u, std::max<uint32_t>(256u, 2 * opt.max_tags));  // decoded candidates per frame before reconcile (k_final.cu MAXC)
  g.max_tags = opt.max_tags;
  g.max_cluster_pts = (uint32_t)(2 * (2 * g.Wd + 2 * g.Hd));

  FitParams &fp = ws.fp;
  int min_tag_width = 1000000, nfam = 0;
  fp.normal_border = fp.reversed_border = 0;
  int rc = 0;
  for (int i = 0; i < B200AT_MAX_FAMILIES && rc == 0; i++) {
    if (!(opt.family_mask & (1u << i))) continue;
    DevFamily &df = ws.fams[nfam++];
    const unsigned long long *host_codes = nullptr;
    if (i < B200AT_NUM_FAMILIES) {
      const HostFamily &hf = kHostFamilies[i];
      df.nbits = hf.nbits;
      df.ncodes = hf.ncodes;
      df.width_at_border = hf.width_at_border;
      df.total_width = hf.total_width;
      df.reversed_border = hf.reversed_border;
      for (int k = 0; k < hf.nbits; k++) {
        df.bit_x[k] = (int8_t)hf.bit_x[k];
        df.bit_y[k] = (int8_t)hf.bit_y[k];
      }
      host_codes = hf.codes;
    } else {
      std::lock_guard<std::mutex> lk(g_custom_mutex);
      const CustomFamily &cf = g_custom_families[i - B200AT_NUM_FAMILIES];
      df.nbits = cf.nbits;
      df.ncodes = cf.ncodes;
      df.width_at_border = cf.width_at_border;
      df.total_width = cf.total_width;
      df.reversed_border = cf.reversed_border;
      for (int k = 0; k < cf.nbits; k++) {
        df.bit_x[k] = cf.bit_x[k];
        df.bit_y[k] = cf.bit_y[k];
      }
      host_codes = cf.codes.data();
    }
    df.index = i;
    rc = dev_alloc(h, &h->dev_codes[i], (size_t)df.ncodes);
    if (rc == 0 && cudaMemcpy(h->dev_codes[i], host_codes, sizeof(unsigned long long) * df.ncodes, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = B200AT_ERR_CUDA;
    df.codes = h->dev_codes[i];
    if (df.width_at_border < min_tag_width) min_tag_width = df.width_at_border;
    fp.normal_border |= !df.reversed_border;
    fp.reversed_border |= df.reversed_border;
  }
  if (qd > 1) min_tag_width = (int)(min_tag_width / qd);
  if (min_tag_width < 3) min_tag_width = 3;
  fp.tag_width = min_tag_width;
  fp.max_nmaxima = opt.max_nmaxima;
  fp.max_line_fit_mse = opt.max_line_fit_mse;
  fp.cos_critical_rad = (float)cos((double)opt.critical_rad);
  {
    const double sigma = 1;
    for (int i = 0; i < 7; i++) {
      int j = i - 3;
      fp.smooth[i] = (float)exp(-j * j / (2 * sigma * sigma));
    }
  }
  fp.quad_decimate = qd;
  fp.refine_edges = opt.refine_edges;
  fp.decode_sharpening = opt.decode_sharpening;
  fp.max_hamming = opt.max_hamming;
  fp.nfam = nfam;
  if (cam) {
    fp.fx = cam->fx;
    fp.fy = cam->fy;
    fp.cx = cam->cx;
    fp.cy = cam->cy;
  } else {
    fp.fx = fp.fy = 1;
    fp.cx = fp.cy = 0;
  }
  fp.tagsize = tag_dim;
  // blur taps (apriltag_detector_detect's quad_sigma block + image_u8_gaussian_blur)
  ws.blur_ksz = 0;
  ws.blur_sharpen = 0;
  if (opt.quad_sigma != 0) {
    float sigma = fabsf(opt.quad_sigma);
    int ksz = (int)(4 * sigma);
    if ((ksz & 1) == 0) ksz++;
    if (ksz > 31) ksz = 31;
    if (ksz > 1) {
      std::vector<double> dk(ksz);
      double acc = 0;
      for (int i = 0; i < ksz; i++) {
        int x = -ksz / 2 + i;
        dk[i] = exp(-.5 * ((double)x / sigma) * ((double)x / sigma));
        acc += dk[i];
      }
      for (int i = 0; i < ksz; i++) ws.blur_k[i] = (uint8_t)(dk[i] / acc * 255);
      ws.blur_ksz = ksz;
      ws.blur_sharpen = opt.quad_sigma < 0 ? 1 : 0;
    }
  }

#define ALLOC(ptr, count)                  \
  if (rc == 0) rc = dev_alloc(h, &(ptr), (size_t)(count))
  ALLOC(ws.frames, B);
  ALLOC(ws.dec, B * Pp);
  ALLOC(ws.dec_tmp, ws.blur_ksz > 1 ? B * Pp : 16);
  ALLOC(ws.tmin, (size_t)B * g.th * at_twp(g));
  ALLOC(ws.tmax, (size_t)B * g.th * at_twp(g));
  ALLOC(ws.tth, (size_t)B * ((g.Hd + 3) / 4) * (Wp / 4));
  ALLOC(ws.tlow, (size_t)B * ((g.Hd + 3) / 4) * (Wp / 4));
  ALLOC(ws.thr, B * Pp);
  ALLOC(ws.thr2, B * Pp);
  ALLOC(ws.lab, B * Pp);
  ALLOC(ws.lab0, B * Pp);
  ALLOC(ws.csize, B * Pp);
  ALLOC(ws.roots, B * Pp);
  {
    const size_t ntl = (size_t)((g.Wd + 31) / 32) * ((g.Hd + 31) / 32);
    ALLOC(ws.ccl_req, (size_t)B * ntl * 192);
    ALLOC(ws.ccl_reqcnt, (size_t)B * ntl);
  }
  ALLOC(ws.nroots, B);
  ALLOC(ws.hkey, (size_t)B * g.hcap);
  ALLOC(ws.hcnt, (size_t)B * g.hcap);
  ALLOC(ws.hoff, (size_t)B * g.hcap);
  ALLOC(ws.hcur, (size_t)B * g.hcap);
  ALLOC(ws.clusters, g.clu_cap);
  ALLOC(ws.pts, g.pts_cap);
  ALLOC(ws.keys, g.pts_cap);
  ALLOC(ws.lfps, g.pts_cap);
  ALLOC(ws.errs, (size_t)2 * g.pts_cap);
  ALLOC(ws.quads, g.quad_cap);
  ALLOC(ws.quads_refined, g.quad_cap);
  ALLOC(ws.cands, (size_t)B * g.cand_cap);
  ALLOC(ws.cand_count, B);
  ALLOC(ws.out, (size_t)B * g.max_tags);
  ALLOC(ws.out_count, B);
  ALLOC(ws.counters, (size_t)kMaxChunks * CNT_N);
  ALLOC(ws.bin_idx, (size_t)kQuadBins * g.clu_cap);
  ws.tune = parse_tune();
  ALLOC(ws.quad_H, (size_t)g.quad_cap * 10);
  if (!ws.tune.qf_exact) {
    ws.qwork_cap = (uint32_t)std::min<size_t>((size_t)g.pts_cap / kQfChunkMax + g.clu_cap, 0xfffffff0ull);
    ALLOC(ws.qinfo, g.clu_cap);
    ALLOC(ws.qwbase, g.clu_cap);
    ALLOC(ws.qwork, ws.qwork_cap);
    ALLOC(ws.qwtot, (size_t)ws.qwork_cap * 6);
    ALLOC(ws.qwnmax, ws.qwork_cap);
  }
  {
    // combination tables: for every nm, all m0<m1<m2<m3<nm in lexicographic order (the serial loops' visiting order)
    std::vector<unsigned char> tab;
    for (int nm = 0; nm <= 17; nm++) {
      ws.combo_off[nm] = (int)(tab.size() / 4);
      if (nm < 4 || nm > kMaxNMaxima) continue;
      for (int a = 0; a < nm - 3; a++)
        for (int b = a + 1; b < nm - 2; b++)
          for (int c = b + 1; c < nm - 1; c++)
            for (int d = c + 1; d < nm; d++) {
              tab.push_back((unsigned char)a);
              tab.push_back((unsigned char)b);
              tab.push_back((unsigned char)c);
              tab.push_back((unsigned char)d);
            }
    }
    unsigned char *dcomb = nullptr;
    ALLOC(dcomb, tab.size());
    if (rc == 0 && cudaMemcpy(dcomb, tab.data(), tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = B200AT_ERR_CUDA;
    ws.combos = dcomb;
  }
  // TMA descriptor for the CCL tile staging: thr viewed as a (Wp, Hd, B) u8 tensor, box = 64 x 33 x 1 (the box start
  // must be 16-byte aligned in global memory, so it begins 16 pixels left of the tile: x0-16 .. x0+47).
  ws.use_tma = 0;
  if (rc == 0) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
        qres == cudaDriverEntryPointSuccess && getenv("B200AT_NO_TMA") == nullptr) {
      const cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)g.Hd, (cuuint64_t)B};
      const cuuint64_t strides[2] = {(cuuint64_t)Wp, (cuuint64_t)Wp * g.Hd};  // bytes, dims 1..2
      const cuuint32_t box[3] = {64, 33, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      CUresult cr = ((EncodeFn)fn)(&ws.thr_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ws.thr, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr == CUDA_SUCCESS) ws.use_tma = 1;
    } else {
      cudaGetLastError();
    }
  }
  for (int i = 0; i < kQuadAux && rc == 0; i++) {
    if (cudaStreamCreateWithFlags(&ws.aux[i], cudaStreamNonBlocking) != cudaSuccess) rc = B200AT_ERR_CUDA;
    if (rc == 0 && cudaEventCreateWithFlags(&ws.ev_join[i], cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
  }
  if (rc == 0 && cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
#undef ALLOC
  for (auto &d : h->dslot) {
    if (rc == 0 && cudaMallocHost(&d.h_frames, sizeof(FrameDesc) * B) != cudaSuccess) rc = B200AT_ERR_NOMEM;
    if (rc == 0 && cudaMallocHost(&d.h_out, sizeof(b200AprilTagsDetection_t) * B * g.max_tags) != cudaSuccess) rc = B200AT_ERR_NOMEM;
    if (rc == 0 && cudaMallocHost(&d.h_out_count, sizeof(uint32_t) * B) != cudaSuccess) rc = B200AT_ERR_NOMEM;
    if (rc == 0 && cudaMallocHost(&d.h_counters, sizeof(uint32_t) * CNT_N * kMaxChunks) != cudaSuccess) rc = B200AT_ERR_NOMEM;
    if (rc == 0) memset(d.h_counters, 0, sizeof(uint32_t) * CNT_N * kMaxChunks);
    if (rc == 0 && cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
  }
  h->h_out = h->dslot[0].h_out;
  h->h_out_count = h->dslot[0].h_out_count;
  h->h_counters = h->dslot[0].h_counters;
  if (rc == 0 && cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) rc = B200AT_ERR_CUDA;
  for (int i = 0; i < kQuadAux && rc == 0; i++) {
    if (cudaStreamCreateWithFlags(&h->lane_aux[i], cudaStreamNonBlocking) != cudaSuccess) rc = B200AT_ERR_CUDA;
    if (rc == 0 && cudaEventCreateWithFlags(&h->lane_join[i], cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
  }
  if (rc == 0 && cudaEventCreateWithFlags(&h->lane_fork, cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
  if (rc == 0 && cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming) != cudaSuccess) rc = B200AT_ERR_CUDA;
  for (int i = 0; i <= B200AT_NUM_STAGES && rc == 0; i++)
    if (cudaEventCreate(&h->ev[i]) != cudaSuccess) rc = B200AT_ERR_CUDA;
  if (rc != 0) {
    destroy_handle(h);
    return rc;
  }
  *out = h;
  return B200AT_OK;
}

int nvCreateAprilTagsDetector(cuAprilTagsHandle *hApriltags, const uint32_t img_width, const uint32_t img_height,
                              const uint32_t tile_size, const cuAprilTagsFamily tag_family,
                              const cuAprilTagsCameraIntrinsics_t *cam, float tag_dim) {
  if (tag_family != NVAT_TAG36H11) return B200AT_ERR_UNSUPPORTED;  // the reference's cuAprilTags path: tag36h11 only
  b200AprilTagsOptions_t opt;
  b200AprilTagsDefaultOptions(&opt);
  opt.tile_size = tile_size;
  // the create call carries no max_tags (the node passes its parameter to cuAprilTagsDetect, apriltag_node.cpp:490-493): size the
  // handle so that the caller's max_tags is the only truncation
  opt.max_tags = 1024;
  return b200AprilTagsCreate(hApriltags, img_width, img_height, cam, tag_dim, &opt);
}

int cuAprilTagsDestroy(cuAprilTagsHandle h) {
  if (!h) return B200AT_ERR_INVALID_ARG;
  destroy_handle(h);
  return B200AT_OK;
}

int b200AprilTagsSetInputEncoding(cuAprilTagsHandle h, int32_t enc) {
  if (!h || bpp_of(enc) == 0) return B200AT_ERR_INVALID_ARG;
  h->ws.g.enc = enc;
  h->ws.g.bpp = bpp_of(enc);  // (a cached graph is keyed on the encoding and is re-captured when it changes)
  h->opt.input_encoding = enc;
  return B200AT_OK;
}

int b200AprilTagsSetRectification(cuAprilTagsHandle h, const b200AprilTagsRectify_t *r) {
  if (!h || h->in_flight || h->calls[0].active || h->calls[1].active) return B200AT_ERR_INVALID_ARG;
  if (r && (r->struct_size != sizeof(*r) || r->src_width == 0 || r->src_height == 0 || r->src_width > 16382 || r->src_height > 16382))
    return B200AT_ERR_INVALID_ARG;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto &d : h->dslot)
    if (d.graph_exec) {  // the cached graphs were captured without / with another pre-stage
      cudaGraphExecDestroy(d.graph_exec);
      d.graph_exec = nullptr;
    }
  Workspace &ws = h->ws;
  const Geo &g = ws.g;
  int rc = B200AT_OK;
  if (!r) {
    ws.rect_map = nullptr;  // (buffers stay allocated for the next SetRectification)
  } else {
    const size_t npx = (size_t)g.W * g.H;
    if (!h->rect_map_buf) {
      ws.rect_pitch = (g.W + 15) & ~15;
      rc = dev_alloc(h, &h->rect_map_buf, npx);
      if (rc == 0) rc = dev_alloc(h, &ws.rect_img, (size_t)h->max_batch * g.H * ws.rect_pitch);
      if (rc == 0) rc = dev_alloc(h, &ws.rect_frames, h->max_batch);
      if (rc == 0) {
        std::vector<FrameDesc> tab(h->max_batch);
        for (uint32_t i = 0; i < h->max_batch; i++) {
          tab[i].ptr = ws.rect_img + (size_t)i * g.H * ws.rect_pitch;
          tab[i].pitch = (unsigned long long)ws.rect_pitch;
        }
        if (cudaMemcpy(ws.rect_frames, tab.data(), sizeof(FrameDesc) * tab.size(), cudaMemcpyHostToDevice) != cudaSuccess) rc = B200AT_ERR_CUDA;
      }
    }
    if (rc == 0) {
      // OpenCV initUndistortRectifyMap in double, evaluated directly per pixel (oracle/rectify.py restates exactly this sequence)
      double A[9], iR[9];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i * 3 + j] = r->P[i * 3 + 0] * r->R[0 * 3 + j] + r->P[i * 3 + 1] * r->R[1 * 3 + j] + r->P[i * 3 + 2] * r->R[2 * 3 + j];
      const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
      if (det == 0) rc = B200AT_ERR_INVALID_ARG;
      const double id = 1.0 / det;
      iR[0] = (A[4] * A[8] - A[5] * A[7]) * id;
      iR[1] = (A[2] * A[7] - A[1] * A[8]) * id;
      iR[2] = (A[1] * A[5] - A[2] * A[4]) * id;
      iR[3] = (A[5] * A[6] - A[3] * A[8]) * id;
      iR[4] = (A[0] * A[8] - A[2] * A[6]) * id;
      iR[5] = (A[2] * A[3] - A[0] * A[5]) * id;
      iR[6] = (A[3] * A[7] - A[4] * A[6]) * id;
      iR[7] = (A[1] * A[6] - A[0] * A[7]) * id;
      iR[8] = (A[0] * A[4] - A[1] * A[3]) * id;
      const double fx = r->K[0], fy = r->K[4], u0 = r->K[2], v0 = r->K[5];
      const double k1 = r->D[0], k2 = r->D[1], p1 = r->D[2], p2 = r->D[3], k3 = r->D[4], k4 = r->D[5], k5 = r->D[6], k6 = r->D[7];
      std::vector<float2> map(rc == 0 ? npx : 0);
      for (int v = 0; v < g.H && rc == 0; v++) {
        for (int u = 0; u < g.W; u++) {
          const double X = u * iR[0] + v * iR[1] + iR[2], Y = u * iR[3] + v * iR[4] + iR[5], Wq = u * iR[6] + v * iR[7] + iR[8];
          const double w = 1.0 / Wq, x = X * w, y = Y * w;
          const double x2 = x * x, y2 = y * y, r2 = x2 + y2, xy2 = 2 * x * y;
          const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
          const double xd = x * kr + p1 * xy2 + p2 * (r2 + 2 * x2);
          const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * xy2;
          float2 m;
          m.x = (float)(fx * xd + u0);
          m.y = (float)(fy * yd + v0);
          map[(size_t)v * g.W + u] = m;
        }
      }
      if (rc == 0 && cudaMemcpy(h->rect_map_buf, map.data(), sizeof(float2) * npx, cudaMemcpyHostToDevice) != cudaSuccess) rc = B200AT_ERR_CUDA;
      if (rc == 0) {
        ws.rect_map = h->rect_map_buf;
        ws.rect_src_w = (int)r->src_width;
        ws.rect_src_h = (int)r->src_height;
      }
    }
  }
  if (prev != h->device) cudaSetDevice(prev);
  return rc;
}

int b200AprilTagsEnableStageTiming(cuAprilTagsHandle h, int enable) {
  if (!h) return B200AT_ERR_INVALID_ARG;
  h->timing = enable != 0;
  return B200AT_OK;
}

// Launch every stage of one batch on `stream` and queue the D2H of its results.  `hf` is a pinned frame-table slice
// that must stay untouched until the stream has consumed it; results land in the pinned arrays `out`, `cnt`, `ctr`.
// host side of a batch: validate the frames, fill the pinned frame table, classify the alignment
static int fill_frame_table(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n, FrameDesc *hf, int *fast_out) {
  const Geo &g = h->ws.g;
  int fast = 1;
  const size_t min_pitch = (size_t)(h->ws.rect_map ? h->ws.rect_src_w : g.W) * g.bpp;  // (with a rectification set the frames are the raw ones)
  for (uint32_t i = 0; i < n; i++) {
    if (!frames[i].ptr || frames[i].pitch < min_pitch) return B200AT_ERR_INVALID_ARG;
    hf[i].ptr = (const uint8_t *)frames[i].ptr;
    hf[i].pitch = frames[i].pitch;
    if (((uintptr_t)frames[i].ptr & 15) || (frames[i].pitch & 15)) fast = 0;
  }
  *fast_out = fast;
  return B200AT_OK;
}

// A view of the workspace for frames [f0, f0+..) processed as chunk c of nch: per-frame buffers are offset by f0 frames,
// the pools (points, clusters, quads) and the counters are split evenly between the chunks.
static Workspace make_view(cuAprilTagsHandle h, int f0, int c, int nch) {
  Workspace v = h->ws;
  const Geo &g = h->ws.g;
  const size_t Pp = (size_t)at_Wp(g) * g.Hd;
  v.frames += f0;
  v.dec += (size_t)f0 * Pp;
  v.dec_tmp += (h->ws.blur_ksz > 1) ? (size_t)f0 * Pp : 0;
  v.thr += (size_t)f0 * Pp;
  v.thr2 += (size_t)f0 * Pp;
  v.tmin += (size_t)f0 * g.th * at_twp(g);
  v.tmax += (size_t)f0 * g.th * at_twp(g);
  v.tth += (size_t)f0 * ((g.Hd + 3) / 4) * (at_Wp(g) / 4);
  v.tlow += (size_t)f0 * ((g.Hd + 3) / 4) * (at_Wp(g) / 4);
  v.lab += (size_t)f0 * Pp;
  v.lab0 += (size_t)f0 * Pp;
  v.csize += (size_t)f0 * Pp;
  v.roots += (size_t)f0 * Pp;
  {
    const size_t ntl = (size_t)((g.Wd + 31) / 32) * ((g.Hd + 31) / 32);
    v.ccl_req += (size_t)f0 * ntl * 192;
    v.ccl_reqcnt += (size_t)f0 * ntl;
  }
  v.nroots += f0;
  v.hkey += (size_t)f0 * g.hcap;
  v.hcnt += (size_t)f0 * g.hcap;
  v.hoff += (size_t)f0 * g.hcap;
  v.hcur += (size_t)f0 * g.hcap;
  const uint32_t pc = g.pts_cap / nch, cc = g.clu_cap / nch, qc = g.quad_cap / nch;
  v.g.pts_cap = pc;
  v.g.clu_cap = cc;
  v.g.quad_cap = qc;
  v.pts += (size_t)c * pc;
  v.keys += (size_t)c * pc;
  v.lfps += (size_t)c * pc;
  v.errs += (size_t)2 * c * pc;
  v.clusters += (size_t)c * cc;
  v.bin_idx += (size_t)c * kQuadBins * cc;
  if (v.qinfo) {
    const uint32_t wc = h->ws.qwork_cap / nch;
    v.qinfo += (size_t)c * cc;
    v.qwbase += (size_t)c * cc;
    v.qwork += (size_t)c * wc;
    v.qwtot += (size_t)c * wc * 6;
    v.qwnmax += (size_t)c * wc;
    v.qwork_cap = wc;
  }
  v.quads += (size_t)c * qc;
  v.quads_refined += (size_t)c * qc;
  v.cands += (size_t)f0 * g.cand_cap;
  v.cand_count += f0;
  v.out += (size_t)f0 * g.max_tags;
  v.out_count += f0;
  v.counters += (size_t)c * CNT_N;
  if (v.need1) {
    v.need1 += (size_t)f0 * g.H;
    v.need2 += (size_t)f0 * g.H;
    v.src_frames += f0;
  }
  if (v.quad_H) v.quad_H += (size_t)c * qc * 10;
  v.g.tma_frame0 = f0;
  if (c & 1) {  // lane 1 has its own side streams / events
    for (int i = 0; i < kQuadAux; i++) {
      v.aux[i] = h->lane_aux[i];
      v.ev_join[i] = h->lane_join[i];
    }
    v.ev_fork = h->lane_fork;
  }
  return v;
}

static void sum_counters(const uint32_t *c, uint32_t *out) {
  for (int k = 0; k < CNT_N; k++) out[k] = 0;
  for (int ch = 0; ch < kMaxChunks; ch++)
    for (int k = 0; k < CNT_N; k++) {
      if (k == CNT_STATUS)
        out[k] |= c[ch * CNT_N + k];
      else
        out[k] += c[ch * CNT_N + k];
    }
}

// Launch every stage of one batch and queue the D2H of its results.  `hf` is a pinned frame-table slice that must stay
// untouched until the stream has consumed it; results land in the pinned arrays `out`, `cnt`, `ctr` (kMaxChunks*CNT_N).
static int enqueue_core(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n, cudaStream_t stream, FrameDesc *hf,
                        b200AprilTagsDetection_t *out, uint32_t *cnt, uint32_t *ctr, bool timing, int *launches_out,
                        bool table_filled = false, const FrameDesc *sparse_src = nullptr) {
  Workspace &ws = h->ws;
  Geo &g = ws.g;
  // sparse host path: `frames` are staged copies that hold only every f-th row; `sparse_src` (pinned) lists the
  // device-mapped host frames the missing rows are fetched from where a quad needs them (k_decode.cu)
  struct SparseScope {
    Geo &g;
    ~SparseScope() { g.row_step = 0; }
  } sparse_scope{g};
  g.row_step = 0;
  if (sparse_src) {
    g.row_step = g.f;
    g.seg_shift = 5;
    while (((g.W + (1 << g.seg_shift) - 1) >> g.seg_shift) > 64) g.seg_shift++;
  }
  if (!table_filled) {
    int fast = 1;
    int rcf = fill_frame_table(h, frames, n, hf, &fast);
    if (rcf != B200AT_OK) return rcf;
    g.fast_align = fast;
  }
  int launches = 0;
  cudaError_t e = cudaMemcpyAsync(ws.frames, hf, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws.counters, 0, sizeof(uint32_t) * CNT_N * kMaxChunks, stream);
  if (sparse_src) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(ws.src_frames, sparse_src, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.need1, 0, sizeof(unsigned long long) * (size_t)n * g.H, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ws.need2, 0, sizeof(unsigned long long) * (size_t)n * g.H, stream);
  }
  const bool tm = timing;
  g.tma_frame0 = 0;
#define STAMP(i) \
  if (tm) cudaEventRecord(h->ev[i], stream)
  STAMP(0);
  // rectify / resize pre-stage: raw frames -> internal gray frames; everything after it sees mono8 frames of the handle's size
  struct RectScope {
    Workspace &w;
    FrameDesc *frames;
    int enc, bpp, fast;
    bool on;
    ~RectScope() {
      if (!on) return;
      w.frames = frames;
      w.g.enc = enc;
      w.g.bpp = bpp;
      w.g.fast_align = fast;
    }
  } rect_scope{ws, ws.frames, g.enc, g.bpp, g.fast_align, ws.rect_map != nullptr};
  if (rect_scope.on) {
    launches += launch_rectify(ws, (int)n, stream);
    ws.frames = ws.rect_frames;
    g.enc = B200AT_ENC_MONO8;
    g.bpp = 1;
    g.fast_align = 1;
  }
  launches += launch_preprocess(ws, (int)n, stream);
  STAMP(1);
  launches += launch_threshold(ws, (int)n, stream);
  STAMP(2);
  launches += launch_ccl(ws, (int)n, stream);
  STAMP(3);
  launches += launch_cluster(ws, (int)n, stream);
  STAMP(4);
  launches += launch_quadfit(ws, (int)n, stream);
  STAMP(5);
  launches += launch_decode(ws, (int)n, stream);
  STAMP(6);
  launches += launch_finalize(ws, (int)n, stream);
  STAMP(7);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(out, ws.out, sizeof(b200AprilTagsDetection_t) * (size_t)n * g.max_tags, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, ws.out_count, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(ctr, ws.counters, sizeof(uint32_t) * CNT_N * kMaxChunks, cudaMemcpyDeviceToHost, stream);
  STAMP(8);
#undef STAMP
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200apriltags] enqueue failed: %s\n", cudaGetErrorString(e));
    return B200AT_ERR_CUDA;
  }
  if (launches_out) *launches_out = launches;
  return B200AT_OK;
}

// One sub-batch of the host path on a VIEW of the workspace (frames [f0, f0 + n), pool slice c of 2), so that two sub-batches
// can be in flight on two streams: the latency-bound tails of one (large-cluster quad fits, decode, pose: a few busy SMs)
// overlap the dense stages of the next.  Everything the launch chain writes is inside the view.
enum { VIEW_ALL = 0, VIEW_FRONT = 1, VIEW_FETCH = 2, VIEW_BACK = 3 };
static void view_sparse_geo(Geo &g, bool sparse) {
  g.row_step = 0;
  if (sparse) {
    g.row_step = g.f;
    g.seg_shift = 5;
    while (((g.W + (1 << g.seg_shift) - 1) >> g.seg_shift) > 64) g.seg_shift++;
  }
}
// Pipelined sparse host path: the same chain cut in three.  FRONT = table upload + preprocess .. quad fit (compute stream);
// FETCH = mark + fetch of the rows refine_edges needs (fetch stream, overlaps the next sub-batch's FRONT); BACK = refine,
// second fetch, decode, reconcile, pose, D2H (compute stream, after FETCH).
static int enqueue_view_part(cuAprilTagsHandle h, Workspace v, int part, uint32_t n, cudaStream_t stream, b200AprilTagsDetection_t *out,
                             uint32_t *cnt, uint32_t *ctr, int *launches_out, cudaStream_t tail = nullptr, cudaEvent_t ev_decoded = nullptr) {
  (void)h;
  Geo &g = v.g;
  view_sparse_geo(g, true);
  int launches = 0;
  cudaError_t e = cudaSuccess;
  if (part == VIEW_FETCH) {
    launches += launch_sparse_fetch1(v, (int)n, stream);
  } else {  // VIEW_BACK: decode on the compute stream; reconcile, pose and the D2H of the results on the tail stream
    launches += launch_sparse_back(v, (int)n, stream);
    cudaStream_t ts = tail ? tail : stream;
    if (tail) {
      e = cudaEventRecord(ev_decoded, stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(tail, ev_decoded, 0);
    }
    launches += launch_finalize(v, (int)n, ts);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, v.out, sizeof(b200AprilTagsDetection_t) * (size_t)n * g.max_tags, cudaMemcpyDeviceToHost, ts);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, v.out_count, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ts);
    memset(ctr, 0, sizeof(uint32_t) * CNT_N * kMaxChunks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctr, v.counters, sizeof(uint32_t) * CNT_N, cudaMemcpyDeviceToHost, ts);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200apriltags] enqueue failed: %s\n", cudaGetErrorString(e));
    return B200AT_ERR_CUDA;
  }
  if (launches_out) *launches_out = launches;
  return B200AT_OK;
}

static int enqueue_view(cuAprilTagsHandle h, Workspace v, const b200AprilTagsFrame_t *frames, uint32_t n, cudaStream_t stream,
                        FrameDesc *hf, b200AprilTagsDetection_t *out, uint32_t *cnt, uint32_t *ctr, int *launches_out,
                        const FrameDesc *sparse_src, int part = VIEW_ALL) {
  Geo &g = v.g;
  int fast = 1;
  int rcf = fill_frame_table(h, frames, n, hf, &fast);
  if (rcf != B200AT_OK) return rcf;
  g.fast_align = fast;
  view_sparse_geo(g, sparse_src != nullptr);
  cudaError_t e = cudaMemcpyAsync(v.frames, hf, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(v.counters, 0, sizeof(uint32_t) * CNT_N, stream);
  if (sparse_src) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(v.src_frames, sparse_src, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(v.need1, 0, sizeof(unsigned long long) * (size_t)n * g.H, stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(v.need2, 0, sizeof(unsigned long long) * (size_t)n * g.H, stream);
  }
  int launches = 0;
  launches += launch_preprocess(v, (int)n, stream);
  launches += launch_threshold(v, (int)n, stream);
  launches += launch_ccl(v, (int)n, stream);
  launches += launch_cluster(v, (int)n, stream);
  launches += launch_quadfit(v, (int)n, stream);
  if (part == VIEW_ALL) {
    launches += launch_decode(v, (int)n, stream);
    launches += launch_finalize(v, (int)n, stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(out, v.out, sizeof(b200AprilTagsDetection_t) * (size_t)n * g.max_tags, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(cnt, v.out_count, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, stream);
    memset(ctr, 0, sizeof(uint32_t) * CNT_N * kMaxChunks);  // (host) only the first CNT_N entries are produced by this view
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctr, v.counters, sizeof(uint32_t) * CNT_N, cudaMemcpyDeviceToHost, stream);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200apriltags] enqueue failed: %s\n", cudaGetErrorString(e));
    return B200AT_ERR_CUDA;
  }
  if (launches_out) *launches_out = launches;
  return B200AT_OK;
}

int b200AprilTagsEnqueueBatch(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n, cudaStream_t stream) {
  if (!h || !frames || n == 0 || n > h->max_batch) return B200AT_ERR_INVALID_ARG;
  if (h->calls[0].active || h->calls[1].active) return B200AT_ERR_INVALID_ARG;  // one workspace: collect the host calls in flight first
  // up to two device batches in flight, on ONE stream (stream order is what makes the reuse of the workspace safe); with stage
  // timing on, one (the stage events are per handle)
  const uint64_t queued = h->dev_enq - h->dev_col;
  if (queued >= 2 || (queued == 1 && (h->timing || stream != h->cur_stream))) return B200AT_ERR_INVALID_ARG;
  cuAprilTagsHandle_st::DevSlot &d = h->dslot[h->dev_enq & 1];
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  int launches = h->launches;
  int fast = 1;
  int rc = fill_frame_table(h, frames, n, d.h_frames, &fast);
  if (rc == B200AT_OK) {
    h->ws.g.fast_align = fast;
    const bool graph_ok = h->use_graph && !h->timing && h->plain_calls >= 1;  // first call runs plain (lazy attribute setup)
    if (graph_ok && d.graph_exec && d.graph_n == n && d.graph_fast == fast && d.graph_enc == h->ws.g.enc && d.graph_stream == stream) {
      if (cudaGraphLaunch(d.graph_exec, stream) != cudaSuccess) rc = B200AT_ERR_CUDA;
      launches = d.launches;
    } else if (graph_ok) {
      if (d.graph_exec) {
        cudaGraphExecDestroy(d.graph_exec);
        d.graph_exec = nullptr;
      }
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
      if (e == cudaSuccess) {
        rc = enqueue_core(h, frames, n, stream, d.h_frames, d.h_out, d.h_out_count, d.h_counters, false, &launches, true);
        e = cudaStreamEndCapture(stream, &graph);
        if (rc == B200AT_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&d.graph_exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc == B200AT_OK && e == cudaSuccess && d.graph_exec) {
          d.graph_n = n;
          d.graph_fast = fast;
          d.graph_enc = h->ws.g.enc;
          d.graph_stream = stream;
          if (cudaGraphLaunch(d.graph_exec, stream) != cudaSuccess) rc = B200AT_ERR_CUDA;
        } else {
          // capture not possible in this context: fall back to plain launches from now on
          cudaGetLastError();
          h->use_graph = false;
          d.graph_exec = nullptr;
          rc = enqueue_core(h, frames, n, stream, d.h_frames, d.h_out, d.h_out_count, d.h_counters, h->timing, &launches, true);
        }
      } else {
        cudaGetLastError();
        h->use_graph = false;
        rc = enqueue_core(h, frames, n, stream, d.h_frames, d.h_out, d.h_out_count, d.h_counters, h->timing, &launches, true);
      }
    } else {
      rc = enqueue_core(h, frames, n, stream, d.h_frames, d.h_out, d.h_out_count, d.h_counters, h->timing, &launches, true);
      h->plain_calls++;
    }
  }
  if (rc == B200AT_OK && cudaEventRecord(d.done, stream) != cudaSuccess) rc = B200AT_ERR_CUDA;
  if (prev != h->device) cudaSetDevice(prev);
  if (rc != B200AT_OK) return rc;
  d.launches = launches;
  d.n = n;
  h->cur_stream = stream;
  h->dev_enq++;
  h->in_flight = true;
  return B200AT_OK;
}

int b200AprilTagsCollectBatch(cuAprilTagsHandle h, b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out, uint32_t *counts) {
  if (!h || h->dev_enq == h->dev_col) return B200AT_ERR_INVALID_ARG;
  cuAprilTagsHandle_st::DevSlot &d = h->dslot[h->dev_col & 1];  // the oldest batch in flight
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaError_t e = cudaEventSynchronize(d.done);
  h->dev_col++;
  h->in_flight = h->dev_enq != h->dev_col;
  if (e == cudaSuccess && h->timing) {
    for (int i = 0; i < B200AT_NUM_STAGES; i++) cudaEventElapsedTime(&h->stage_ms[i], h->ev[i], h->ev[i + 1]);
  }
  if (prev != h->device) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200apriltags] batch failed: %s\n", cudaGetErrorString(e));
    return B200AT_ERR_CUDA;
  }
  h->h_out = d.h_out;
  h->h_out_count = d.h_out_count;
  h->h_counters = d.h_counters;
  h->launches = d.launches;
  h->cur_n = d.n;
  const uint32_t mt = h->ws.g.max_tags;
  for (uint32_t i = 0; i < d.n; i++) {
    uint32_t c = d.h_out_count[i];
    if (c > mt) c = mt;
    if (counts) counts[i] = c;
    if (dets_out) memcpy(dets_out + (size_t)i * mt, d.h_out + (size_t)i * mt, sizeof(b200AprilTagsDetection_t) * c);
    if (ids_out)
      for (uint32_t k = 0; k < c; k++) to_id_struct(d.h_out[(size_t)i * mt + k], ids_out + (size_t)i * mt + k);
  }
  uint32_t agg[CNT_N];
  sum_counters(d.h_counters, agg);
  h->last_status = agg[CNT_STATUS];
  h->last_counters[0] = (uint64_t)d.launches;
  h->last_counters[1] = agg[CNT_POINTS];
  h->last_counters[2] = agg[CNT_CLUSTERS];
  h->last_counters[3] = agg[CNT_QUADS];
  uint64_t nc = 0;
  h->last_counters[4] = nc;
  h->last_counters[5] = agg[CNT_DETS];
  return h->last_status ? B200AT_ERR_OVERFLOW : B200AT_OK;
}

static cudaStream_t resolve_sync_stream(cuAprilTagsHandle h, cudaStream_t stream);

int b200AprilTagsDetectBatch(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n, b200AprilTagsDetection_t *dets_out,
                             cuAprilTagsID_t *ids_out, uint32_t *counts, cudaStream_t stream) {
  if (!h) return B200AT_ERR_INVALID_ARG;
  int rc = b200AprilTagsEnqueueBatch(h, frames, n, resolve_sync_stream(h, stream));
  if (rc != B200AT_OK) return rc;
  return b200AprilTagsCollectBatch(h, dets_out, ids_out, counts);
}

// ---------------------------------------------------------------------------------------------------------------------
// Host entry points: frames in HOST memory.  One persistent pipeline of SUB-BATCHES per handle; a call appends its sub-batches
// to it, and up to two calls may be in flight (b200AprilTagsEnqueueBatchHost / CollectBatchHost), so that the DMA and the quad
// detection of the next call overlap the decode / pose / D2H of the previous one -- a synchronous call pays the pipeline's fill
// and drain (~7 ms of a 28 ms call at batch 256, profiles/r02_host_path_trace.txt) every time.
//
// Sparse staging (frames pinned + device-mapped + 16-byte aligned, integer quad_decimate f >= 2; k_decode.cu): only rows 0, f,
// 2f.. are DMA'd, the full-resolution rows around the fitted quads are fetched afterwards straight from the caller's frames.
// Per sub-batch k (global counter: staging slot k % 3, workspace view k & 1, counters block k % 4):
//   copy stream    DMA(k)                                   after BACK(k-3) released the slot
//   compute stream FRONT(k) = frame table .. quad fit       after DMA(k)
//                  BACK(k-1) = refine, 2nd fetch, decode    after FETCH(k-1) and the tail of k-3 (same view: candidates / outputs)
//   fetch stream   FETCH(k) = mark + fetch for refine_edges after FRONT(k) and BACK(k-1)
//   tail stream    reconcile, pose, D2H of k-1              after BACK(k-1)
// The BACK of a call's last sub-batch is deferred until the next call's first FRONT has been queued (or until the call is
// collected), so the compute stream never idles behind a fetch.
// Anything else (pageable frames, f < 2, tiny max_batch) takes the full copy: sequential sub-batches on the whole workspace, two
// staging slots, copy(k+1) overlapping compute(k).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

constexpr uint32_t kHostSlotsMax = 3;

int host_streams_create(cuAprilTagsHandle h) {
  if (h->copy_stream) return B200AT_OK;
  // the fetch and tail kernels are a handful of latency-bound CTAs (PCIe reads; a warp per detection): at the highest priority
  // they get the first free SM slots instead of queueing behind the next sub-batch's full grids
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  bool ok = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaStreamCreateWithPriority(&h->fetch_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  ok = ok && cudaStreamCreateWithPriority(&h->tail_stream, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
  for (int i = 0; i < 3 && ok; i++) {
    ok = ok && cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming) == cudaSuccess;
  }
  for (int i = 0; i < 2 && ok; i++) {
    ok = ok && cudaEventCreateWithFlags(&h->ev_front[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_fetched[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_decoded[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_tail[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_backdone[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->calls[i].done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (ok) return B200AT_OK;
  // all or nothing: a half-created set would be taken for a complete one by the next call
  cudaGetLastError();
  auto ds = [](cudaStream_t &s) {
    if (s) cudaStreamDestroy(s);
    s = nullptr;
  };
  auto de = [](cudaEvent_t &e) {
    if (e) cudaEventDestroy(e);
    e = nullptr;
  };
  ds(h->copy_stream);
  ds(h->fetch_stream);
  ds(h->tail_stream);
  for (int i = 0; i < 3; i++) {
    de(h->ev_copied[i]);
    de(h->ev_consumed[i]);
  }
  for (int i = 0; i < 2; i++) {
    de(h->ev_front[i]);
    de(h->ev_fetched[i]);
    de(h->ev_decoded[i]);
    de(h->ev_tail[i]);
    de(h->ev_backdone[i]);
    de(h->calls[i].done);
  }
  return B200AT_ERR_CUDA;
}

// every stream the host path queues work on is idle afterwards
void host_drain(cuAprilTagsHandle h) {
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  if (h->fetch_stream) cudaStreamSynchronize(h->fetch_stream);
  if (h->tail_stream) cudaStreamSynchronize(h->tail_stream);
}

Workspace host_view_of(cuAprilTagsHandle h, uint64_t k, uint32_t S) {
  Workspace v = make_view(h, (int)(k & 1) * (int)S, (int)(k & 1), 2);
  v.counters = h->ws.counters + (size_t)(k % kMaxChunks) * CNT_N;  // FRONT(k + 2) zeroes another block than the tail of k reads
  return v;
}

// BACK of sub-batch `pb` (pipelined sparse path): refine, second fetch, decode on the compute stream; reconcile, pose and the D2H of
// the results on the tail stream
int host_enqueue_back(cuAprilTagsHandle h, const HostPendingBack &pb) {
  HostCall &call = h->calls[pb.call];
  const uint32_t mt = h->ws.g.max_tags;
  const int vb = (int)(pb.k & 1);
  cudaStream_t cs = h->own_stream;
  bool ok = cudaStreamWaitEvent(cs, h->ev_fetched[vb], 0) == cudaSuccess;
  // the decoder overwrites the candidates / outputs that the tail of this view's previous user (k - 2) still reads
  ok = ok && cudaStreamWaitEvent(cs, h->ev_tail[vb], 0) == cudaSuccess;
  if (!ok) return B200AT_ERR_CUDA;
  int l = 0;
  int rc = enqueue_view_part(h, host_view_of(h, pb.k, h->host_S), VIEW_BACK, pb.len, cs, call.out + (size_t)pb.start * mt, call.out_count + pb.start,
                             call.counters + (size_t)pb.sub * CNT_N * kMaxChunks, &l, h->tail_stream, h->ev_decoded[vb]);
  call.launches += l;
  if (rc != B200AT_OK) return rc;
  // the staged frames are dead once the decoder has run (the tail does not read them)
  ok = cudaEventRecord(h->ev_consumed[pb.k % kHostSlotsMax], cs) == cudaSuccess;
  ok = ok && cudaEventRecord(h->ev_backdone[vb], cs) == cudaSuccess;
  ok = ok && cudaEventRecord(h->ev_tail[vb], h->tail_stream) == cudaSuccess;
  if (ok && pb.last) ok = cudaEventRecord(call.done, h->tail_stream) == cudaSuccess;
  return ok ? B200AT_OK : B200AT_ERR_CUDA;
}

int host_flush_pending(cuAprilTagsHandle h) {
  if (!h->pending_back.valid) return B200AT_OK;
  h->pending_back.valid = false;
  return host_enqueue_back(h, h->pending_back);
}

int host_call_buffers(HostCall &c, uint32_t n, uint32_t nsub, uint32_t mt) {
  if (c.cap_frames >= n && c.cap_subs >= nsub) return B200AT_OK;
  if (c.frames_tab) cudaFreeHost(c.frames_tab);
  if (c.src_tab) cudaFreeHost(c.src_tab);
  if (c.out) cudaFreeHost(c.out);
  if (c.out_count) cudaFreeHost(c.out_count);
  if (c.counters) cudaFreeHost(c.counters);
  c.frames_tab = c.src_tab = nullptr;
  c.out = nullptr;
  c.out_count = c.counters = nullptr;
  c.cap_frames = c.cap_subs = 0;
  const uint32_t nf = std::max(n, (uint32_t)c.cap_frames), ns = std::max(nsub, (uint32_t)c.cap_subs);
  if (cudaMallocHost(&c.frames_tab, sizeof(FrameDesc) * nf) != cudaSuccess || cudaMallocHost(&c.src_tab, sizeof(FrameDesc) * nf) != cudaSuccess ||
      cudaMallocHost(&c.out, sizeof(b200AprilTagsDetection_t) * (size_t)nf * mt) != cudaSuccess ||
      cudaMallocHost(&c.out_count, sizeof(uint32_t) * nf) != cudaSuccess ||
      cudaMallocHost(&c.counters, sizeof(uint32_t) * CNT_N * kMaxChunks * ns) != cudaSuccess) {
    cudaGetLastError();
    return B200AT_ERR_NOMEM;
  }
  c.cap_frames = nf;
  c.cap_subs = ns;
  return B200AT_OK;
}

int host_enqueue(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n) {
  const Geo &g = h->ws.g;
  if (h->ws.rect_map) return B200AT_ERR_UNSUPPORTED;  // the rectify pre-stage takes device frames (include/b200_apriltags.h)
  const uint32_t mt = g.max_tags;
  const size_t row = (size_t)g.W * g.bpp;
  // ---- everything that can fail on the arguments is checked before anything is queued ----
  for (uint32_t i = 0; i < n; i++)
    if (!frames[i].ptr || frames[i].pitch < row) return B200AT_ERR_INVALID_ARG;
  const int slot_id = h->calls[0].active ? 1 : 0;
  if (h->calls[slot_id].active) return B200AT_ERR_INVALID_ARG;  // two calls in flight already: collect one first
  const uint32_t in_flight = (h->calls[0].active ? 1u : 0u) + (h->calls[1].active ? 1u : 0u);
  int rc = host_streams_create(h);
  if (rc != B200AT_OK) return rc;
  HostCall &call = h->calls[slot_id];
  // (the encoding, hence the row size, can change between calls: b200AprilTagsSetInputEncoding)
  const size_t need_pitch = (row + 255) & ~(size_t)255;
  // Sparse staging needs pinned, device-mapped, 16-byte aligned host frames; anything else takes the full copy.
  // B200AT_SPARSE_H2D=0/1 overrides the default.  A frame that is covered with tags (a calibration grid) needs nearly every row:
  // the sparse path then moves as many bytes as the full copy, in smaller pieces -- after such a call the next eight take the
  // full copy, then sparse staging is tried again.
  bool sparse = true, sparse_forced = false;
  if (const char *es = getenv("B200AT_SPARSE_H2D")) {
    sparse = atoi(es) != 0;
    sparse_forced = true;
  }
  if (sparse && !sparse_forced && h->sparse_skip > 0) {
    h->sparse_skip--;
    sparse = false;
  }
  if (g.f < 2) sparse = false;
  std::vector<FrameDesc> src(sparse ? n : 0);
  for (uint32_t i = 0; i < n && sparse; i++) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, frames[i].ptr) != cudaSuccess) {
      cudaGetLastError();
      sparse = false;
      break;
    }
    if (pa.type != cudaMemoryTypeHost || !pa.devicePointer || ((uintptr_t)pa.devicePointer & 15) || (frames[i].pitch & 15)) {
      sparse = false;
      break;
    }
    src[i].ptr = (const uint8_t *)pa.devicePointer;
    src[i].pitch = frames[i].pitch;
  }
  // sub-batch size.  Sparse: the call is compute-bound and likes large sub-batches (the fixed cost per sub-batch -- ~30 launches,
  // persistent-kernel tails -- is ~0.4 ms): max_batch / 4, two workspace views.  Full copy: PCIe-bound, short pipeline fill.
  uint32_t S = sparse ? std::max<uint32_t>(1, std::min<uint32_t>(h->max_batch / 2, std::max<uint32_t>(16, (h->max_batch + 3) / 4)))
                      : std::max<uint32_t>(1, std::min<uint32_t>(h->max_batch, std::max<uint32_t>(16, (h->max_batch + 15) / 16)));
  if (const char *es = getenv("B200AT_HOST_SUB")) {
    const int v = atoi(es);
    if (v >= 1) S = std::min<uint32_t>(sparse ? std::max<uint32_t>(1, h->max_batch / 2) : h->max_batch, (uint32_t)v);
  }
  const bool pipe = sparse && 2 * S <= h->max_batch;
  if (!pipe) S = std::min<uint32_t>(S, h->max_batch);
  const uint32_t nslots = pipe ? 3 : 2;
  // a change of mode, sub-batch size or staging geometry re-shapes what the calls in flight are using: drain first
  const bool reshape = !h->d_stage || need_pitch > h->stage_pitch || h->stage_sub < S || h->stage_slots < nslots;
  if (in_flight && (reshape || h->host_pipe != pipe || h->host_S != S)) {
    rc = host_flush_pending(h);
    host_drain(h);
    if (rc != B200AT_OK) return rc;
  }
  if (reshape) {
    host_drain(h);
    if (h->d_stage) {
      for (size_t i = 0; i < h->dev_allocs.size(); i++)
        if (h->dev_allocs[i] == h->d_stage) {
          h->dev_allocs.erase(h->dev_allocs.begin() + (long)i);
          break;
        }
      cudaFree(h->d_stage);
      h->d_stage = nullptr;
    }
    void *p = nullptr;
    const size_t pitch = std::max(need_pitch, h->stage_pitch);
    const uint32_t ss = std::max(S, h->stage_sub), sl = std::max(nslots, h->stage_slots);
    if (cudaMalloc(&p, pitch * g.H * (size_t)ss * sl) != cudaSuccess) {
      cudaGetLastError();
      h->stage_sub = h->stage_slots = 0;
      return B200AT_ERR_NOMEM;
    }
    h->dev_allocs.push_back(p);
    h->d_stage = (uint8_t *)p;
    h->stage_pitch = pitch;
    h->stage_sub = ss;
    h->stage_slots = sl;
  }
  if (sparse && !h->sparse_bufs) {
    Workspace &w = h->ws;
    const size_t B = h->max_batch;
    int rca = dev_alloc(h, &w.need1, B * g.H);
    if (rca == 0) rca = dev_alloc(h, &w.need2, B * g.H);
    if (rca == 0) rca = dev_alloc(h, &w.src_frames, B);
    if (rca == 0 && !w.quad_H) rca = dev_alloc(h, &w.quad_H, (size_t)g.quad_cap * 10);
    if (rca != 0) return rca;
    h->sparse_bufs = true;
  }
  // sub-batches.  With nothing in flight the first two are small, so that the kernels start after a short first DMA
  // (B200AT_HOST_RAMP=0 turns the ramp off); with a call in flight the pipeline is full already.
  bool ramp = pipe && in_flight == 0;
  if (const char *es = getenv("B200AT_HOST_RAMP")) ramp = ramp && atoi(es) != 0;
  std::vector<uint32_t> sub_start, sub_len;
  for (uint32_t pos = 0, j = 0; pos < n; j++) {
    uint32_t len = S;
    if (ramp && j == 0) len = std::max<uint32_t>(1, S / 4);
    if (ramp && j == 1) len = std::max<uint32_t>(1, S / 2);
    len = std::min<uint32_t>(len, n - pos);
    sub_start.push_back(pos);
    sub_len.push_back(len);
    pos += len;
  }
  const uint32_t nsub = (uint32_t)sub_start.size();
  rc = host_call_buffers(call, n, nsub, mt);
  if (rc != B200AT_OK) return rc;
  if (sparse) memcpy(call.src_tab, src.data(), sizeof(FrameDesc) * n);
  call.n = n;
  call.nsub = nsub;
  call.sparse = sparse;
  call.launches = 0;
  call.dma_bytes = 0;
  call.active = true;
  call.seq = h->host_seq++;
  h->host_pipe = pipe;
  h->host_S = S;
  const int row_step = sparse ? g.f : 1;
  const int rows_dma = 1 + (g.H - 1) / row_step;
  static const bool sparse_debug = getenv("B200AT_SPARSE_DEBUG") != nullptr;  // poison the slot: an unfetched row cannot go unnoticed
  std::vector<b200AprilTagsFrame_t> dframes(S);
  cudaStream_t cs = h->own_stream;
  rc = B200AT_OK;
  for (uint32_t j = 0; j < nsub && rc == B200AT_OK; j++) {
    const uint64_t k = h->host_k++;
    const uint32_t i0 = sub_start[j], m = sub_len[j];
    const int slot = (int)(k % nslots);
    uint8_t *slot_base = h->d_stage + (size_t)slot * h->stage_sub * h->stage_pitch * g.H;
    cudaError_t e = cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0);  // staging slot free again (no-op before its first use)
    if (sparse && sparse_debug && e == cudaSuccess) e = cudaMemsetAsync(slot_base, 0xA5, (size_t)m * h->stage_pitch * g.H, h->copy_stream);
    for (uint32_t q = 0; q < m && e == cudaSuccess;) {
      uint8_t *dst = slot_base + (size_t)q * h->stage_pitch * g.H;
      // rows 0, f, 2f, ... (all rows on the full-copy path).  Frames that follow each other in the caller's memory (one allocation
      // for the batch) go out as ONE 2-D copy: with H a multiple of the row step the row pattern continues across the frame
      // boundary on both sides.  (Measured: 256 copies of 540 rows 51.4 GB/s, 4 copies of 34,560 rows 55.2 GB/s = the rate of a
      // contiguous copy on the same box.)
      uint32_t cnt = 1;
      if (g.H % row_step == 0)
        while (q + cnt < m && frames[i0 + q + cnt].pitch == frames[i0 + q].pitch &&
               (const uint8_t *)frames[i0 + q + cnt].ptr == (const uint8_t *)frames[i0 + q].ptr + (size_t)cnt * frames[i0 + q].pitch * g.H)
          cnt++;
      e = cudaMemcpy2DAsync(dst, h->stage_pitch * row_step, frames[i0 + q].ptr, frames[i0 + q].pitch * row_step, row, (size_t)rows_dma * cnt,
                            cudaMemcpyHostToDevice, h->copy_stream);
      call.dma_bytes += (uint64_t)row * rows_dma * cnt;
      for (uint32_t u = 0; u < cnt; u++) {
        dframes[q + u].ptr = slot_base + (size_t)(q + u) * h->stage_pitch * g.H;
        dframes[q + u].pitch = h->stage_pitch;
      }
      q += cnt;
    }
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_copied[slot], h->copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, h->ev_copied[slot], 0);
    if (e != cudaSuccess) {
      rc = B200AT_ERR_CUDA;
      break;
    }
    int l = 0;
    if (pipe) {
      const int vw = (int)(k & 1);
      rc = enqueue_view(h, host_view_of(h, k, S), dframes.data(), m, cs, call.frames_tab + i0, nullptr, nullptr, nullptr, &l, call.src_tab + i0, VIEW_FRONT);
      call.launches += l;
      if (rc == B200AT_OK && cudaEventRecord(h->ev_front[vw], cs) != cudaSuccess) rc = B200AT_ERR_CUDA;
      // BACK of the previous sub-batch (this call's, or the one the previous call left pending)
      if (rc == B200AT_OK) rc = host_flush_pending(h);
      // FETCH(k): after FRONT(k) and after BACK(k - 1), whose second fetch would otherwise queue behind this one on PCIe
      if (rc == B200AT_OK && cudaStreamWaitEvent(h->fetch_stream, h->ev_front[vw], 0) != cudaSuccess) rc = B200AT_ERR_CUDA;
      if (rc == B200AT_OK && cudaStreamWaitEvent(h->fetch_stream, h->ev_backdone[vw ^ 1], 0) != cudaSuccess) rc = B200AT_ERR_CUDA;
      if (rc == B200AT_OK) rc = enqueue_view_part(h, host_view_of(h, k, S), VIEW_FETCH, m, h->fetch_stream, nullptr, nullptr, nullptr, &l);
      call.launches += l;
      if (rc == B200AT_OK && cudaEventRecord(h->ev_fetched[vw], h->fetch_stream) != cudaSuccess) rc = B200AT_ERR_CUDA;
      if (rc == B200AT_OK) {
        HostPendingBack &pb = h->pending_back;
        pb.valid = true;
        pb.call = slot_id;
        pb.k = k;
        pb.sub = j;
        pb.start = i0;
        pb.len = m;
        pb.last = j + 1 == nsub;
      }
      continue;
    }
    rc = enqueue_core(h, dframes.data(), m, cs, call.frames_tab + i0, call.out + (size_t)i0 * mt, call.out_count + i0,
                      call.counters + (size_t)j * CNT_N * kMaxChunks, false, &l, false, sparse ? call.src_tab + i0 : nullptr);
    call.launches += l;
    if (rc == B200AT_OK && cudaEventRecord(h->ev_consumed[slot], cs) != cudaSuccess) rc = B200AT_ERR_CUDA;
    if (rc == B200AT_OK && j + 1 == nsub && cudaEventRecord(call.done, cs) != cudaSuccess) rc = B200AT_ERR_CUDA;
  }
  if (rc != B200AT_OK) {
    // nothing of this call may still be running when the error is reported (the caller is free to release its frames)
    h->pending_back.valid = false;
    host_drain(h);
    call.active = false;
    if (h->calls[slot_id ^ 1].active) h->calls[slot_id ^ 1].failed = true;  // (the drain cut its pending work short)
    return rc;
  }
  return B200AT_OK;
}

int host_collect(cuAprilTagsHandle h, b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out, uint32_t *counts) {
  // the older of the calls in flight
  int id = -1;
  for (int i = 0; i < 2; i++)
    if (h->calls[i].active && (id < 0 || h->calls[i].seq < h->calls[id].seq)) id = i;
  if (id < 0) return B200AT_ERR_INVALID_ARG;
  HostCall &call = h->calls[id];
  const Geo &g = h->ws.g;
  const uint32_t mt = g.max_tags, n = call.n;
  int rc = B200AT_OK;
  if (h->pending_back.valid && h->pending_back.call == id) rc = host_flush_pending(h);
  cudaError_t es = rc == B200AT_OK ? cudaEventSynchronize(call.done) : cudaSuccess;
  call.active = false;
  if (rc != B200AT_OK || es != cudaSuccess || call.failed) {
    call.failed = false;
    host_drain(h);
    if (es != cudaSuccess) fprintf(stderr, "[b200apriltags] batch failed: %s\n", cudaGetErrorString(es));
    return rc != B200AT_OK ? rc : B200AT_ERR_CUDA;
  }
  uint32_t status = 0;
  uint64_t pts = 0, clu = 0, qd = 0, dt = 0, fetched = 0;
  for (uint32_t k = 0; k < call.nsub; k++) {
    uint32_t c[CNT_N];
    sum_counters(call.counters + (size_t)k * CNT_N * kMaxChunks, c);
    status |= c[CNT_STATUS];
    fetched += c[CNT_FETCHED];
    pts += c[CNT_POINTS];
    clu += c[CNT_CLUSTERS];
    qd += c[CNT_QUADS];
    dt += c[CNT_DETS];
  }
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t c = std::min(call.out_count[i], mt);
    if (counts) counts[i] = c;
    if (dets_out) memcpy(dets_out + (size_t)i * mt, call.out + (size_t)i * mt, sizeof(b200AprilTagsDetection_t) * c);
    if (ids_out)
      for (uint32_t k = 0; k < c; k++) to_id_struct(call.out[(size_t)i * mt + k], ids_out + (size_t)i * mt + k);
  }
  h->last_status = status;
  h->launches = call.launches;
  h->last_counters[0] = (uint64_t)call.launches;
  h->last_counters[1] = pts;
  h->last_counters[2] = clu;
  h->last_counters[3] = qd;
  h->last_counters[5] = dt;
  h->last_counters[6] = call.dma_bytes + fetched * 16;  // host->device bytes of this call: DMA + rows fetched on demand
  h->last_counters[7] = call.sparse ? 1 : 0;
  {
    const char *eb = getenv("B200AT_SPARSE_BACKOFF");
    const double backoff = eb ? atof(eb) : 0.85;
    const size_t row = (size_t)g.W * g.bpp;
    if (call.sparse && (double)(call.dma_bytes + fetched * 16) > backoff * (double)row * g.H * n) h->sparse_skip = 8;
  }
  return status ? B200AT_ERR_OVERFLOW : B200AT_OK;
}

}  // namespace

int b200AprilTagsEnqueueBatchHost(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n) {
  if (!h || !frames || n == 0 || h->in_flight) return B200AT_ERR_INVALID_ARG;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  const int rc = host_enqueue(h, frames, n);
  if (prev != h->device) cudaSetDevice(prev);
  return rc;
}

int b200AprilTagsCollectBatchHost(cuAprilTagsHandle h, b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out, uint32_t *counts) {
  if (!h) return B200AT_ERR_INVALID_ARG;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  const int rc = host_collect(h, dets_out, ids_out, counts);
  if (prev != h->device) cudaSetDevice(prev);
  return rc;
}

int b200AprilTagsDetectBatchHost(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n, b200AprilTagsDetection_t *dets_out,
                                 cuAprilTagsID_t *ids_out, uint32_t *counts) {
  if (!h || h->calls[0].active || h->calls[1].active) return B200AT_ERR_INVALID_ARG;  // (asynchronous calls in flight: collect them first)
  const int rc = b200AprilTagsEnqueueBatchHost(h, frames, n);
  if (rc != B200AT_OK) return rc;
  return b200AprilTagsCollectBatchHost(h, dets_out, ids_out, counts);
}

// A caller that passes the legacy default stream (nullptr) still gets the CUDA-graph replay path: capture is not possible on
// the legacy stream, so the batch runs on the handle's own stream, ordered after whatever the caller queued on the legacy
// stream before the call (the call is synchronous, so everything the caller queues afterwards is ordered by the host).
static cudaStream_t resolve_sync_stream(cuAprilTagsHandle h, cudaStream_t stream) {
  if (stream != nullptr) return stream;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaStream_t use = h->own_stream;
  if (cudaEventRecord(h->ev_in, cudaStreamLegacy) != cudaSuccess || cudaStreamWaitEvent(h->own_stream, h->ev_in, 0) != cudaSuccess) {
    cudaGetLastError();
    use = nullptr;  // (cannot order against the legacy stream: run there, without the graph)
  }
  if (prev != h->device) cudaSetDevice(prev);
  return use;
}

uint32_t cuAprilTagsDetect(cuAprilTagsHandle h, const cuAprilTagsImageInput_t *img, cuAprilTagsID_t *tags_out, uint32_t *num_tags,
                           const uint32_t max_tags, cudaStream_t stream) {
  if (!h || !img || !tags_out || !num_tags) return B200AT_ERR_INVALID_ARG;
  if (h->ws.rect_map ? (img->width != h->ws.rect_src_w || img->height != h->ws.rect_src_h) : (img->width != h->ws.g.W || img->height != h->ws.g.H))
    return B200AT_ERR_INVALID_ARG;
  if (h->ws.g.bpp != 3) return B200AT_ERR_INVALID_ARG;  // the uchar3 entry point carries rgb8/bgr8 only
  b200AprilTagsFrame_t fr;
  fr.ptr = img->dev_ptr;
  fr.pitch = img->pitch;
  uint32_t cnt = 0;
  int rc = b200AprilTagsEnqueueBatch(h, &fr, 1, resolve_sync_stream(h, stream));
  if (rc != B200AT_OK) return (uint32_t)rc;
  rc = b200AprilTagsCollectBatch(h, nullptr, nullptr, &cnt);
  if (rc != B200AT_OK && rc != B200AT_ERR_OVERFLOW) return (uint32_t)rc;
  if (cnt > max_tags) cnt = max_tags;
  for (uint32_t k = 0; k < cnt; k++) to_id_struct(h->h_out[k], tags_out + k);
  *num_tags = cnt;
  // A bounded buffer that overflowed truncates the list; the detections that were written are valid.  The reference's caller
  // drops the whole frame on ANY non-zero return (apriltag_node.cpp:494-497), so truncation is reported through
  // b200AprilTagsLastStatus only, like a cuAprilTags detector that fills max_tags entries and returns success.
  return 0u;
}

int b200AprilTagsLastStatus(cuAprilTagsHandle h, uint32_t *status) {
  if (!h || !status) return B200AT_ERR_INVALID_ARG;
  *status = h->last_status;
  return B200AT_OK;
}

int b200AprilTagsGetStageTimes(cuAprilTagsHandle h, float *ms) {
  if (!h || !ms) return B200AT_ERR_INVALID_ARG;
  for (int i = 0; i < B200AT_NUM_STAGES; i++) ms[i] = h->stage_ms[i];
  return B200AT_OK;
}

int b200AprilTagsGetCounters(cuAprilTagsHandle h, uint64_t *c) {
  if (!h || !c) return B200AT_ERR_INVALID_ARG;
  for (int i = 0; i < 8; i++) c[i] = h->last_counters[i];
  return B200AT_OK;
}

int b200AprilTagsGetDims(cuAprilTagsHandle h, uint32_t *wd, uint32_t *hd, uint32_t *tw, uint32_t *th) {
  if (!h) return B200AT_ERR_INVALID_ARG;
  if (wd) *wd = h->ws.g.Wd;
  if (hd) *hd = h->ws.g.Hd;
  if (tw) *tw = h->ws.g.tw;
  if (th) *th = h->ws.g.th;
  return B200AT_OK;
}

int b200AprilTagsReadBuffer(cuAprilTagsHandle h, int which, uint32_t frame, void *dst, size_t cap, size_t *n_elems) {
  if (!h || frame >= h->max_batch) return B200AT_ERR_INVALID_ARG;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != h->device) cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  const Workspace &ws = h->ws;
  const Geo &g = ws.g;
  const size_t Wp = (size_t)at_Wp(g), Pp = Wp * g.Hd;
  cudaError_t e = cudaSuccess;
  size_t n = 0;
  auto conv_label = [&](uint32_t v) -> uint32_t { return (uint32_t)((v / Wp) * g.Wd + (v % Wp)); };
  auto conv_key = [&](unsigned long long k) -> unsigned long long {
    return ((unsigned long long)conv_label((uint32_t)(k >> 32)) << 32) | conv_label((uint32_t)(k & 0xffffffffu));
  };
  switch (which) {
    case B200AT_BUF_RECTIFIED: {
      if (!ws.rect_img) {
        if (prev != h->device) cudaSetDevice(prev);
        return B200AT_ERR_INVALID_ARG;
      }
      n = (size_t)g.W * g.H;
      if (dst && cap >= n)
        e = cudaMemcpy2D(dst, g.W, ws.rect_img + (size_t)frame * g.H * ws.rect_pitch, ws.rect_pitch, g.W, g.H, cudaMemcpyDeviceToHost);
      break;
    }
    case B200AT_BUF_DECIMATED:
    case B200AT_BUF_THRESHOLD: {
      const uint8_t *src = (which == B200AT_BUF_DECIMATED ? ws.dec : ws.thr) + (size_t)frame * Pp;
      n = (size_t)g.Wd * g.Hd;
      if (dst && cap >= n) e = cudaMemcpy2D(dst, g.Wd, src, Wp, g.Wd, g.Hd, cudaMemcpyDeviceToHost);
      break;
    }
    case B200AT_BUF_TILE_MIN:
    case B200AT_BUF_TILE_MAX: {
      const size_t twp = at_twp(g);
      const uint8_t *src = (which == B200AT_BUF_TILE_MIN ? ws.tmin : ws.tmax) + (size_t)frame * g.th * twp;
      n = (size_t)g.tw * g.th;
      if (dst && cap >= n) e = cudaMemcpy2D(dst, g.tw, src, twp, g.tw, g.th, cudaMemcpyDeviceToHost);
      break;
    }
    case B200AT_BUF_LABELS:
    case B200AT_BUF_SIZES: {
      const uint32_t *src = (which == B200AT_BUF_LABELS ? ws.lab : ws.csize) + (size_t)frame * Pp;
      n = (size_t)g.Wd * g.Hd;
      if (dst && cap >= n * 4) {
        e = cudaMemcpy2D(dst, (size_t)g.Wd * 4, src, Wp * 4, (size_t)g.Wd * 4, g.Hd, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && which == B200AT_BUF_LABELS && Wp != (size_t)g.Wd) {
          uint32_t *d = (uint32_t *)dst;
          for (size_t i = 0; i < n; i++) d[i] = conv_label(d[i]);
        }
        if (e == cudaSuccess && which == B200AT_BUF_SIZES) {
          // the device image holds a count at component representatives only (everything else is never written): present it as
          // "size at the representative, 1 at 127-pixels (AprilRobotics never connects them: singletons), 0 elsewhere"
          std::vector<uint32_t> lab_h(n);
          std::vector<uint8_t> thr_h(n);
          e = cudaMemcpy2D(lab_h.data(), (size_t)g.Wd * 4, ws.lab + (size_t)frame * Pp, Wp * 4, (size_t)g.Wd * 4, g.Hd, cudaMemcpyDeviceToHost);
          if (e == cudaSuccess) e = cudaMemcpy2D(thr_h.data(), g.Wd, ws.thr + (size_t)frame * Pp, Wp, g.Wd, g.Hd, cudaMemcpyDeviceToHost);
          uint32_t *d = (uint32_t *)dst;
          if (e == cudaSuccess)
            for (size_t i = 0; i < n; i++) {
              const size_t self = (i / g.Wd) * Wp + (i % g.Wd);
              d[i] = thr_h[i] == 127 ? 1u : (lab_h[i] == (uint32_t)self ? d[i] : 0u);
            }
        }
      }
      break;
    }
    case B200AT_BUF_CLUSTERS: {
      n = std::min<size_t>(h->h_counters[CNT_CLUSTERS], g.clu_cap);
      size_t m = std::min(n, cap / sizeof(ClusterRec));
      if (dst && m) {
        e = cudaMemcpy(dst, ws.clusters, m * sizeof(ClusterRec), cudaMemcpyDeviceToHost);
        ClusterRec *d = (ClusterRec *)dst;
        if (e == cudaSuccess && Wp != (size_t)g.Wd)
          for (size_t i = 0; i < m; i++) d[i].key = conv_key(d[i].key);
      }
      break;
    }
    case B200AT_BUF_POINTS: {
      n = std::min<size_t>(h->h_counters[CNT_POINTS], g.pts_cap);
      size_t m = std::min(n, cap / sizeof(unsigned long long));
      if (dst && m) e = cudaMemcpy(dst, ws.keys, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      break;
    }
    case B200AT_BUF_POINTS_RAW: {
      n = std::min<size_t>(h->h_counters[CNT_POINTS], g.pts_cap);
      size_t m = std::min(n, cap / sizeof(uint32_t));
      if (dst && m) e = cudaMemcpy(dst, ws.pts, m * sizeof(uint32_t), cudaMemcpyDeviceToHost);
      break;
    }
    case B200AT_BUF_QUADS:
    case B200AT_BUF_QUADS_REFINED: {
      n = std::min<size_t>(h->h_counters[CNT_QUADS], g.quad_cap);
      size_t m = std::min(n, cap / sizeof(QuadRec));
      if (dst && m) {
        e = cudaMemcpy(dst, which == B200AT_BUF_QUADS ? ws.quads : ws.quads_refined, m * sizeof(QuadRec), cudaMemcpyDeviceToHost);
        QuadRec *d = (QuadRec *)dst;
        if (e == cudaSuccess && Wp != (size_t)g.Wd)
          for (size_t i = 0; i < m; i++) d[i].key = conv_key(d[i].key);
      }
      break;
    }
    default:
      if (prev != h->device) cudaSetDevice(prev);
      return B200AT_ERR_INVALID_ARG;
  }
  if (n_elems) *n_elems = n;
  if (prev != h->device) cudaSetDevice(prev);
  return e == cudaSuccess ? B200AT_OK : B200AT_ERR_CUDA;
}

}  // extern "C"
