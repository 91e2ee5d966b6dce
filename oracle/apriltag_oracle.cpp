/*
 * apriltag_oracle.cpp -- CPU ORACLE: restatement of AprilRobotics apriltag 3.x (v3.4.x semantics).
 *
 * TEST INFRASTRUCTURE ONLY (see apriltag_oracle.h).  Every function cites the upstream file/function
 * it restates; the upstream sources are not in /root/reference (the reference node only links closed
 * libraries: /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:450-452,491-493,228-231,290-301),
 * so citations are by upstream name and SURVEY.md Appendix A section.
 *
 * Deliberate canonicalisations (upstream leaves these orders to hash-bucket / thread-merge accidents;
 * they are stated here so the CUDA path can reproduce them exactly):
 *   C1. connected-component representative = minimum pixel index of the component.
 *   C2. clusters are processed in ascending key order, key = (max(rep)<<32 | min(rep)).
 *   C3. boundary points with equal float `slope` are ordered by (y, x) (upstream: its merge sort's
 *       tie order, which depends on insertion order and recursion shape).
 *   C4. final detections are ordered by (id, family, c.y, c.x) (upstream: qsort by id, unstable).
 * Everything else follows upstream operation by operation, in the same floating-point types and
 * evaluation order; build with -ffp-contract=off so no FMA contraction changes roundings.
 */
#include "apriltag_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <string>
#include <vector>

#include "tag_families_data.inc"

namespace {

using std::vector;

struct Family {
  const char *name;
  int nbits, h, ncodes, width_at_border, total_width;
  bool reversed_border;
  const signed char *bit_x, *bit_y;  // relative to the border's first cell: negative / >= width_at_border = outside the border
  const unsigned long long *codes;
};

const Family kFamilies[ATO_NUM_FAMILIES] = {
    {"tag36h11", tag36h11_nbits, tag36h11_h, tag36h11_ncodes, tag36h11_width_at_border, tag36h11_total_width, false,
     (const signed char *)tag36h11_bit_x, (const signed char *)tag36h11_bit_y, tag36h11_codes},
    {"tag25h9", tag25h9_nbits, tag25h9_h, tag25h9_ncodes, tag25h9_width_at_border, tag25h9_total_width, false,
     (const signed char *)tag25h9_bit_x, (const signed char *)tag25h9_bit_y, tag25h9_codes},
    {"tag16h5", tag16h5_nbits, tag16h5_h, tag16h5_ncodes, tag16h5_width_at_border, tag16h5_total_width, false,
     (const signed char *)tag16h5_bit_x, (const signed char *)tag16h5_bit_y, tag16h5_codes},
    {"tag36h10", tag36h10_nbits, tag36h10_h, tag36h10_ncodes, tag36h10_width_at_border, tag36h10_total_width, false,
     (const signed char *)tag36h10_bit_x, (const signed char *)tag36h10_bit_y, tag36h10_codes},
};

// Families registered at run time (ato_register_family): the same upstream tag family struct, filled by the caller -- how the
// families without a built-in table (upstream tagStandard41h12.c, tagCircle21h7.c, ...: reversed border, bits outside the border) are
// supplied, and what the parity tests use for a synthetic reversed-border family.
struct CustomFamily {
  Family f{};
  std::string name;
  std::vector<signed char> bx, by;
  std::vector<unsigned long long> codes;
  bool used = false;
};
CustomFamily g_custom[ATO_MAX_FAMILIES - ATO_NUM_FAMILIES];

const Family &family_at(int f) { return f < ATO_NUM_FAMILIES ? kFamilies[f] : g_custom[f - ATO_NUM_FAMILIES].f; }

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Image {
  int w = 0, h = 0;
  vector<uint8_t> buf;  // stride == w
  uint8_t at(int x, int y) const { return buf[(size_t)y * w + x]; }
};

struct Pt {
  uint16_t x, y;  // 2*actual value
  int16_t gx, gy;
  float slope;
};

struct LineFitPt {
  double Mx, My, Mxx, Mxy, Myy, W;
};

struct Quad {
  float p[4][2];
  bool reversed_border;
  uint64_t key;
  double H[9];
};

struct Cluster {
  uint64_t key;
  vector<Pt> pts;
};

// ---------------------------------------------------------------------------------------------
// upstream common/unionfind.h : unionfind_t (size-weighted union, path-halving find)
// ---------------------------------------------------------------------------------------------
struct UnionFind {
  vector<uint32_t> parent, size;
  void init(uint32_t n) {
    parent.resize(n);
    size.assign(n, 1);
    for (uint32_t i = 0; i < n; i++) parent[i] = i;
  }
  uint32_t find(uint32_t id) {
    uint32_t root = id;
    while (parent[root] != root) root = parent[root];
    while (parent[id] != root) {
      uint32_t nx = parent[id];
      parent[id] = root;
      id = nx;
    }
    return root;
  }
  uint32_t set_size(uint32_t id) { return size[find(id)]; }
  void connect(uint32_t a, uint32_t b) {
    uint32_t ra = find(a), rb = find(b);
    if (ra == rb) return;
    if (size[ra] > size[rb]) {
      parent[rb] = ra;
      size[ra] += size[rb];
    } else {
      parent[ra] = rb;
      size[rb] += size[ra];
    }
  }
};

struct Detector {
  ato_params_t prm;
  vector<int> fams;  // registered family indices
  // intermediates of last call
  Image quad_im, thresh;
  vector<uint8_t> tmin, tmax;
  int tw = 0, th = 0;
  vector<uint32_t> label, lsize;
  vector<Cluster> clusters;  // kept clusters, sorted (post fit-ordering: pts sorted by slope)
  int total_points = 0;
  vector<Quad> quads_fit, quads_refined;
  ato_times_t times{};
};

// ---------------------------------------------------------------------------------------------
// upstream common/image_u8.c : image_u8_decimate   
// ---------------------------------------------------------------------------------------------
Image decimate(const uint8_t *im, int width, int height, int stride, float ffactor) {
  Image d;
  if (ffactor == 1.5f) {
    int swidth = width / 3 * 2, sheight = height / 3 * 2;
    d.w = swidth;
    d.h = sheight;
    d.buf.assign((size_t)swidth * sheight, 0);
    int y = 0, sy = 0;
    while (sy < sheight) {
      int x = 0, sx = 0;
      while (sx < swidth) {
        const uint8_t *r0 = im + (size_t)(y + 0) * stride + x;
        const uint8_t *r1 = im + (size_t)(y + 1) * stride + x;
        const uint8_t *r2 = im + (size_t)(y + 2) * stride + x;
        int a = r0[0], b = r0[1], c = r0[2], dd = r1[0], e = r1[1], f = r1[2], g = r2[0], hh = r2[1], i = r2[2];
        d.buf[(size_t)(sy + 0) * swidth + sx + 0] = (uint8_t)((4 * a + 2 * b + 2 * dd + e) / 9);
        d.buf[(size_t)(sy + 0) * swidth + sx + 1] = (uint8_t)((4 * c + 2 * b + 2 * f + e) / 9);
        d.buf[(size_t)(sy + 1) * swidth + sx + 0] = (uint8_t)((4 * g + 2 * dd + 2 * hh + e) / 9);
        d.buf[(size_t)(sy + 1) * swidth + sx + 1] = (uint8_t)((4 * i + 2 * f + 2 * hh + e) / 9);
        x += 3;
        sx += 2;
      }
      y += 3;
      sy += 2;
    }
    return d;
  }
  int factor = (int)ffactor;
  int swidth = 1 + (width - 1) / factor;
  int sheight = 1 + (height - 1) / factor;
  d.w = swidth;
  d.h = sheight;
  d.buf.assign((size_t)swidth * sheight, 0);
  int sy = 0;
  for (int y = 0; y < height; y += factor) {
    int sx = 0;
    for (int x = 0; x < width; x += factor) {
      d.buf[(size_t)sy * swidth + sx] = im[(size_t)y * stride + x];
      sx++;
    }
    sy++;
  }
  return d;
}

// ---------------------------------------------------------------------------------------------
// upstream common/image_u8.c : convolve / image_u8_convolve_2D / image_u8_gaussian_blur (App. A.1)
// ---------------------------------------------------------------------------------------------
void convolve1d(const uint8_t *x, uint8_t *y, int sz, const uint8_t *k, int ksz) {
  for (int i = 0; i < ksz / 2 && i < sz; i++) y[i] = x[i];
  for (int i = 0; i < sz - ksz; i++) {
    uint32_t acc = 0;
    for (int j = 0; j < ksz; j++) acc += k[j] * x[i + j];
    y[ksz / 2 + i] = (uint8_t)(acc >> 8);
  }
  for (int i = sz - ksz + ksz / 2; i < sz; i++) {
    if (i >= 0) y[i] = x[i];
  }
}

void gaussian_taps(double sigma, int ksz, uint8_t *k) {
  vector<double> dk(ksz);
  for (int i = 0; i < ksz; i++) {
    int x = -ksz / 2 + i;
    double v = exp(-.5 * (x / sigma) * (x / sigma));
    dk[i] = v;
  }
  double acc = 0;
  for (int i = 0; i < ksz; i++) acc += dk[i];
  for (int i = 0; i < ksz; i++) dk[i] /= acc;
  for (int i = 0; i < ksz; i++) k[i] = (uint8_t)(dk[i] * 255);
}

void gaussian_blur(Image &im, double sigma, int ksz) {
  if (sigma == 0) return;
  vector<uint8_t> k(ksz);
  gaussian_taps(sigma, ksz, k.data());
  // rows
  vector<uint8_t> tmp(std::max(im.w, im.h)), col(im.h);
  for (int y = 0; y < im.h; y++) {
    convolve1d(&im.buf[(size_t)y * im.w], tmp.data(), im.w, k.data(), ksz);
    memcpy(&im.buf[(size_t)y * im.w], tmp.data(), im.w);
  }
  // columns
  for (int x = 0; x < im.w; x++) {
    for (int y = 0; y < im.h; y++) col[y] = im.buf[(size_t)y * im.w + x];
    convolve1d(col.data(), tmp.data(), im.h, k.data(), ksz);
    for (int y = 0; y < im.h; y++) im.buf[(size_t)y * im.w + x] = tmp[y];
  }
}

// upstream apriltag.c : apriltag_detector_detect, the quad_sigma block
void blur_or_sharpen(Image &quad_im, float quad_sigma) {
  if (quad_sigma == 0) return;
  float sigma = fabsf(quad_sigma);
  int ksz = (int)(4 * sigma);
  if ((ksz & 1) == 0) ksz++;
  if (ksz <= 1) return;
  if (quad_sigma > 0) {
    gaussian_blur(quad_im, sigma, ksz);
  } else {
    Image orig = quad_im;
    gaussian_blur(quad_im, sigma, ksz);
    for (size_t i = 0; i < quad_im.buf.size(); i++) {
      int v = 2 * (int)orig.buf[i] - (int)quad_im.buf[i];
      if (v < 0) v = 0;
      if (v > 255) v = 255;
      quad_im.buf[i] = (uint8_t)v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag_quad_thresh.c : threshold   
// ---------------------------------------------------------------------------------------------
void threshold(Detector &D) {
  const Image &im = D.quad_im;
  int w = im.w, h = im.h;
  const int tilesz = D.prm.tile_size;
  int tw = w / tilesz, th = h / tilesz;
  D.tw = tw;
  D.th = th;
  Image &t = D.thresh;
  t.w = w;
  t.h = h;
  t.buf.assign((size_t)w * h, 0);
  vector<uint8_t> im_max((size_t)tw * th), im_min((size_t)tw * th);
  for (int ty = 0; ty < th; ty++)
    for (int tx = 0; tx < tw; tx++) {
      uint8_t mx = 0, mn = 255;
      for (int dy = 0; dy < tilesz; dy++)
        for (int dx = 0; dx < tilesz; dx++) {
          uint8_t v = im.at(tx * tilesz + dx, ty * tilesz + dy);
          if (v < mn) mn = v;
          if (v > mx) mx = v;
        }
      im_max[(size_t)ty * tw + tx] = mx;
      im_min[(size_t)ty * tw + tx] = mn;
    }
  D.tmax.assign((size_t)tw * th, 0);
  D.tmin.assign((size_t)tw * th, 0);
  for (int ty = 0; ty < th; ty++)
    for (int tx = 0; tx < tw; tx++) {
      uint8_t mx = 0, mn = 255;
      for (int dy = -1; dy <= 1; dy++) {
        if (ty + dy < 0 || ty + dy >= th) continue;
        for (int dx = -1; dx <= 1; dx++) {
          if (tx + dx < 0 || tx + dx >= tw) continue;
          uint8_t m = im_max[(size_t)(ty + dy) * tw + tx + dx];
          if (m > mx) mx = m;
          m = im_min[(size_t)(ty + dy) * tw + tx + dx];
          if (m < mn) mn = m;
        }
      }
      D.tmax[(size_t)ty * tw + tx] = mx;
      D.tmin[(size_t)ty * tw + tx] = mn;
    }
  for (int ty = 0; ty < th; ty++)
    for (int tx = 0; tx < tw; tx++) {
      int mn = D.tmin[(size_t)ty * tw + tx], mx = D.tmax[(size_t)ty * tw + tx];
      if (mx - mn < D.prm.min_white_black_diff) {
        for (int dy = 0; dy < tilesz; dy++)
          for (int dx = 0; dx < tilesz; dx++) t.buf[(size_t)(ty * tilesz + dy) * w + tx * tilesz + dx] = 127;
        continue;
      }
      uint8_t thresh = (uint8_t)(mn + (mx - mn) / 2);
      for (int dy = 0; dy < tilesz; dy++)
        for (int dx = 0; dx < tilesz; dx++) {
          int x = tx * tilesz + dx, y = ty * tilesz + dy;
          t.buf[(size_t)y * w + x] = im.at(x, y) > thresh ? 255 : 0;
        }
    }
  // partial tiles on the right / bottom use the clamped nearest full tile, never 127
  if (tw > 0 && th > 0) {
    for (int y = 0; y < h; y++) {
      int x0 = (y >= th * tilesz) ? 0 : tw * tilesz;
      int ty = y / tilesz;
      if (ty >= th) ty = th - 1;
      for (int x = x0; x < w; x++) {
        int tx = x / tilesz;
        if (tx >= tw) tx = tw - 1;
        int mx = D.tmax[(size_t)ty * tw + tx], mn = D.tmin[(size_t)ty * tw + tx];
        int thresh = mn + (mx - mn) / 2;
        t.buf[(size_t)y * w + x] = im.at(x, y) > thresh ? 255 : 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag_quad_thresh.c : do_unionfind_first_line / do_unionfind_line2 / connected_components
// (SURVEY App. A.3).  The link elisions are restated literally: at the left/right image columns they
// DO change which links exist.
// ---------------------------------------------------------------------------------------------
void connected_components(Detector &D, UnionFind &uf) {
  const Image &im = D.thresh;
  int w = im.w, h = im.h;
  uf.init((uint32_t)w * h);
  const uint8_t *b = im.buf.data();
#define DO_UNIONFIND2(dx, dy)                                    \
  if (b[(size_t)(y + dy) * w + x + dx] == v) uf.connect((uint32_t)(y * w + x), (uint32_t)((y + dy) * w + x + dx));
  {
    int y = 0;
    for (int x = 1; x < w - 1; x++) {
      uint8_t v = b[(size_t)y * w + x];
      if (v == 127) continue;
      DO_UNIONFIND2(-1, 0);
    }
  }
  for (int y = 1; y < h; y++) {
    uint8_t v_m1_m1;
    uint8_t v_0_m1 = b[(size_t)(y - 1) * w];
    uint8_t v_1_m1 = b[(size_t)(y - 1) * w + 1];
    uint8_t v_m1_0;
    uint8_t v = b[(size_t)y * w];
    for (int x = 1; x < w - 1; x++) {
      v_m1_m1 = v_0_m1;
      v_0_m1 = v_1_m1;
      v_1_m1 = b[(size_t)(y - 1) * w + x + 1];
      v_m1_0 = v;
      v = b[(size_t)y * w + x];
      if (v == 127) continue;
      DO_UNIONFIND2(-1, 0);
      if (x == 1 || !((v_m1_0 == v_m1_m1) && (v_m1_m1 == v_0_m1))) {
        DO_UNIONFIND2(0, -1);
      }
      if (v == 255) {
        if (x == 1 || !(v_m1_0 == v_m1_m1 || v_0_m1 == v_m1_m1)) {
          DO_UNIONFIND2(-1, -1);
        }
        if (!(v_0_m1 == v_1_m1)) {
          DO_UNIONFIND2(1, -1);
        }
      }
    }
  }
#undef DO_UNIONFIND2
  // canonicalisation C1: label = min pixel index of the set; size per pixel = set size
  size_t n = (size_t)w * h;
  D.label.assign(n, 0);
  D.lsize.assign(n, 0);
  vector<uint32_t> minidx(n, 0xffffffffu);
  for (size_t i = 0; i < n; i++) {
    uint32_t r = uf.find((uint32_t)i);
    if (minidx[r] == 0xffffffffu) minidx[r] = (uint32_t)i;  // ascending scan => first seen is the minimum
  }
  for (size_t i = 0; i < n; i++) {
    uint32_t r = uf.find((uint32_t)i);
    D.label[i] = minidx[r];
    D.lsize[i] = uf.size[r];
  }
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag_quad_thresh.c : do_gradient_clusters / gradient_clusters  
// v3.4.x semantics including the `connected_last` duplicate suppression.
// ---------------------------------------------------------------------------------------------
void gradient_clusters(Detector &D, std::unordered_map<uint64_t, vector<Pt>> &map) {
  const Image &t = D.thresh;
  int w = t.w, h = t.h;
  const uint8_t *b = t.buf.data();
  int total = 0;
  for (int y = 1; y < h - 1; y++) {
    bool connected_last = false;
    for (int x = 1; x < w - 1; x++) {
      uint8_t v0 = b[(size_t)y * w + x];
      if (v0 == 127) {
        connected_last = false;
        continue;
      }
      uint64_t rep0 = D.label[(size_t)y * w + x];
      if (D.lsize[(size_t)y * w + x] < 25) {
        connected_last = false;
        continue;
      }
      bool connected = false;
#define DO_CONN(dx, dy)                                                                       \
  if (1) {                                                                                    \
    uint8_t v1 = b[(size_t)(y + dy) * w + x + dx];                                            \
    if (v0 + v1 == 255) {                                                                     \
      uint64_t rep1 = D.label[(size_t)(y + dy) * w + x + dx];                                 \
      if (D.lsize[(size_t)(y + dy) * w + x + dx] > 24) {                                      \
        uint64_t clusterid = rep0 < rep1 ? (rep1 << 32) + rep0 : (rep0 << 32) + rep1;         \
        Pt p;                                                                                 \
        p.x = (uint16_t)(2 * x + dx);                                                         \
        p.y = (uint16_t)(2 * y + dy);                                                         \
        p.gx = (int16_t)(dx * ((int)v1 - v0));                                                \
        p.gy = (int16_t)(dy * ((int)v1 - v0));                                                \
        p.slope = 0;                                                                          \
        map[clusterid].push_back(p);                                                          \
        total++;                                                                              \
        connected = true;                                                                     \
      }                                                                                       \
    }                                                                                         \
  }
      // do 4 connectivity
      DO_CONN(1, 0);
      DO_CONN(0, 1);
      // do 8 connectivity
      if (!connected_last) {
        DO_CONN(-1, 1);
      }
      connected = false;
      DO_CONN(1, 1);
      connected_last = connected;
#undef DO_CONN
    }
  }
  D.total_points = total;
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag_quad_thresh.c : fit_line   
// ---------------------------------------------------------------------------------------------
void fit_line(const LineFitPt *lfps, int sz, int i0, int i1, double *lineparm, double *err, double *mse) {
  double Mx, My, Mxx, Myy, Mxy, W;
  int N;
  if (i0 < i1) {
    N = i1 - i0 + 1;
    Mx = lfps[i1].Mx;
    My = lfps[i1].My;
    Mxx = lfps[i1].Mxx;
    Mxy = lfps[i1].Mxy;
    Myy = lfps[i1].Myy;
    W = lfps[i1].W;
    if (i0 > 0) {
      Mx -= lfps[i0 - 1].Mx;
      My -= lfps[i0 - 1].My;
      Mxx -= lfps[i0 - 1].Mxx;
      Mxy -= lfps[i0 - 1].Mxy;
      Myy -= lfps[i0 - 1].Myy;
      W -= lfps[i0 - 1].W;
    }
  } else {
    Mx = lfps[sz - 1].Mx - lfps[i0 - 1].Mx;
    My = lfps[sz - 1].My - lfps[i0 - 1].My;
    Mxx = lfps[sz - 1].Mxx - lfps[i0 - 1].Mxx;
    Mxy = lfps[sz - 1].Mxy - lfps[i0 - 1].Mxy;
    Myy = lfps[sz - 1].Myy - lfps[i0 - 1].Myy;
    W = lfps[sz - 1].W - lfps[i0 - 1].W;
    Mx += lfps[i1].Mx;
    My += lfps[i1].My;
    Mxx += lfps[i1].Mxx;
    Mxy += lfps[i1].Mxy;
    Myy += lfps[i1].Myy;
    W += lfps[i1].W;
    N = sz - i0 + i1 + 1;
  }
  double Ex = Mx / W;
  double Ey = My / W;
  double Cxx = Mxx / W - Ex * Ex;
  double Cxy = Mxy / W - Ex * Ey;
  double Cyy = Myy / W - Ey * Ey;
  double eig_small = 0.5 * (Cxx + Cyy - sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy)));
  if (lineparm) {
    lineparm[0] = Ex;
    lineparm[1] = Ey;
    double eig = 0.5 * (Cxx + Cyy + sqrtf((float)((Cxx - Cyy) * (Cxx - Cyy) + 4 * Cxy * Cxy)));
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
    double length = sqrtf((float)M);
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

// upstream apriltag_quad_thresh.c : compute_lfps   -- sequential prefix sums
void compute_lfps(const vector<Pt> &pts, const Image &im, vector<LineFitPt> &lfps) {
  int sz = (int)pts.size();
  lfps.assign(sz, LineFitPt{0, 0, 0, 0, 0, 0});
  for (int i = 0; i < sz; i++) {
    const Pt *p = &pts[i];
    if (i > 0) lfps[i] = lfps[i - 1];
    double delta = 0.5;
    double x = p->x * .5 + delta;
    double y = p->y * .5 + delta;
    int ix = (int)x, iy = (int)y;
    double W = 1;
    if (ix > 0 && ix + 1 < im.w && iy > 0 && iy + 1 < im.h) {
      int grad_x = im.at(ix + 1, iy) - im.at(ix - 1, iy);
      int grad_y = im.at(ix, iy + 1) - im.at(ix, iy - 1);
      W = sqrt((double)(grad_x * grad_x + grad_y * grad_y)) + 1;
    }
    double fx = x, fy = y;
    lfps[i].Mx += W * fx;
    lfps[i].My += W * fy;
    lfps[i].Mxx += W * fx * fx;
    lfps[i].Mxy += W * fx * fy;
    lfps[i].Myy += W * fy * fy;
    lfps[i].W += W;
  }
}

// upstream apriltag_quad_thresh.c : quad_segment_maxima  
int quad_segment_maxima(const Detector &D, int sz, const LineFitPt *lfps, int indices[4]) {
  int ksz = std::min(20, sz / 12);
  if (ksz < 2) return 0;
  vector<double> errs(sz);
  for (int i = 0; i < sz; i++) fit_line(lfps, sz, (i + sz - ksz) % sz, (i + ksz) % sz, nullptr, &errs[i], nullptr);
  {
    vector<double> y(sz);
    double sigma = 1;
    double cutoff = 0.05;
    int fsz = (int)(sqrt(-log(cutoff) * 2 * sigma * sigma) + 1);
    fsz = 2 * fsz + 1;
    vector<float> f(fsz);
    for (int i = 0; i < fsz; i++) {
      int j = i - fsz / 2;
      f[i] = (float)exp(-j * j / (2 * sigma * sigma));
    }
    for (int iy = 0; iy < sz; iy++) {
      double acc = 0;
      for (int i = 0; i < fsz; i++) acc += errs[(iy + i - fsz / 2 + sz) % sz] * f[i];
      y[iy] = acc;
    }
    errs = y;
  }
  vector<int> maxima;
  vector<double> maxima_errs;
  for (int i = 0; i < sz; i++) {
    if (errs[i] > errs[(i + 1) % sz] && errs[i] > errs[(i + sz - 1) % sz]) {
      maxima.push_back(i);
      maxima_errs.push_back(errs[i]);
    }
  }
  int nmaxima = (int)maxima.size();
  if (nmaxima < 4) return 0;
  int max_nmaxima = D.prm.max_nmaxima;
  if (nmaxima > max_nmaxima) {
    vector<double> copy = maxima_errs;
    std::sort(copy.begin(), copy.end(), [](double a, double b) { return a > b; });
    double maxima_thresh = copy[max_nmaxima];
    int out = 0;
    for (int in = 0; in < nmaxima; in++) {
      if (maxima_errs[in] <= maxima_thresh) continue;
      maxima[out++] = maxima[in];
    }
    nmaxima = out;
  }
  int best_indices[4] = {0, 0, 0, 0};
  double best_error = HUGE_VALF;
  double err01, err12, err23, err30;
  double mse01, mse12, mse23, mse30;
  double params01[4], params12[4], params23[4], params30[4];
  double max_dot = (float)cos((double)D.prm.critical_rad);  // td->qtp.cos_critical_rad is a float field
  for (int m0 = 0; m0 < nmaxima - 3; m0++) {
    int i0 = maxima[m0];
    for (int m1 = m0 + 1; m1 < nmaxima - 2; m1++) {
      int i1 = maxima[m1];
      fit_line(lfps, sz, i0, i1, params01, &err01, &mse01);
      if (mse01 > D.prm.max_line_fit_mse) continue;
      for (int m2 = m1 + 1; m2 < nmaxima - 1; m2++) {
        int i2 = maxima[m2];
        fit_line(lfps, sz, i1, i2, params12, &err12, &mse12);
        if (mse12 > D.prm.max_line_fit_mse) continue;
        double dot = params01[2] * params12[2] + params01[3] * params12[3];
        if (fabs(dot) > max_dot) continue;
        for (int m3 = m2 + 1; m3 < nmaxima; m3++) {
          int i3 = maxima[m3];
          fit_line(lfps, sz, i2, i3, params23, &err23, &mse23);
          if (mse23 > D.prm.max_line_fit_mse) continue;
          fit_line(lfps, sz, i3, i0, params30, &err30, &mse30);
          if (mse30 > D.prm.max_line_fit_mse) continue;
          double err = err01 + err12 + err23 + err30;
          if (err < best_error) {
            best_error = err;
            best_indices[0] = i0;
            best_indices[1] = i1;
            best_indices[2] = i2;
            best_indices[3] = i3;
          }
        }
      }
    }
  }
  if (best_error == HUGE_VALF) return 0;
  for (int i = 0; i < 4; i++) indices[i] = best_indices[i];
  if (best_error / sz < D.prm.max_line_fit_mse) return 1;
  return 0;
}

inline double sq(double v) { return v * v; }

// upstream apriltag_quad_thresh.c : fit_quad  
int fit_quad(const Detector &D, const Image &im, vector<Pt> &cluster, Quad *quad, int tag_width, bool normal_border,
             bool reversed_border) {
  int sz = (int)cluster.size();
  if (sz < 24) return 0;
  uint16_t xmax = cluster[0].x, xmin = cluster[0].x, ymax = cluster[0].y, ymin = cluster[0].y;
  for (int pidx = 1; pidx < sz; pidx++) {
    const Pt *p = &cluster[pidx];
    if (p->x > xmax)
      xmax = p->x;
    else if (p->x < xmin)
      xmin = p->x;
    if (p->y > ymax)
      ymax = p->y;
    else if (p->y < ymin)
      ymin = p->y;
  }
  if ((xmax - xmin) * (ymax - ymin) < tag_width) return 0;
  float cx = (float)((xmin + xmax) * 0.5 + 0.05118);
  float cy = (float)((ymin + ymax) * 0.5 + -0.028581);
  float dot = 0;
  float quadrants[2][2] = {{-1 * (2 << 15), 0}, {2 * (2 << 15), 2 << 15}};
  for (int pidx = 0; pidx < sz; pidx++) {
    Pt *p = &cluster[pidx];
    float dx = p->x - cx;
    float dy = p->y - cy;
    dot += dx * p->gx + dy * p->gy;
    float quadrant = quadrants[dy > 0][dx > 0];
    if (dy < 0) {
      dy = -dy;
      dx = -dx;
    }
    if (dx < 0) {
      float tmp = dx;
      dx = dy;
      dy = -tmp;
    }
    p->slope = quadrant + dy / dx;
  }
  quad->reversed_border = dot < 0;
  if (!reversed_border && quad->reversed_border) return 0;
  if (!normal_border && !quad->reversed_border) return 0;
  // ptsort, canonicalisation C3: ties on slope ordered by (y, x)
  std::sort(cluster.begin(), cluster.end(), [](const Pt &a, const Pt &b) {
    if (a.slope != b.slope) return a.slope < b.slope;
    if (a.y != b.y) return a.y < b.y;
    return a.x < b.x;
  });
  // remove duplicate points (a no-op with connected_last, kept as upstream does)
  {
    int outpos = 1;
    Pt last = cluster[0];
    for (int i = 1; i < sz; i++) {
      Pt p = cluster[i];
      if (p.x != last.x || p.y != last.y) {
        if (i != outpos) cluster[outpos] = p;
        outpos++;
      }
      last = p;
    }
    cluster.resize(outpos);
    sz = outpos;
  }
  if (sz < 24) return 0;
  vector<LineFitPt> lfps;
  compute_lfps(cluster, im, lfps);
  int indices[4];
  if (!quad_segment_maxima(D, sz, lfps.data(), indices)) return 0;
  double lines[4][4];
  for (int i = 0; i < 4; i++) {
    int i0 = indices[i];
    int i1 = indices[(i + 1) & 3];
    double mse;
    fit_line(lfps.data(), sz, i0, i1, lines[i], nullptr, &mse);
    if (mse > D.prm.max_line_fit_mse) return 0;
  }
  for (int i = 0; i < 4; i++) {
    double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
    double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
    double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
    double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
    double det = A00 * A11 - A10 * A01;
    double W00 = A11 / det, W01 = -A01 / det;
    if (fabs(det) < 0.001) return 0;
    double L0 = W00 * B0 + W01 * B1;
    quad->p[i][0] = (float)(lines[i][0] + L0 * A00);
    quad->p[i][1] = (float)(lines[i][1] + L0 * A10);
  }
  {
    double area = 0;
    double length[3], p;
    for (int i = 0; i < 3; i++) {
      int idxa = i;
      int idxb = (i + 1) % 3;
      length[i] = sqrt(sq(quad->p[idxb][0] - quad->p[idxa][0]) + sq(quad->p[idxb][1] - quad->p[idxa][1]));
    }
    p = (length[0] + length[1] + length[2]) / 2;
    area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
    for (int i = 0; i < 3; i++) {
      int idxs[] = {2, 3, 0, 2};
      int idxa = idxs[i];
      int idxb = idxs[i + 1];
      length[i] = sqrt(sq(quad->p[idxb][0] - quad->p[idxa][0]) + sq(quad->p[idxb][1] - quad->p[idxa][1]));
    }
    p = (length[0] + length[1] + length[2]) / 2;
    area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
    if (area < 0.95 * tag_width * tag_width) return 0;
  }
  {
    double ccr = (float)cos((double)D.prm.critical_rad);  // float field upstream
    for (int i = 0; i < 4; i++) {
      int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
      double dx1 = quad->p[i1][0] - quad->p[i0][0];
      double dy1 = quad->p[i1][1] - quad->p[i0][1];
      double dx2 = quad->p[i2][0] - quad->p[i1][0];
      double dy2 = quad->p[i2][1] - quad->p[i1][1];
      double cos_dtheta = (dx1 * dx2 + dy1 * dy2) / sqrt((dx1 * dx1 + dy1 * dy1) * (dx2 * dx2 + dy2 * dy2));
      if ((cos_dtheta > ccr || cos_dtheta < -ccr) || dx1 * dy2 < dy1 * dx2) return 0;
    }
  }
  return 1;
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag.c : refine_edges   
// ---------------------------------------------------------------------------------------------
void refine_edges(const Detector &D, const uint8_t *im, int width, int height, int stride, Quad *quad) {
  double lines[4][4];
  for (int edge = 0; edge < 4; edge++) {
    int a = edge, b = (edge + 1) & 3;
    double nx = quad->p[b][1] - quad->p[a][1];
    double ny = -quad->p[b][0] + quad->p[a][0];
    double mag = sqrt(nx * nx + ny * ny);
    nx /= mag;
    ny /= mag;
    if (quad->reversed_border) {
      nx = -nx;
      ny = -ny;
    }
    int nsamples = std::max(16, (int)(mag / 8));
    double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
    for (int s = 0; s < nsamples; s++) {
      double alpha = (1.0 + s) / (nsamples + 1);
      double x0 = alpha * quad->p[a][0] + (1 - alpha) * quad->p[b][0];
      double y0 = alpha * quad->p[a][1] + (1 - alpha) * quad->p[b][1];
      double Mn = 0;
      double Mcount = 0;
      double range = D.prm.quad_decimate + 1;
      for (double n = -range; n <= range; n += 0.25) {
        double grange = 1;
        int x1 = (int)(x0 + (n + grange) * nx);
        int y1 = (int)(y0 + (n + grange) * ny);
        if (x1 < 0 || x1 >= width || y1 < 0 || y1 >= height) continue;
        int x2 = (int)(x0 + (n - grange) * nx);
        int y2 = (int)(y0 + (n - grange) * ny);
        if (x2 < 0 || x2 >= width || y2 < 0 || y2 >= height) continue;
        int g1 = im[(size_t)y1 * stride + x1];
        int g2 = im[(size_t)y2 * stride + x2];
        if (g1 < g2) continue;
        double weight = (g2 - g1) * (g2 - g1);
        Mn += weight * n;
        Mcount += weight;
      }
      if (Mcount == 0) continue;
      double n0 = Mn / Mcount;
      double bestx = x0 + n0 * nx;
      double besty = y0 + n0 * ny;
      Mx += bestx;
      My += besty;
      Mxx += bestx * bestx;
      Mxy += bestx * besty;
      Myy += besty * besty;
      N++;
    }
    double Ex = Mx / N, Ey = My / N;
    double Cxx = Mxx / N - Ex * Ex;
    double Cxy = Mxy / N - Ex * Ey;
    double Cyy = Myy / N - Ey * Ey;
    double normal_theta = .5 * atan2f((float)(-2 * Cxy), (float)(Cyy - Cxx));
    nx = cosf((float)normal_theta);
    ny = sinf((float)normal_theta);
    lines[edge][0] = Ex;
    lines[edge][1] = Ey;
    lines[edge][2] = nx;
    lines[edge][3] = ny;
  }
  for (int i = 0; i < 4; i++) {
    double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
    double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
    double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
    double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
    double det = A00 * A11 - A10 * A01;
    if (fabs(det) > 0.001) {
      double W00 = A11 / det, W01 = -A01 / det;
      double L0 = W00 * B0 + W01 * B1;
      quad->p[i][0] = (float)(lines[i][0] + L0 * A00);
      quad->p[i][1] = (float)(lines[i][1] + L0 * A10);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// upstream common/homography.c : homography_compute2 ; apriltag.c : quad_update_homographies
// ---------------------------------------------------------------------------------------------
bool homography_compute2(const double c[4][4], double H[9]) {
  double A[] = {
      c[0][0], c[0][1], 1, 0, 0, 0, -c[0][0] * c[0][2], -c[0][1] * c[0][2], c[0][2],
      0, 0, 0, c[0][0], c[0][1], 1, -c[0][0] * c[0][3], -c[0][1] * c[0][3], c[0][3],
      c[1][0], c[1][1], 1, 0, 0, 0, -c[1][0] * c[1][2], -c[1][1] * c[1][2], c[1][2],
      0, 0, 0, c[1][0], c[1][1], 1, -c[1][0] * c[1][3], -c[1][1] * c[1][3], c[1][3],
      c[2][0], c[2][1], 1, 0, 0, 0, -c[2][0] * c[2][2], -c[2][1] * c[2][2], c[2][2],
      0, 0, 0, c[2][0], c[2][1], 1, -c[2][0] * c[2][3], -c[2][1] * c[2][3], c[2][3],
      c[3][0], c[3][1], 1, 0, 0, 0, -c[3][0] * c[3][2], -c[3][1] * c[3][2], c[3][2],
      0, 0, 0, c[3][0], c[3][1], 1, -c[3][0] * c[3][3], -c[3][1] * c[3][3], c[3][3],
  };
  double epsilon = 1e-10;
  for (int col = 0; col < 8; col++) {
    double max_val = 0;
    int max_val_idx = -1;
    for (int row = col; row < 8; row++) {
      double val = fabs(A[row * 9 + col]);
      if (val > max_val) {
        max_val = val;
        max_val_idx = row;
      }
    }
    if (max_val < epsilon) return false;
    if (max_val_idx != col) {
      for (int i = col; i < 9; i++) {
        double tmp = A[col * 9 + i];
        A[col * 9 + i] = A[max_val_idx * 9 + i];
        A[max_val_idx * 9 + i] = tmp;
      }
    }
    for (int i = col + 1; i < 8; i++) {
      double f = A[i * 9 + col] / A[col * 9 + col];
      A[i * 9 + col] = 0;
      for (int j = col + 1; j < 9; j++) A[i * 9 + j] -= f * A[col * 9 + j];
    }
  }
  for (int col = 7; col >= 0; col--) {
    double sum = 0;
    for (int i = col + 1; i < 8; i++) sum += A[col * 9 + i] * A[i * 9 + 8];
    A[col * 9 + 8] = (A[col * 9 + 8] - sum) / A[col * 9 + col];
  }
  H[0] = A[8];
  H[1] = A[17];
  H[2] = A[26];
  H[3] = A[35];
  H[4] = A[44];
  H[5] = A[53];
  H[6] = A[62];
  H[7] = A[71];
  H[8] = 1;
  return true;
}

double det33(const double *m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

int quad_update_homographies(Quad *quad) {
  double corr[4][4];
  for (int i = 0; i < 4; i++) {
    corr[i][0] = (i == 0 || i == 3) ? -1 : 1;
    corr[i][1] = (i == 0 || i == 1) ? -1 : 1;
    corr[i][2] = quad->p[i][0];
    corr[i][3] = quad->p[i][1];
  }
  if (!homography_compute2(corr, quad->H)) return -1;
  // upstream also requires matd_inverse(H) to exist (non-singular)
  if (det33(quad->H) == 0) return -1;
  return 0;
}

inline void homography_project(const double *H, double x, double y, double *ox, double *oy) {
  double xx = H[0] * x + H[1] * y + H[2];
  double yy = H[3] * x + H[4] * y + H[5];
  double zz = H[6] * x + H[7] * y + H[8];
  *ox = xx / zz;
  *oy = yy / zz;
}

// upstream common/math_util / apriltag.c : graymodel + mat33_chol / mat33_lower_tri_inv / mat33_sym_solve
struct GrayModel {
  double A[3][3];
  double B[3];
  double C[3];
};
void graymodel_add(GrayModel *gm, double x, double y, double gray) {
  gm->A[0][0] += x * x;
  gm->A[0][1] += x * y;
  gm->A[0][2] += x;
  gm->A[1][1] += y * y;
  gm->A[1][2] += y;
  gm->A[2][2] += 1;
  gm->B[0] += x * gray;
  gm->B[1] += y * gray;
  gm->B[2] += gray;
}
void mat33_sym_solve(const double *A, const double *B, double *R) {
  double L[9];
  L[0] = sqrt(A[0]);
  L[3] = A[1] / L[0];
  L[6] = A[2] / L[0];
  L[4] = sqrt(A[4] - L[3] * L[3]);
  L[7] = (A[5] - L[3] * L[6]) / L[4];
  L[8] = sqrt(A[8] - L[6] * L[6] - L[7] * L[7]);
  L[1] = 0;
  L[2] = 0;
  L[5] = 0;
  double M[9];
  M[0] = 1 / L[0];
  M[3] = -L[3] * M[0] / L[4];
  M[4] = 1 / L[4];
  M[6] = (-L[6] * M[0] - L[7] * M[3]) / L[8];
  M[7] = -L[7] * M[4] / L[8];
  M[8] = 1 / L[8];
  double tmp[3];
  tmp[0] = M[0] * B[0];
  tmp[1] = M[3] * B[0] + M[4] * B[1];
  tmp[2] = M[6] * B[0] + M[7] * B[1] + M[8] * B[2];
  R[0] = M[0] * tmp[0] + M[3] * tmp[1] + M[6] * tmp[2];
  R[1] = M[4] * tmp[1] + M[7] * tmp[2];
  R[2] = M[8] * tmp[2];
}
void graymodel_solve(GrayModel *gm) { mat33_sym_solve((double *)gm->A, gm->B, gm->C); }
double graymodel_interpolate(const GrayModel *gm, double x, double y) { return gm->C[0] * x + gm->C[1] * y + gm->C[2]; }

// upstream apriltag.c : value_for_pixel (bilinear, pixel centres at +0.5)
double value_for_pixel(const uint8_t *im, int width, int height, int stride, double px, double py) {
  int x1 = (int)floor(px - 0.5);
  int x2 = (int)ceil(px - 0.5);
  double x = px - 0.5 - x1;
  int y1 = (int)floor(py - 0.5);
  int y2 = (int)ceil(py - 0.5);
  double y = py - 0.5 - y1;
  if (x1 < 0 || x2 >= width || y1 < 0 || y2 >= height) return -1;
  return im[(size_t)y1 * stride + x1] * (1 - x) * (1 - y) + im[(size_t)y1 * stride + x2] * x * (1 - y) +
         im[(size_t)y2 * stride + x1] * (1 - x) * y + im[(size_t)y2 * stride + x2] * x * y;
}

// upstream apriltag.c : sharpen
void sharpen(double decode_sharpening, double *values, int size) {
  vector<double> sharpened((size_t)size * size);
  double kernel[9] = {0, -1, 0, -1, 4, -1, 0, -1, 0};
  for (int y = 0; y < size; y++)
    for (int x = 0; x < size; x++) {
      sharpened[y * size + x] = 0;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
          if ((y + i - 1) < 0 || (y + i - 1) > size - 1 || (x + j - 1) < 0 || (x + j - 1) > size - 1) continue;
          sharpened[y * size + x] += values[(y + i - 1) * size + (x + j - 1)] * kernel[i * 3 + j];
        }
    }
  for (int y = 0; y < size; y++)
    for (int x = 0; x < size; x++) values[y * size + x] = values[y * size + x] + decode_sharpening * sharpened[y * size + x];
}

// upstream apriltag.c : rotate90
uint64_t rotate90(uint64_t w, int numBits) {
  int p = numBits;
  uint64_t l = 0;
  if (numBits % 4 == 1) {
    p = numBits - 1;
    l = 1;
  }
  w = ((w >> l) << (p / 4 + l)) | (w >> (3 * p / 4 + l) << l) | (w & l);
  w &= ((uint64_t(1) << numBits) - 1);
  return w;
}

struct DecodeEntry {
  int id, hamming, rotation;
};

// upstream apriltag.c : quick_decode_codeword.  Upstream looks `rcode` up in a hash of every code with
// <= maxhamming bit flips, rotation by rotation, first hit wins.  A direct popcount scan is the same
// function because the families' minimum distance (>= 2*maxhamming+1) makes the hit unique.
void quick_decode_codeword(const Family &fam, int maxhamming, uint64_t rcode, DecodeEntry *entry) {
  for (int ridx = 0; ridx < 4; ridx++) {
    for (int k = 0; k < fam.ncodes; k++) {
      int hd = __builtin_popcountll(rcode ^ (uint64_t)fam.codes[k]);
      if (hd <= maxhamming) {
        entry->id = k;
        entry->hamming = hd;
        entry->rotation = ridx;
        return;
      }
    }
    rcode = rotate90(rcode, fam.nbits);
  }
  entry->id = 65535;
  entry->hamming = 255;
  entry->rotation = 0;
}

// upstream apriltag.c : quad_decode  
float quad_decode(const Detector &D, const Family &family, const uint8_t *im, int width, int height, int stride,
                  const Quad *quad, DecodeEntry *entry) {
  float wab = (float)family.width_at_border;
  float patterns[] = {
      -0.5f, 0.5f, 0, 1, 1,        // left white column
      0.5f, 0.5f, 0, 1, 0,         // left black column
      wab + 0.5f, .5f, 0, 1, 1,    // right white column
      wab - 0.5f, .5f, 0, 1, 0,    // right black column
      0.5f, -0.5f, 1, 0, 1,        // top white row
      0.5f, 0.5f, 1, 0, 0,         // top black row
      0.5f, wab + 0.5f, 1, 0, 1,   // bottom white row
      0.5f, wab - 0.5f, 1, 0, 0,   // bottom black row
  };
  GrayModel whitemodel, blackmodel;
  memset(&whitemodel, 0, sizeof(whitemodel));
  memset(&blackmodel, 0, sizeof(blackmodel));
  for (size_t pattern_idx = 0; pattern_idx < sizeof(patterns) / (5 * sizeof(float)); pattern_idx++) {
    float *pattern = &patterns[pattern_idx * 5];
    int is_white = (int)pattern[4];
    for (int i = 0; i < family.width_at_border; i++) {
      double tagx01 = (pattern[0] + i * pattern[2]) / (family.width_at_border);
      double tagy01 = (pattern[1] + i * pattern[3]) / (family.width_at_border);
      double tagx = 2 * (tagx01 - 0.5);
      double tagy = 2 * (tagy01 - 0.5);
      double px, py;
      homography_project(quad->H, tagx, tagy, &px, &py);
      int ix = (int)px;
      int iy = (int)py;
      if (ix < 0 || iy < 0 || ix >= width || iy >= height) continue;
      int v = im[(size_t)iy * stride + ix];
      if (is_white)
        graymodel_add(&whitemodel, tagx, tagy, v);
      else
        graymodel_add(&blackmodel, tagx, tagy, v);
    }
  }
  if (family.width_at_border > 1) {
    graymodel_solve(&whitemodel);
    graymodel_solve(&blackmodel);
  } else {
    graymodel_solve(&whitemodel);
    blackmodel.C[0] = 0;
    blackmodel.C[1] = 0;
    blackmodel.C[2] = blackmodel.B[2] / 4;
  }
  if ((graymodel_interpolate(&whitemodel, 0, 0) - graymodel_interpolate(&blackmodel, 0, 0) < 0) !=
      family.reversed_border)
    return -1;
  float black_score = 0, white_score = 0;
  float black_score_count = 1, white_score_count = 1;
  int tw = family.total_width;
  vector<double> values((size_t)tw * tw, 0.0);
  int min_coord = (family.width_at_border - family.total_width) / 2;
  for (int i = 0; i < family.nbits; i++) {
    int bity = family.bit_y[i];
    int bitx = family.bit_x[i];
    double tagx01 = (bitx + 0.5) / (family.width_at_border);
    double tagy01 = (bity + 0.5) / (family.width_at_border);
    double tagx = 2 * (tagx01 - 0.5);
    double tagy = 2 * (tagy01 - 0.5);
    double px, py;
    homography_project(quad->H, tagx, tagy, &px, &py);
    double v = value_for_pixel(im, width, height, stride, px, py);
    if (v == -1) continue;
    double thresh = (graymodel_interpolate(&blackmodel, tagx, tagy) + graymodel_interpolate(&whitemodel, tagx, tagy)) / 2.0;
    values[(size_t)tw * (bity - min_coord) + bitx - min_coord] = v - thresh;
  }
  sharpen(D.prm.decode_sharpening, values.data(), tw);
  uint64_t rcode = 0;
  for (int i = 0; i < family.nbits; i++) {
    int bity = family.bit_y[i];
    int bitx = family.bit_x[i];
    rcode = (rcode << 1);
    double v = values[(size_t)(bity - min_coord) * tw + bitx - min_coord];
    if (v > 0) {
      white_score = (float)((double)white_score + v);  // float += double: summed in double, rounded to float
      white_score_count++;
      rcode |= 1;
    } else {
      black_score = (float)((double)black_score - v);
      black_score_count++;
    }
  }
  quick_decode_codeword(family, D.prm.max_hamming, rcode, entry);
  return fminf(white_score / white_score_count, black_score / black_score_count);
}

// ---------------------------------------------------------------------------------------------
// upstream common/g2d.c : g2d_polygon_overlaps_polygon (used by reconcile)
// ---------------------------------------------------------------------------------------------
struct Seg {
  double p[2], u[2];  // origin, unit direction; p1 = end point
  double p1[2];
};
bool seg_intersect(const double a0[2], const double a1[2], const double b0[2], const double b1[2]) {
  // g2d_line_segment_intersect_segment: intersect the two infinite lines, then check the point lies
  // within both segments (by projection on each segment's direction).
  double ua[2] = {a1[0] - a0[0], a1[1] - a0[1]};
  double ub[2] = {b1[0] - b0[0], b1[1] - b0[1]};
  double la = sqrt(ua[0] * ua[0] + ua[1] * ua[1]), lb = sqrt(ub[0] * ub[0] + ub[1] * ub[1]);
  ua[0] /= la;
  ua[1] /= la;
  ub[0] /= lb;
  ub[1] /= lb;
  // g2d_line_intersect_line
  double m00 = ua[0], m01 = -ub[0], m10 = ua[1], m11 = -ub[1];
  double det = m00 * m11 - m01 * m10;
  if (fabs(det) < 0.00000001) return false;
  double i00 = m11 / det, i01 = -m01 / det;
  double b00 = b0[0] - a0[0], b10 = b0[1] - a0[1];
  double x00 = i00 * b00 + i01 * b10;
  double px = ua[0] * x00 + a0[0], py = ua[1] * x00 + a0[1];
  // within segment a?
  double ta = (px - a0[0]) * ua[0] + (py - a0[1]) * ua[1];
  double tb = (px - b0[0]) * ub[0] + (py - b0[1]) * ub[1];
  double a_lo = 0, a_hi = (a1[0] - a0[0]) * ua[0] + (a1[1] - a0[1]) * ua[1];
  double b_lo = 0, b_hi = (b1[0] - b0[0]) * ub[0] + (b1[1] - b0[1]) * ub[1];
  if (ta < std::min(a_lo, a_hi) || ta > std::max(a_lo, a_hi)) return false;
  if (tb < std::min(b_lo, b_hi) || tb > std::max(b_lo, b_hi)) return false;
  return true;
}
// g2d_polygon_contains_point: winding by accumulated signed angle quadrant changes (upstream uses a
// quadrant-crossing count); for the convex quads here an orientation test is the same predicate.
bool poly_contains_point(const double poly[4][2], const double q[2]) {
  int pos = 0, neg = 0;
  for (int i = 0; i < 4; i++) {
    const double *a = poly[i], *b = poly[(i + 1) & 3];
    double cr = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0]);
    if (cr > 0) pos++;
    if (cr < 0) neg++;
  }
  return !(pos > 0 && neg > 0);
}
bool polygon_overlaps_polygon(const double a[4][2], const double b[4][2]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      if (seg_intersect(a[i], a[(i + 1) & 3], b[j], b[(j + 1) & 3])) return true;
  // no edge crossings: either disjoint or one contains the other
  if (poly_contains_point(a, b[0])) return true;
  if (poly_contains_point(b, a[0])) return true;
  return false;
}

int prefer_smaller(int pref, double q0, double q1) {
  if (pref) return pref;
  if (q0 < q1) return -1;
  if (q1 < q0) return 1;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// apriltag_pose.c  (SURVEY App. A.9).  3x3 helpers are plain row-major arrays.
// ---------------------------------------------------------------------------------------------
void mm33(const double *A, const double *B, double *C) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i * 3 + j] = A[i * 3 + 0] * B[0 * 3 + j] + A[i * 3 + 1] * B[1 * 3 + j] + A[i * 3 + 2] * B[2 * 3 + j];
  memcpy(C, t, sizeof(t));
}
void mv33(const double *A, const double *v, double *o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[i * 3 + 0] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
  memcpy(o, t, sizeof(t));
}
void tr33(const double *A, double *T) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[j * 3 + i] = A[i * 3 + j];
  memcpy(T, t, sizeof(t));
}
bool inv33(const double *m, double *o) {
  double d = det33(m);
  if (d == 0) return false;
  double id = 1.0 / d;
  double t[9];
  t[0] = (m[4] * m[8] - m[5] * m[7]) * id;
  t[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  t[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  t[3] = (m[5] * m[6] - m[3] * m[8]) * id;
  t[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  t[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  t[6] = (m[3] * m[7] - m[4] * m[6]) * id;
  t[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  t[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  memcpy(o, t, sizeof(t));
  return true;
}

// Polar factor U*V' of a 3x3 matrix (what upstream gets from matd_svd then "M*M'").  Implemented as a
// cyclic Jacobi eigen-decomposition of A'A; singular directions with (near-)zero singular value are
// completed by cross products, which is all the planar (rank-2) case of orthogonal_iteration needs.
void polar_UVt(const double *A, double *R) {
  double At[9], S[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  tr33(A, At);
  mm33(At, A, S);
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = fabs(S[1]) + fabs(S[2]) + fabs(S[5]);
    if (off <= 1e-32 * (fabs(S[0]) + fabs(S[4]) + fabs(S[8]))) break;  // converged far below double epsilon
    for (int pi = 0; pi < 3; pi++) {
      int p = pi == 2 ? 0 : pi, q = pi == 0 ? 1 : 2;  // (0,1), (1,2), (0,2)
      double apq = S[p * 3 + q];
      if (apq == 0) continue;
      double app = S[p * 3 + p], aqq = S[q * 3 + q];
      double theta = (aqq - app) / (2 * apq);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
      double c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < 3; k++) {  // S = S * J
        double skp = S[k * 3 + p], skq = S[k * 3 + q];
        S[k * 3 + p] = c * skp - s * skq;
        S[k * 3 + q] = s * skp + c * skq;
      }
      for (int k = 0; k < 3; k++) {  // S = J' * S
        double spk = S[p * 3 + k], sqk = S[q * 3 + k];
        S[p * 3 + k] = c * spk - s * sqk;
        S[q * 3 + k] = s * spk + c * sqk;
      }
      for (int k = 0; k < 3; k++) {
        double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq;
        V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  // order eigenvalues descending
  int idx[3] = {0, 1, 2};
  double ev[3] = {S[0], S[4], S[8]};
  for (int i = 0; i < 3; i++)
    for (int j = i + 1; j < 3; j++)
      if (ev[idx[j]] > ev[idx[i]]) std::swap(idx[i], idx[j]);
  double Vs[9], U[9];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) Vs[r * 3 + c] = V[r * 3 + idx[c]];
  double smax = sqrt(std::max(ev[idx[0]], 0.0));
  int rank = 0;
  for (int c = 0; c < 3; c++) {
    double s = sqrt(std::max(ev[idx[c]], 0.0));
    if (s > 1e-12 * smax && s > 0) {
      double v[3] = {Vs[0 * 3 + c], Vs[1 * 3 + c], Vs[2 * 3 + c]}, u[3];
      mv33(A, v, u);
      for (int r = 0; r < 3; r++) U[r * 3 + c] = u[r] / s;
      rank = c + 1;
    } else {
      break;
    }
  }
  auto cross_col = [](double *M, int a, int b, int o) {
    double ax = M[0 * 3 + a], ay = M[1 * 3 + a], az = M[2 * 3 + a];
    double bx = M[0 * 3 + b], by = M[1 * 3 + b], bz = M[2 * 3 + b];
    M[0 * 3 + o] = ay * bz - az * by;
    M[1 * 3 + o] = az * bx - ax * bz;
    M[2 * 3 + o] = ax * by - ay * bx;
  };
  if (rank == 2) {
    cross_col(U, 0, 1, 2);
  } else if (rank < 2) {
    // degenerate input; return identity-like completion
    for (int i = 0; i < 9; i++) U[i] = (i % 4 == 0) ? 1 : 0;
    for (int i = 0; i < 9; i++) Vs[i] = (i % 4 == 0) ? 1 : 0;
  }
  double Vt[9];
  tr33(Vs, Vt);
  mm33(U, Vt, R);
}

// upstream common/homography.c : homography_to_pose
void homography_to_pose(const double *H, double fx, double fy, double cx, double cy, double R[9], double T[3]) {
  double R20 = H[6];
  double R21 = H[7];
  double TZ = H[8];
  double R00 = (H[0] - cx * R20) / fx;
  double R01 = (H[1] - cx * R21) / fx;
  double TX = (H[2] - cx * TZ) / fx;
  double R10 = (H[3] - cy * R20) / fy;
  double R11 = (H[4] - cy * R21) / fy;
  double TY = (H[5] - cy * TZ) / fy;
  double length1 = sqrtf((float)(R00 * R00 + R10 * R10 + R20 * R20));
  double length2 = sqrtf((float)(R01 * R01 + R11 * R11 + R21 * R21));
  double s = 1.0 / sqrtf((float)(length1 * length2));
  if (TZ > 0) s *= -1;
  R20 *= s;
  R21 *= s;
  TZ *= s;
  R00 *= s;
  R01 *= s;
  TX *= s;
  R10 *= s;
  R11 *= s;
  TY *= s;
  double R02 = R10 * R21 - R20 * R11;
  double R12 = R20 * R01 - R00 * R21;
  double R22 = R00 * R11 - R10 * R01;
  double M[9] = {R00, R01, R02, R10, R11, R12, R20, R21, R22};
  polar_UVt(M, R);
  T[0] = TX;
  T[1] = TY;
  T[2] = TZ;
}

void calculate_F(const double v[3], double F[9]) {
  double n = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) F[i * 3 + j] = v[i] * v[j] / n;
}

// upstream apriltag_pose.c : orthogonal_iteration (Lu, Hager, Mjolsness 2000)
double orthogonal_iteration(const double v[4][3], const double p[4][3], double t[3], double R[9], int n_points, int n_steps) {
  double p_mean[3] = {0, 0, 0};
  for (int i = 0; i < n_points; i++)
    for (int k = 0; k < 3; k++) p_mean[k] += p[i][k];
  for (int k = 0; k < 3; k++) p_mean[k] *= 1.0 / n_points;
  double p_res[4][3];
  for (int i = 0; i < n_points; i++)
    for (int k = 0; k < 3; k++) p_res[i][k] = p[i][k] - p_mean[k];
  double F[4][9], avg_F[9] = {0};
  for (int i = 0; i < n_points; i++) {
    calculate_F(v[i], F[i]);
    for (int k = 0; k < 9; k++) avg_F[k] += F[i][k];
  }
  for (int k = 0; k < 9; k++) avg_F[k] *= 1.0 / n_points;
  double M1[9], M1_inv[9];
  for (int k = 0; k < 9; k++) M1[k] = ((k % 4 == 0) ? 1.0 : 0.0) - avg_F[k];
  inv33(M1, M1_inv);
  double prev_error = HUGE_VAL;
  for (int it = 0; it < n_steps; it++) {
    double M2[3] = {0, 0, 0};
    for (int j = 0; j < n_points; j++) {
      double FmI[9], Rp[3], u[3];
      for (int k = 0; k < 9; k++) FmI[k] = F[j][k] - ((k % 4 == 0) ? 1.0 : 0.0);
      mv33(R, p[j], Rp);
      mv33(FmI, Rp, u);
      for (int k = 0; k < 3; k++) M2[k] += u[k];
    }
    for (int k = 0; k < 3; k++) M2[k] *= 1.0 / n_points;
    mv33(M1_inv, M2, t);
    double q[4][3], q_mean[3] = {0, 0, 0};
    for (int j = 0; j < n_points; j++) {
      double Rp[3];
      mv33(R, p[j], Rp);
      for (int k = 0; k < 3; k++) Rp[k] += t[k];
      mv33(F[j], Rp, q[j]);
      for (int k = 0; k < 3; k++) q_mean[k] += q[j][k];
    }
    for (int k = 0; k < 3; k++) q_mean[k] *= 1.0 / n_points;
    double M3[9] = {0};
    for (int j = 0; j < n_points; j++)
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) M3[a * 3 + b] += (q[j][a] - q_mean[a]) * p_res[j][b];
    polar_UVt(M3, R);
    if (det33(R) < 0) {
      R[2] *= -1;
      R[5] *= -1;
      R[8] *= -1;
    }
    double error = 0;
    for (int j = 0; j < 4; j++) {
      double ImF[9], Rp[3], e[3];
      for (int k = 0; k < 9; k++) ImF[k] = ((k % 4 == 0) ? 1.0 : 0.0) - F[j][k];
      mv33(R, p[j], Rp);
      for (int k = 0; k < 3; k++) Rp[k] += t[k];
      mv33(ImF, Rp, e);
      error += e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
    }
    prev_error = error;
  }
  return prev_error;
}

double polyval(const double *p, int degree, double x) {
  // upstream: ret += p[i]*pow(x, i); restated with repeated multiplication so the CUDA pose kernel can
  // reproduce it bit for bit (pow() is not reproducible across libms; the difference is <= 1 ulp per term)
  double ret = 0, xp = 1;
  for (int i = 0; i <= degree; i++) {
    ret += p[i] * xp;
    xp *= x;
  }
  return ret;
}

// upstream apriltag_pose.c : solve_poly_approx
void solve_poly_approx(const double *p, int degree, double *roots, int *n_roots) {
  static const int MAX_ROOT = 1000;
  if (degree == 1) {
    if (fabs(p[0]) > MAX_ROOT * fabs(p[1])) {
      *n_roots = 0;
    } else {
      roots[0] = -p[0] / p[1];
      *n_roots = 1;
    }
    return;
  }
  double p_der[8];
  for (int i = 0; i < degree; i++) p_der[i] = (i + 1) * p[i + 1];
  double der_roots[8];
  int n_der_roots;
  solve_poly_approx(p_der, degree - 1, der_roots, &n_der_roots);
  *n_roots = 0;
  for (int i = 0; i <= n_der_roots; i++) {
    double mn = (i == 0) ? -MAX_ROOT : der_roots[i - 1];
    double mx = (i == n_der_roots) ? MAX_ROOT : der_roots[i];
    if (polyval(p, degree, mn) * polyval(p, degree, mx) < 0) {
      double lower, upper;
      if (polyval(p, degree, mn) < polyval(p, degree, mx)) {
        lower = mn;
        upper = mx;
      } else {
        lower = mx;
        upper = mn;
      }
      double root = 0.5 * (lower + upper);
      double dx_old = upper - lower;
      double dx = dx_old;
      double f = polyval(p, degree, root);
      double df = polyval(p_der, degree - 1, root);
      for (int j = 0; j < 100; j++) {
        if (((f + df * (upper - root)) * (f + df * (lower - root)) > 0) || (fabs(2 * f) > fabs(dx_old * df))) {
          dx_old = dx;
          dx = 0.5 * (upper - lower);
          root = lower + dx;
        } else {
          dx_old = dx;
          dx = -f / df;
          root += dx;
        }
        if (root == upper || root == lower) break;
        f = polyval(p, degree, root);
        df = polyval(p_der, degree - 1, root);
        if (f > 0)
          upper = root;
        else
          lower = root;
      }
      roots[(*n_roots)++] = root;
    } else if (polyval(p, degree, mx) == 0) {
      roots[(*n_roots)++] = mx;
    }
  }
}

// upstream apriltag_pose.c : fix_pose_ambiguities (Schweighofer & Pinz 2006).  Returns true and R2 if a
// second local minimum exists.
bool fix_pose_ambiguities(const double v[4][3], const double p[4][3], const double t[3], const double R[9], int n_points,
                          double R2[9]) {
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  // 1. R_t
  double tn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  double Rt3[3] = {t[0] / tn, t[1] / tn, t[2] / tn};
  double ex[3] = {1, 0, 0};
  double d = ex[0] * Rt3[0] + ex[1] * Rt3[1] + ex[2] * Rt3[2];
  double tmp[3] = {ex[0] - d * Rt3[0], ex[1] - d * Rt3[1], ex[2] - d * Rt3[2]};
  double tmn = sqrt(tmp[0] * tmp[0] + tmp[1] * tmp[1] + tmp[2] * tmp[2]);
  double Rt1[3] = {tmp[0] / tmn, tmp[1] / tmn, tmp[2] / tmn};
  double Rt2[3] = {Rt3[1] * Rt1[2] - Rt3[2] * Rt1[1], Rt3[2] * Rt1[0] - Rt3[0] * Rt1[2], Rt3[0] * Rt1[1] - Rt3[1] * Rt1[0]};
  double R_t[9] = {Rt1[0], Rt1[1], Rt1[2], Rt2[0], Rt2[1], Rt2[2], Rt3[0], Rt3[1], Rt3[2]};
  // 2. R_z
  double R_1_prime[9];
  mm33(R_t, R, R_1_prime);
  double r31 = R_1_prime[6];
  double r32 = R_1_prime[7];
  double hypotenuse = sqrt(r31 * r31 + r32 * r32);
  if (hypotenuse < 1e-100) {
    r31 = 1;
    r32 = 0;
    hypotenuse = 1;
  }
  double R_z[9] = {r31 / hypotenuse, -r32 / hypotenuse, 0, r32 / hypotenuse, r31 / hypotenuse, 0, 0, 0, 1};
  // 3. parameters of Eos
  double R_trans[9];
  mm33(R_1_prime, R_z, R_trans);
  double sin_gamma = -R_trans[1];
  double cos_gamma = R_trans[4];
  double R_gamma[9] = {cos_gamma, -sin_gamma, 0, sin_gamma, cos_gamma, 0, 0, 0, 1};
  double sin_beta = -R_trans[6];
  double cos_beta = R_trans[8];
  double t_initial = atan2(sin_beta, cos_beta);
  double v_trans[4][3], p_trans[4][3], F_trans[4][9], avg_F_trans[9] = {0};
  double R_zT[9];
  tr33(R_z, R_zT);
  for (int i = 0; i < n_points; i++) {
    mv33(R_zT, p[i], p_trans[i]);
    mv33(R_t, v[i], v_trans[i]);
    calculate_F(v_trans[i], F_trans[i]);
    for (int k = 0; k < 9; k++) avg_F_trans[k] += F_trans[i][k];
  }
  for (int k = 0; k < 9; k++) avg_F_trans[k] *= 1.0 / n_points;
  double G[9], ImA[9];
  for (int k = 0; k < 9; k++) ImA[k] = I3[k] - avg_F_trans[k];
  inv33(ImA, G);
  for (int k = 0; k < 9; k++) G[k] *= 1.0 / n_points;
  const double M1[9] = {0, 0, 2, 0, 0, 0, -2, 0, 0};
  const double M2[9] = {-1, 0, 0, 0, 1, 0, 0, 0, -1};
  double b0[3] = {0, 0, 0}, b1[3] = {0, 0, 0}, b2[3] = {0, 0, 0};
  double RgM1[9], RgM2[9];
  mm33(R_gamma, M1, RgM1);
  mm33(R_gamma, M2, RgM2);
  for (int i = 0; i < n_points; i++) {
    double FmI[9], a[3], o[3];
    for (int k = 0; k < 9; k++) FmI[k] = F_trans[i][k] - I3[k];
    mv33(R_gamma, p_trans[i], a);
    mv33(FmI, a, o);
    for (int k = 0; k < 3; k++) b0[k] += o[k];
    mv33(RgM1, p_trans[i], a);
    mv33(FmI, a, o);
    for (int k = 0; k < 3; k++) b1[k] += o[k];
    mv33(RgM2, p_trans[i], a);
    mv33(FmI, a, o);
    for (int k = 0; k < 3; k++) b2[k] += o[k];
  }
  double b0_[3], b1_[3], b2_[3];
  mv33(G, b0, b0_);
  mv33(G, b1, b1_);
  mv33(G, b2, b2_);
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  for (int i = 0; i < n_points; i++) {
    double ImF[9], a[3], c0[3], c1[3], c2[3];
    for (int k = 0; k < 9; k++) ImF[k] = I3[k] - F_trans[i][k];
    mv33(R_gamma, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b0_[k];
    mv33(ImF, a, c0);
    mv33(RgM1, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b1_[k];
    mv33(ImF, a, c1);
    mv33(RgM2, p_trans[i], a);
    for (int k = 0; k < 3; k++) a[k] += b2_[k];
    mv33(ImF, a, c2);
    auto dot3 = [](const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
    a0 += dot3(c0, c0);
    a1 += 2 * dot3(c0, c1);
    a2 += dot3(c1, c1) + 2 * dot3(c0, c2);
    a3 += 2 * dot3(c1, c2);
    a4 += dot3(c2, c2);
  }
  // 4. minima of Eos
  double p0 = a1;
  double p1 = 2 * a2 - 4 * a0;
  double p2 = 3 * a3 - 3 * a1;
  double p3 = 4 * a4 - 2 * a2;
  double p4 = -a3;
  double roots[4];
  int n_roots;
  double poly[5] = {p0, p1, p2, p3, p4};
  solve_poly_approx(poly, 4, roots, &n_roots);
  double minima[4];
  int n_minima = 0;
  for (int i = 0; i < n_roots; i++) {
    double t1 = roots[i];
    double t2 = t1 * t1;
    double t3 = t1 * t2;
    double t4 = t1 * t3;
    double t5 = t1 * t4;
    if (a2 - 2 * a0 + (3 * a3 - 6 * a1) * t1 + (6 * a4 - 8 * a2 + 10 * a0) * t2 + (-8 * a3 + 6 * a1) * t3 +
            (-6 * a4 + 3 * a2) * t4 + a3 * t5 >=
        0) {
      double tt = 2 * atan(roots[i]);
      if (fabs(tt - t_initial) > 0.1) minima[n_minima++] = roots[i];
    }
  }
  // 5. pose for the minimum
  if (n_minima == 1) {
    double tt = minima[0];
    double R_beta[9];
    for (int k = 0; k < 9; k++) R_beta[k] = M2[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= tt;
    for (int k = 0; k < 9; k++) R_beta[k] += M1[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= tt;
    for (int k = 0; k < 9; k++) R_beta[k] += I3[k];
    for (int k = 0; k < 9; k++) R_beta[k] *= 1 / (1 + tt * tt);
    double R_tT[9], A[9], B[9];
    tr33(R_t, R_tT);
    mm33(R_tT, R_gamma, A);
    mm33(A, R_beta, B);
    mm33(B, R_zT, R2);
    return true;
  }
  return false;
}

// upstream apriltag_pose.c : estimate_pose_for_tag_homography + estimate_tag_pose_orthogonal_iteration + estimate_tag_pose
void estimate_pose(const ato_detection_t *det, double fx, double fy, double cx, double cy, double tagsize, ato_pose_t *best,
                   ato_pose_t *o1, ato_pose_t *o2) {
  double scale = tagsize / 2.0;
  double p[4][3] = {{-scale, scale, 0}, {scale, scale, 0}, {scale, -scale, 0}, {-scale, -scale, 0}};
  double v[4][3];
  for (int i = 0; i < 4; i++) {
    v[i][0] = (det->p[i][0] - cx) / fx;
    v[i][1] = (det->p[i][1] - cy) / fy;
    v[i][2] = 1;
  }
  ato_pose_t s1, s2;
  {
    double R[9], T[3];
    homography_to_pose(det->H, -fx, fy, cx, cy, R, T);
    T[0] *= scale;
    T[1] *= scale;
    T[2] *= scale;
    // fix = diag(1,-1,-1,1) * M_H
    for (int j = 0; j < 3; j++) {
      s1.R[0 * 3 + j] = R[0 * 3 + j];
      s1.R[1 * 3 + j] = -R[1 * 3 + j];
      s1.R[2 * 3 + j] = -R[2 * 3 + j];
    }
    s1.t[0] = T[0];
    s1.t[1] = -T[1];
    s1.t[2] = -T[2];
  }
  s1.err = orthogonal_iteration(v, p, s1.t, s1.R, 4, 50);
  memset(&s2, 0, sizeof(s2));
  if (fix_pose_ambiguities(v, p, s1.t, s1.R, 4, s2.R)) {
    s2.t[0] = s2.t[1] = s2.t[2] = 0;
    s2.err = orthogonal_iteration(v, p, s2.t, s2.R, 4, 50);
  } else {
    s2.err = HUGE_VAL;
  }
  if (o1) *o1 = s1;
  if (o2) *o2 = s2;
  if (best) *best = (s1.err <= s2.err) ? s1 : s2;
}

// ---------------------------------------------------------------------------------------------
// upstream apriltag.c : apriltag_detector_detect
// ---------------------------------------------------------------------------------------------
int detect(Detector &D, const uint8_t *im_orig, int width, int height, int stride, ato_detection_t *out, int max_out) {
  double t0 = now_s(), t1;
  memset(&D.times, 0, sizeof(D.times));
  // Step 1. decimate / blur
  if (D.prm.quad_decimate > 1) {
    D.quad_im = decimate(im_orig, width, height, stride, D.prm.quad_decimate);
  } else {
    D.quad_im.w = width;
    D.quad_im.h = height;
    D.quad_im.buf.resize((size_t)width * height);
    for (int y = 0; y < height; y++) memcpy(&D.quad_im.buf[(size_t)y * width], im_orig + (size_t)y * stride, width);
  }
  t1 = now_s();
  D.times.decimate = t1 - t0;
  blur_or_sharpen(D.quad_im, D.prm.quad_sigma);
  double t2 = now_s();
  D.times.blur = t2 - t1;
  // Step 2. apriltag_quad_thresh
  threshold(D);
  double t3 = now_s();
  D.times.threshold = t3 - t2;
  UnionFind uf;
  connected_components(D, uf);
  double t4 = now_s();
  D.times.unionfind = t4 - t3;
  std::unordered_map<uint64_t, vector<Pt>> cmap;
  gradient_clusters(D, cmap);
  int w = D.quad_im.w, h = D.quad_im.h;
  // keep what do_quad_task would look at (size gates), canonical order C2
  D.clusters.clear();
  for (auto &kv : cmap) {
    int n = (int)kv.second.size();
    if (n < D.prm.min_cluster_pixels) continue;
    if (n > 2 * (2 * w + 2 * h)) continue;
    if (n < 24) continue;  // fit_quad's own first gate; dropped here so dumps match the GPU's kept set
    Cluster c;
    c.key = kv.first;
    c.pts.swap(kv.second);
    D.clusters.push_back(std::move(c));
  }
  std::sort(D.clusters.begin(), D.clusters.end(), [](const Cluster &a, const Cluster &b) { return a.key < b.key; });
  double t5 = now_s();
  D.times.clusters = t5 - t4;
  // fit_quads
  bool normal_border = false, reversed_border = false;
  int min_tag_width = 1000000;
  for (int f : D.fams) {
    if (family_at(f).width_at_border < min_tag_width) min_tag_width = family_at(f).width_at_border;
    normal_border |= !family_at(f).reversed_border;
    reversed_border |= family_at(f).reversed_border;
  }
  if (D.prm.quad_decimate > 1) min_tag_width = (int)(min_tag_width / D.prm.quad_decimate);
  if (min_tag_width < 3) min_tag_width = 3;
  D.quads_fit.clear();
  for (auto &c : D.clusters) {
    Quad q;
    memset(&q, 0, sizeof(q));
    if (fit_quad(D, D.quad_im, c.pts, &q, min_tag_width, normal_border, reversed_border)) {
      q.key = c.key;
      D.quads_fit.push_back(q);
    }
  }
  double t6 = now_s();
  D.times.fit_quads = t6 - t5;
  // rescale to full resolution
  D.quads_refined = D.quads_fit;
  if (D.prm.quad_decimate > 1) {
    for (auto &q : D.quads_refined)
      for (int j = 0; j < 4; j++) {
        if (D.prm.quad_decimate == 1.5f) {
          q.p[j][0] *= D.prm.quad_decimate;
          q.p[j][1] *= D.prm.quad_decimate;
        } else {
          q.p[j][0] = (float)((q.p[j][0] - 0.5) * D.prm.quad_decimate + 0.5);
          q.p[j][1] = (float)((q.p[j][1] - 0.5) * D.prm.quad_decimate + 0.5);
        }
      }
  }
  // Step 3. decode
  vector<ato_detection_t> dets;
  static const double kCos[4] = {1.0, 6.123233995736766e-17, -1.0, -1.8369701987210297e-16};
  static const double kSin[4] = {0.0, 1.0, 1.2246467991473532e-16, -1.0};
  for (auto &q : D.quads_refined) {
    if (D.prm.refine_edges) refine_edges(D, im_orig, width, height, stride, &q);
    if (quad_update_homographies(&q) != 0) continue;
    for (int f : D.fams) {
      const Family &family = family_at(f);
      if (family.reversed_border != q.reversed_border) continue;
      DecodeEntry entry;
      float decision_margin = quad_decode(D, family, im_orig, width, height, stride, &q, &entry);
      if (decision_margin >= 0 && entry.hamming < 255) {
        ato_detection_t det;
        memset(&det, 0, sizeof(det));
        det.family = f;
        det.id = entry.id;
        det.hamming = entry.hamming;
        det.decision_margin = decision_margin;
        // theta = rotation*pi/2 ; c = cos(theta), s = sin(theta) (glibc values tabulated)
        double c = kCos[entry.rotation], s = kSin[entry.rotation];
        double Rm[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
        mm33(q.H, Rm, det.H);
        homography_project(det.H, 0, 0, &det.c[0], &det.c[1]);
        for (int i = 0; i < 4; i++) {
          int tcx = (i == 1 || i == 2) ? 1 : -1;
          int tcy = (i < 2) ? 1 : -1;
          homography_project(det.H, tcx, tcy, &det.p[i][0], &det.p[i][1]);
        }
        dets.push_back(det);
      }
    }
  }
  double t7 = now_s();
  D.times.decode = t7 - t6;
  // Step 4. reconcile overlapping duplicates (same control flow as upstream incl. swap-with-last removal)
  {
    for (int i0 = 0; i0 < (int)dets.size(); i0++) {
      bool restart0 = false;
      for (int i1 = i0 + 1; i1 < (int)dets.size(); i1++) {
        ato_detection_t &det0 = dets[i0], &det1 = dets[i1];
        if (det0.id != det1.id || det0.family != det1.family) continue;
        if (polygon_overlaps_polygon(det0.p, det1.p)) {
          int pref = 0;
          pref = prefer_smaller(pref, det0.hamming, det1.hamming);
          pref = prefer_smaller(pref, -det0.decision_margin, -det1.decision_margin);
          for (int i = 0; i < 4; i++) {
            pref = prefer_smaller(pref, det0.p[i][0], det1.p[i][0]);
            pref = prefer_smaller(pref, det0.p[i][1], det1.p[i][1]);
          }
          if (pref < 0) {
            dets[i1] = dets.back();
            dets.pop_back();
            i1--;
          } else {
            dets[i0] = dets.back();
            dets.pop_back();
            i0--;
            restart0 = true;
            break;
          }
        }
      }
      (void)restart0;
    }
  }
  // canonical order C4
  std::sort(dets.begin(), dets.end(), [](const ato_detection_t &a, const ato_detection_t &b) {
    if (a.id != b.id) return a.id < b.id;
    if (a.family != b.family) return a.family < b.family;
    if (a.c[1] != b.c[1]) return a.c[1] < b.c[1];
    return a.c[0] < b.c[0];
  });
  double t8 = now_s();
  D.times.reconcile = t8 - t7;
  D.times.total = t8 - t0;
  int n = std::min((int)dets.size(), max_out);
  for (int i = 0; i < n; i++) out[i] = dets[i];
  return n;
}

}  // namespace

extern "C" {

void ato_default_params(ato_params_t *p) {
  p->quad_decimate = 2.0f;
  p->quad_sigma = 0.0f;
  p->refine_edges = 1;
  p->decode_sharpening = 0.25;
  p->min_cluster_pixels = 5;
  p->max_nmaxima = 10;
  p->critical_rad = (float)(10 * M_PI / 180);
  p->max_line_fit_mse = 10.0f;
  p->min_white_black_diff = 5;
  p->tile_size = 4;
  p->max_hamming = 2;
  p->family_mask = 1u << ATO_FAM_36H11;
}

void *ato_create(const ato_params_t *p) {
  Detector *D = new Detector();
  D->prm = *p;
  for (int f = 0; f < ATO_MAX_FAMILIES; f++)
    if ((p->family_mask & (1u << f)) && (f < ATO_NUM_FAMILIES || g_custom[f - ATO_NUM_FAMILIES].used)) D->fams.push_back(f);
  return D;
}
void ato_destroy(void *h) { delete (Detector *)h; }

int ato_detect(void *h, const uint8_t *gray, int width, int height, int stride, ato_detection_t *out, int max_out) {
  if (!h || !gray) return -1;
  return detect(*(Detector *)h, gray, width, height, stride, out, max_out);
}
void ato_get_times(void *h, ato_times_t *t) { *t = ((Detector *)h)->times; }
void ato_get_quad_dims(void *h, int *w, int *hh) {
  *w = ((Detector *)h)->quad_im.w;
  *hh = ((Detector *)h)->quad_im.h;
}
void ato_get_quad_image(void *h, uint8_t *out) {
  Detector *D = (Detector *)h;
  memcpy(out, D->quad_im.buf.data(), D->quad_im.buf.size());
}
void ato_get_threshold(void *h, uint8_t *out) {
  Detector *D = (Detector *)h;
  memcpy(out, D->thresh.buf.data(), D->thresh.buf.size());
}
void ato_get_tile_minmax(void *h, uint8_t *mn, uint8_t *mx) {
  Detector *D = (Detector *)h;
  memcpy(mn, D->tmin.data(), D->tmin.size());
  memcpy(mx, D->tmax.data(), D->tmax.size());
}
void ato_get_labels(void *h, uint32_t *label, uint32_t *size) {
  Detector *D = (Detector *)h;
  if (label) memcpy(label, D->label.data(), D->label.size() * 4);
  if (size) memcpy(size, D->lsize.data(), D->lsize.size() * 4);
}
int ato_num_clusters(void *h) { return (int)((Detector *)h)->clusters.size(); }
int ato_get_cluster(void *h, int i, uint64_t *key, uint32_t *packed_pts, int max_pts) {
  Detector *D = (Detector *)h;
  const Cluster &c = D->clusters[i];
  *key = c.key;
  int n = std::min((int)c.pts.size(), max_pts);
  for (int k = 0; k < n; k++) packed_pts[k] = (uint32_t)c.pts[k].x | ((uint32_t)c.pts[k].y << 16);
  return (int)c.pts.size();
}
int ato_num_points_total(void *h) { return ((Detector *)h)->total_points; }
int ato_get_quads(void *h, ato_quad_t *out, int max_out, int which) {
  Detector *D = (Detector *)h;
  const vector<Quad> &v = which ? D->quads_refined : D->quads_fit;
  int n = std::min((int)v.size(), max_out);
  for (int i = 0; i < n; i++) {
    memcpy(out[i].p, v[i].p, sizeof(out[i].p));
    out[i].reversed_border = v[i].reversed_border;
    out[i].key = v[i].key;
  }
  return (int)v.size();
}

void ato_estimate_pose(const ato_detection_t *det, double fx, double fy, double cx, double cy, double tagsize, ato_pose_t *best,
                       ato_pose_t *p1, ato_pose_t *p2) {
  estimate_pose(det, fx, fy, cx, cy, tagsize, best, p1, p2);
}

void ato_to_gray(const uint8_t *src, int enc, int width, int height, int stride, uint8_t *dst) {
  for (int y = 0; y < height; y++) {
    const uint8_t *row = src + (size_t)y * stride;
    uint8_t *o = dst + (size_t)y * width;
    if (enc == 0) {
      memcpy(o, row, width);
      continue;
    }
    int bpp = (enc == 1 || enc == 2) ? 3 : 4;
    bool bgr = (enc == 2 || enc == 4);
    for (int x = 0; x < width; x++) {
      int c0 = row[x * bpp], c1 = row[x * bpp + 1], c2 = row[x * bpp + 2];
      int r = bgr ? c2 : c0, g = c1, b = bgr ? c0 : c2;
      o[x] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14);
    }
  }
}

int ato_detect_batch(const ato_params_t *p, const uint8_t *frames, int n, int width, int height, int nthreads,
                     ato_detection_t *out, int *counts, int max_out, ato_times_t *sum_times) {
  return ato_detect_batch_enc(p, frames, 0, n, width, height, nthreads, out, counts, max_out, sum_times);
}

int ato_detect_batch_enc(const ato_params_t *p, const uint8_t *frames, int enc, int n, int width, int height, int nthreads,
                         ato_detection_t *out, int *counts, int max_out, ato_times_t *sum_times) {
  if (nthreads < 1) nthreads = 1;
  const int bpp = enc == 0 ? 1 : ((enc == 1 || enc == 2) ? 3 : 4);
  vector<std::thread> th;
  vector<ato_times_t> tsum(nthreads);
  for (int t = 0; t < nthreads; t++) memset(&tsum[t], 0, sizeof(ato_times_t));
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([=, &tsum]() {
      Detector *D = (Detector *)ato_create(p);
      vector<uint8_t> graybuf;
      if (enc != 0) graybuf.resize((size_t)width * height);
      for (int i = t; i < n; i += nthreads) {
        const uint8_t *src = frames + (size_t)i * width * height * bpp;
        if (enc != 0) {  // colour input: the conversion is part of the CPU path's per-frame work
          ato_to_gray(src, enc, width, height, width * bpp, graybuf.data());
          src = graybuf.data();
        }
        counts[i] = detect(*D, src, width, height, width, out + (size_t)i * max_out, max_out);
        const ato_times_t &x = D->times;
        ato_times_t &s = tsum[t];
        s.decimate += x.decimate;
        s.blur += x.blur;
        s.threshold += x.threshold;
        s.unionfind += x.unionfind;
        s.clusters += x.clusters;
        s.fit_quads += x.fit_quads;
        s.decode += x.decode;
        s.reconcile += x.reconcile;
        s.total += x.total;
      }
      ato_destroy(D);
    });
  }
  for (auto &t : th) t.join();
  if (sum_times) {
    memset(sum_times, 0, sizeof(*sum_times));
    for (int t = 0; t < nthreads; t++) {
      sum_times->decimate += tsum[t].decimate;
      sum_times->blur += tsum[t].blur;
      sum_times->threshold += tsum[t].threshold;
      sum_times->unionfind += tsum[t].unionfind;
      sum_times->clusters += tsum[t].clusters;
      sum_times->fit_quads += tsum[t].fit_quads;
      sum_times->decode += tsum[t].decode;
      sum_times->reconcile += tsum[t].reconcile;
      sum_times->total += tsum[t].total;
    }
  }
  return 0;
}

uint64_t ato_rotate90(uint64_t w, int nbits) { return rotate90(w, nbits); }
int ato_family_info(int fam, int *nbits, int *ncodes, int *width_at_border, int *total_width) {
  if (fam < 0 || fam >= ATO_MAX_FAMILIES || (fam >= ATO_NUM_FAMILIES && !g_custom[fam - ATO_NUM_FAMILIES].used)) return -1;
  *nbits = family_at(fam).nbits;
  *ncodes = family_at(fam).ncodes;
  *width_at_border = family_at(fam).width_at_border;
  *total_width = family_at(fam).total_width;
  return 0;
}
uint64_t ato_family_code(int fam, int idx) { return family_at(fam).codes[idx]; }

int ato_register_family(int slot, const char *name, int nbits, int ncodes, int width_at_border, int total_width, int reversed_border,
                        const signed char *bit_x, const signed char *bit_y, const uint64_t *codes) {
  if (slot < ATO_NUM_FAMILIES || slot >= ATO_MAX_FAMILIES || nbits < 1 || nbits > 64 || ncodes < 1 || !bit_x || !bit_y || !codes) return -1;
  CustomFamily &c = g_custom[slot - ATO_NUM_FAMILIES];
  c.name = name ? name : "custom";
  c.bx.assign(bit_x, bit_x + nbits);
  c.by.assign(bit_y, bit_y + nbits);
  c.codes.assign(codes, codes + ncodes);
  c.f = Family{c.name.c_str(), nbits, 0, ncodes, width_at_border, total_width, reversed_border != 0, c.bx.data(), c.by.data(), c.codes.data()};
  c.used = true;
  return 0;
}

}  // extern "C"
