// k_decode.cu -- per-quad stages (a11-a13): rescale to full resolution, refine_edges on the full-resolution
// image, homography (8x9 Gaussian elimination, double), border gray models, bilinear bit sampling, Laplacian
// sharpen, codeword lookup (<= max_hamming bit errors, 4 rotations).  One WARP per quad.
// Restates AprilRobotics apriltag.c refine_edges / quad_update_homographies / quad_decode / sharpen /
// quick_decode_codeword / rotate90 and common/homography.c homography_compute2 with the
// oracle's types and evaluation order; every order-dependent sum (line moments, gray models, scores) is
// accumulated in the serial order, redundantly by all lanes (no divergence), while the pixel gathers run
// lane-parallel.  Colour frames are converted to gray on the fly at each gather.
#include <math_constants.h>

#include <cstdlib>

#include "detector.h"

namespace b200at {

constexpr int DT = 128;  // threads per CTA (4 warps = 4 quads in flight per CTA)
constexpr int MAXTW = 12;

__device__ __forceinline__ int gray_at(const FrameDesc &fd, int enc, int bpp, int x, int y) {
  const uint8_t *p = fd.ptr + (size_t)y * fd.pitch + (size_t)x * bpp;
  if (enc == B200AT_ENC_MONO8) return p[0];
  uint32_t c0 = p[0], c1 = p[1], c2 = p[2];
  const bool bgr = (enc == B200AT_ENC_BGR8 || enc == B200AT_ENC_BGRA8);
  uint32_t r = bgr ? c2 : c0, b = bgr ? c0 : c2;
  return (int)((r * 4899u + c1 * 9617u + b * 1868u + 8192u) >> 14);
}

// branch-free variant for batched gathers: always three byte loads (for mono8 the three offsets coincide and the luma
// formula reduces to the identity: (v*16384 + 8192) >> 14 == v)
__device__ __forceinline__ int gray_at_bf(const uint8_t *base, size_t pitch, int bpp, int o1, int o2, bool bgr, int x, int y) {
  const uint8_t *p = base + (size_t)y * pitch + (size_t)x * bpp;
  const uint32_t c0 = p[0], c1 = p[o1], c2 = p[o2];
  const uint32_t r = bgr ? c2 : c0, b = bgr ? c0 : c2;
  return (int)((r * 4899u + c1 * 9617u + b * 1868u + 8192u) >> 14);
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ void hproject(const double *H, double x, double y, double *ox, double *oy) {
  double xx = H[0] * x + H[1] * y + H[2];
  double yy = H[3] * x + H[4] * y + H[5];
  double zz = H[6] * x + H[7] * y + H[8];
  *ox = xx / zz;
  *oy = yy / zz;
}

// homography_compute2 (8x9 Gaussian elimination with partial pivoting + back substitution), distributed over the
// warp: lane (l & 7) holds row (l & 7) of the augmented matrix in registers (the four 8-lane groups compute the same
// thing), pivot search / row swap / pivot-row broadcast are width-8 shuffles.  Operation order per element is the
// serial algorithm's, so the result is bit-identical to it.
__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

__device__ bool homography_compute2_warp(const float p[4][2], double H[9]) {
  const int row = threadIdx.x & 7;
  const int ci = row >> 1;
  const double c0 = (ci == 0 || ci == 3) ? -1 : 1, c1 = (ci == 0 || ci == 1) ? -1 : 1;
  const double c2 = p[ci][0], c3 = p[ci][1];
  double r[9];
  if ((row & 1) == 0) {
    r[0] = c0;
    r[1] = c1;
    r[2] = 1;
    r[3] = 0;
    r[4] = 0;
    r[5] = 0;
    r[6] = -c0 * c2;
    r[7] = -c1 * c2;
    r[8] = c2;
  } else {
    r[0] = 0;
    r[1] = 0;
    r[2] = 0;
    r[3] = c0;
    r[4] = c1;
    r[5] = 1;
    r[6] = -c0 * c3;
    r[7] = -c1 * c3;
    r[8] = c3;
  }
  const double epsilon = 1e-10;
  bool ok = true;
#pragma unroll
  for (int col = 0; col < 8; col++) {
    // first row >= col with the largest |A[row][col]| (strictly-greater scan == lowest index among equal maxima)
    double val = row >= col ? fabs(r[col]) : -1.0;
    int idx = row;
#pragma unroll
    for (int of = 4; of > 0; of >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, val, of, 8);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, of, 8);
      if (ov > val || (ov == val && oi < idx)) {
        val = ov;
        idx = oi;
      }
    }
    if (val < epsilon) ok = false;  // "matrix is singular" (uniform across the warp)
    // swap rows col <-> idx
    const int partner = row == col ? idx : (row == idx ? col : row);
#pragma unroll
    for (int j = 0; j < 9; j++)
      if (j >= col) r[j] = shfl8(r[j], partner);
    // broadcast the pivot row, eliminate below
    double pr[9];
#pragma unroll
    for (int j = 0; j < 9; j++) pr[j] = j >= col ? shfl8(r[j], col) : 0.0;
    if (row > col) {
      const double f = r[col] / pr[col];
      r[col] = 0;
#pragma unroll
      for (int j = 0; j < 9; j++)
        if (j > col) r[j] -= f * pr[j];
    }
  }
  // back substitution
  double xs[8];
#pragma unroll
  for (int col = 7; col >= 0; col--) {
    double sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (i > col) sum += r[i] * xs[i];
    const double mine = (r[8] - sum) / r[col];
    xs[col] = shfl8(mine, col);
  }
#pragma unroll
  for (int k = 0; k < 8; k++) H[k] = xs[k];
  H[8] = 1;
  return ok;
}

struct GrayModelD {
  double A00, A01, A02, A11, A12, A22, B0, B1, B2, C0, C1, C2;
};
__device__ __forceinline__ void gm_solve(GrayModelD &m) {
  // mat33_sym_solve: Cholesky, lower-triangular inverse, two triangular products (same op order as the oracle)
  double L0 = sqrt(m.A00);
  double L3 = m.A01 / L0;
  double L6 = m.A02 / L0;
  double L4 = sqrt(m.A11 - L3 * L3);
  double L7 = (m.A12 - L3 * L6) / L4;
  double L8 = sqrt(m.A22 - L6 * L6 - L7 * L7);
  double M0 = 1 / L0;
  double M3 = -L3 * M0 / L4;
  double M4 = 1 / L4;
  double M6 = (-L6 * M0 - L7 * M3) / L8;
  double M7 = -L7 * M4 / L8;
  double M8 = 1 / L8;
  double t0 = M0 * m.B0;
  double t1 = M3 * m.B0 + M4 * m.B1;
  double t2 = M6 * m.B0 + M7 * m.B1 + M8 * m.B2;
  m.C0 = M0 * t0 + M3 * t1 + M6 * t2;
  m.C1 = M4 * t1 + M7 * t2;
  m.C2 = M8 * t2;
}
__device__ __forceinline__ double gm_interp(const GrayModelD &m, double x, double y) { return m.C0 * x + m.C1 * y + m.C2; }

__device__ __forceinline__ unsigned long long rotate90_dev(unsigned long long w, int numBits) {
  int p = numBits;
  unsigned long long l = 0;
  if (numBits % 4 == 1) {
    p = numBits - 1;
    l = 1;
  }
  w = ((w >> l) << (p / 4 + l)) | (w >> (3 * p / 4 + l) << l) | (w & l);
  w &= ((1ULL << numBits) - 1);
  return w;
}

struct DecodeFams {
  DevFamily f[kMaxFamilies];
};

// pixel access of the decoder, in one place so that the row-marking pass of the sparse host path (k_refine<true>) visits
// exactly the pixels k_decode_bits / k_decode will read
struct BorderSample {
  double tagx, tagy;
  int ix, iy, is_white, valid;
};
__device__ __forceinline__ BorderSample border_sample(const double *H, int j, int wab, int width, int height) {
  BorderSample b;
  const float wabf = (float)wab;
  const int pat = j / wab, i = j % wab;
  float p0, p1, p2, p3;
  switch (pat) {
    case 0: p0 = -0.5f; p1 = 0.5f; p2 = 0; p3 = 1; b.is_white = 1; break;
    case 1: p0 = 0.5f; p1 = 0.5f; p2 = 0; p3 = 1; b.is_white = 0; break;
    case 2: p0 = wabf + 0.5f; p1 = .5f; p2 = 0; p3 = 1; b.is_white = 1; break;
    case 3: p0 = wabf - 0.5f; p1 = .5f; p2 = 0; p3 = 1; b.is_white = 0; break;
    case 4: p0 = 0.5f; p1 = -0.5f; p2 = 1; p3 = 0; b.is_white = 1; break;
    case 5: p0 = 0.5f; p1 = 0.5f; p2 = 1; p3 = 0; b.is_white = 0; break;
    case 6: p0 = 0.5f; p1 = wabf + 0.5f; p2 = 1; p3 = 0; b.is_white = 1; break;
    default: p0 = 0.5f; p1 = wabf - 0.5f; p2 = 1; p3 = 0; b.is_white = 0; break;
  }
  // float arithmetic, as the C expression (float + int*float) / int evaluates
  double tagx01 = (double)((p0 + (float)i * p2) / wabf);
  double tagy01 = (double)((p1 + (float)i * p3) / wabf);
  b.tagx = 2 * (tagx01 - 0.5);
  b.tagy = 2 * (tagy01 - 0.5);
  double px, py;
  hproject(H, b.tagx, b.tagy, &px, &py);
  b.ix = (int)px;
  b.iy = (int)py;
  b.valid = !(b.ix < 0 || b.iy < 0 || b.ix >= width || b.iy >= height);
  return b;
}
struct BitSample {
  double tagx, tagy, x, y;
  int x1, x2, y1, y2, valid;
};
__device__ __forceinline__ BitSample bit_sample(const double *H, int bitx, int bity, int wab, int width, int height) {
  BitSample b;
  double tagx01 = (bitx + 0.5) / (wab);
  double tagy01 = (bity + 0.5) / (wab);
  b.tagx = 2 * (tagx01 - 0.5);
  b.tagy = 2 * (tagy01 - 0.5);
  double px, py;
  hproject(H, b.tagx, b.tagy, &px, &py);
  // value_for_pixel: bilinear, pixel centres at +0.5
  b.x1 = (int)floor(px - 0.5);
  b.x2 = (int)ceil(px - 0.5);
  b.x = px - 0.5 - b.x1;
  b.y1 = (int)floor(py - 0.5);
  b.y2 = (int)ceil(py - 0.5);
  b.y = py - 0.5 - b.y1;
  b.valid = !(b.x1 < 0 || b.x2 >= width || b.y1 < 0 || b.y2 >= height);
  return b;
}

// ---- a11: rescale a fitted quad to the full-resolution frame ----
__device__ __forceinline__ void rescale_quad(const FitParams &fp, const QuadRec &q0, float p[4][2]) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (fp.quad_decimate == 1.5f) {
      p[j][0] = q0.p[j][0] * fp.quad_decimate;  // float *= float, as upstream for the 1.5 special case
      p[j][1] = q0.p[j][1] * fp.quad_decimate;
    } else if (fp.quad_decimate > 1) {
      p[j][0] = (float)(((double)q0.p[j][0] - 0.5) * (double)fp.quad_decimate + 0.5);
      p[j][1] = (float)(((double)q0.p[j][1] - 0.5) * (double)fp.quad_decimate + 0.5);
    } else {
      p[j][0] = q0.p[j][0];
      p[j][1] = q0.p[j][1];
    }
  }
}

// ---- a12: refine_edges ----
// one edge (a warp; every lane ends with the same line): centroid + unit normal of the line through the refined sample points
__device__ __forceinline__ void refine_edge_line(const FitParams &fp, const FrameDesc &fd, int width, int height, int bpp, int o1, int o2,
                                                 bool is_bgr, bool reversed, int lane, const float p[4][2], int edge, double line[4]) {
  const double range = (double)(fp.quad_decimate + 1.0f);
  const int nsteps = (int)floor(2.0 * range / 0.25) + 1;
  const int a = edge, b = (edge + 1) & 3;
  double nx = (double)(p[b][1] - p[a][1]);
  double ny = (double)(-p[b][0] + p[a][0]);
  double mag = sqrt(nx * nx + ny * ny);
  nx /= mag;
  ny /= mag;
  if (reversed) {
    nx = -nx;
    ny = -ny;
  }
  const int nsamples = max(16, (int)(mag / 8));
  double Mx = 0, My = 0, Mxx = 0, Mxy = 0, Myy = 0, N = 0;
  for (int s0 = 0; s0 < nsamples; s0 += 32) {
    const int s = s0 + lane;
    double bestx = 0, besty = 0;
    int has = 0;
    if (s < nsamples) {
      double alpha = (1.0 + s) / (nsamples + 1);
      double x0 = alpha * p[a][0] + (1 - alpha) * p[b][0];
      double y0 = alpha * p[a][1] + (1 - alpha) * p[b][1];
      double Mn = 0, Mcount = 0;
      // the pixel gathers of 8 steps are issued together (independent loads), then consumed in step order
      for (int k0 = 0; k0 < nsteps; k0 += 8) {
        int g1[8], g2[8];
        bool okk[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int k = k0 + u;
          const double n = -range + 0.25 * k;
          const double grange = 1;
          const int x1 = (int)(x0 + (n + grange) * nx);
          const int y1 = (int)(y0 + (n + grange) * ny);
          const int x2 = (int)(x0 + (n - grange) * nx);
          const int y2 = (int)(y0 + (n - grange) * ny);
          okk[u] = k < nsteps && !(x1 < 0 || x1 >= width || y1 < 0 || y1 >= height) &&
                   !(x2 < 0 || x2 >= width || y2 < 0 || y2 >= height);
          // unconditional loads from clamped (always valid) coordinates keep the 16 gathers independent
          g1[u] = gray_at_bf(fd.ptr, fd.pitch, bpp, o1, o2, is_bgr, okk[u] ? x1 : 0, okk[u] ? y1 : 0);
          g2[u] = gray_at_bf(fd.ptr, fd.pitch, bpp, o1, o2, is_bgr, okk[u] ? x2 : 0, okk[u] ? y2 : 0);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (!okk[u] || g1[u] < g2[u]) continue;
          const double n = -range + 0.25 * (k0 + u);
          const double weight = (double)((g2[u] - g1[u]) * (g2[u] - g1[u]));
          Mn += weight * n;  // integer-valued multiples of 0.25: exact, order independent
          Mcount += weight;
        }
      }
      if (Mcount != 0) {
        double n0 = Mn / Mcount;
        bestx = x0 + n0 * nx;
        besty = y0 + n0 * ny;
        has = 1;
      }
    }
    // moments of the 32 samples: butterfly sums (every lane ends with the same value).  The oracle adds the samples one
    // after the other; the association differs in the last bits of sums whose line parameters are rounded to float below.
    double m5[5] = {has ? bestx : 0.0, has ? besty : 0.0, has ? bestx * bestx : 0.0, has ? bestx * besty : 0.0, has ? besty * besty : 0.0};
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
#pragma unroll
      for (int q = 0; q < 5; q++) m5[q] += shfl_xor_d(m5[q], of);
    }
    Mx += m5[0];
    My += m5[1];
    Mxx += m5[2];
    Mxy += m5[3];
    Myy += m5[4];
    N += (double)__popc(__ballot_sync(0xffffffffu, has != 0));
  }
  double Ex = Mx / N, Ey = My / N;
  double Cxx = Mxx / N - Ex * Ex;
  double Cxy = Mxy / N - Ex * Ey;
  double Cyy = Myy / N - Ey * Ey;
  // atan2f / cosf / sinf of the C library: evaluated in double and rounded once to float
  float th = (float)atan2((double)(float)(-2 * Cxy), (double)(float)(Cyy - Cxx));
  double normal_theta = .5 * th;
  float nth = (float)normal_theta;
  line[0] = Ex;
  line[1] = Ey;
  line[2] = (double)(float)cos((double)nth);
  line[3] = (double)(float)sin((double)nth);
}

// the refined corners = intersections of adjacent lines (a corner whose lines are nearly parallel keeps its position)
__device__ __forceinline__ void refine_corners(const double lines[4][4], float p[4][2]) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    double A00 = lines[i][3], A01 = -lines[(i + 1) & 3][3];
    double A10 = -lines[i][2], A11 = lines[(i + 1) & 3][2];
    double B0 = -lines[i][0] + lines[(i + 1) & 3][0];
    double B1 = -lines[i][1] + lines[(i + 1) & 3][1];
    double det = A00 * A11 - A10 * A01;
    if (fabs(det) > 0.001) {
      double W00 = A11 / det, W01 = -A01 / det;
      double L0 = W00 * B0 + W01 * B1;
      p[i][0] = (float)(lines[i][0] + L0 * A00);
      p[i][1] = (float)(lines[i][1] + L0 * A10);
    }
  }
}

// one warp per quad: the four edges one after the other; every lane ends with the same refined corners
__device__ __forceinline__ void refine_edges_warp(const FitParams &fp, const FrameDesc &fd, int width, int height, int bpp, int o1,
                                                  int o2, bool is_bgr, bool reversed, int lane, float p[4][2]) {
  double lines[4][4];
#pragma unroll 1
  for (int edge = 0; edge < 4; edge++) refine_edge_line(fp, fd, width, height, bpp, o1, o2, is_bgr, reversed, lane, p, edge, lines[edge]);
  refine_corners(lines, p);
}

// ---- a13: homography of the (refined) corners; false = quad dropped (singular system / zero determinant) ----
__device__ __forceinline__ bool quad_homography_warp(const float p[4][2], double H[9]) {
  if (!homography_compute2_warp(p, H)) return false;
  double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
  return det != 0;
}

// ---- a13: decode one quad against every registered family (one warp) ----
__device__ __forceinline__ void decode_quad_warp(const Geo &g, const FitParams &fp, const DecodeFams &fams, const FrameDesc &fd,
                                                 const double *H, bool reversed, unsigned long long key, uint32_t frame, int lane,
                                                 double *val, double *nval, Cand *__restrict__ cands,
                                                 uint32_t *__restrict__ cand_count, uint32_t *__restrict__ counters) {
  const int width = g.W, height = g.H;
  const int enc = g.enc, bpp = g.bpp;
#pragma unroll 1
  for (int fi = 0; fi < fp.nfam; fi++) {
    const DevFamily &fam = fams.f[fi];
    if ((fam.reversed_border != 0) != reversed) continue;
    const int wab = fam.width_at_border, twd = fam.total_width, nbits = fam.nbits;
    GrayModelD wm = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, bm = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int nsamp = 8 * wab;
    for (int j0 = 0; j0 < nsamp; j0 += 32) {
      const int j = j0 + lane;
      double tagx = 0, tagy = 0;
      int v = 0, valid = 0, is_white = 0;
      if (j < nsamp) {
        const BorderSample b = border_sample(H, j, wab, width, height);
        tagx = b.tagx;
        tagy = b.tagy;
        is_white = b.is_white;
        if (b.valid) {
          v = gray_at(fd, enc, bpp, b.ix, b.iy);
          valid = 1;
        }
      }
      const int cnt = min(32, nsamp - j0);
      for (int k = 0; k < cnt; k++) {
        const int vl = __shfl_sync(0xffffffffu, valid, k);
        const int iw = __shfl_sync(0xffffffffu, is_white, k);
        const int vv = __shfl_sync(0xffffffffu, v, k);
        const double x = shfl_d(tagx, k), y = shfl_d(tagy, k);
        if (vl) {
          GrayModelD &m = iw ? wm : bm;
          const double gray = (double)vv;
          m.A00 += x * x;
          m.A01 += x * y;
          m.A02 += x;
          m.A11 += y * y;
          m.A12 += y;
          m.A22 += 1;
          m.B0 += x * gray;
          m.B1 += y * gray;
          m.B2 += gray;
        }
      }
    }
    if (wab > 1) {
      gm_solve(wm);
      gm_solve(bm);
    } else {
      gm_solve(wm);
      bm.C0 = 0;
      bm.C1 = 0;
      bm.C2 = bm.B2 / 4;
    }
    if ((gm_interp(wm, 0, 0) - gm_interp(bm, 0, 0) < 0) != (fam.reversed_border != 0)) continue;
    // bit samples
    for (int c = lane; c < twd * twd; c += 32) val[c] = 0;
    __syncwarp();
    const int min_coord = (wab - twd) / 2;
    for (int i = lane; i < nbits; i += 32) {
      const int bity = fam.bit_y[i], bitx = fam.bit_x[i];
      const BitSample b = bit_sample(H, bitx, bity, wab, width, height);
      if (!b.valid) continue;
      const double v = gray_at(fd, enc, bpp, b.x1, b.y1) * (1 - b.x) * (1 - b.y) + gray_at(fd, enc, bpp, b.x2, b.y1) * b.x * (1 - b.y) +
                       gray_at(fd, enc, bpp, b.x1, b.y2) * (1 - b.x) * b.y + gray_at(fd, enc, bpp, b.x2, b.y2) * b.x * b.y;
      if (v == -1) continue;
      double thresh = (gm_interp(bm, b.tagx, b.tagy) + gm_interp(wm, b.tagx, b.tagy)) / 2.0;
      val[twd * (bity - min_coord) + bitx - min_coord] = v - thresh;
    }
    __syncwarp();
    // sharpen: values + decode_sharpening * (3x3 Laplacian, zero padded), same term order as the serial loop
    for (int c = lane; c < twd * twd; c += 32) {
      const int y = c / twd, x = c % twd;
      double sh = 0;
#pragma unroll
      for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          if ((y + i - 1) < 0 || (y + i - 1) > twd - 1 || (x + j - 1) < 0 || (x + j - 1) > twd - 1) continue;
          const double kern = (i == 1 && j == 1) ? 4.0 : ((i == 1 || j == 1) ? -1.0 : 0.0);
          sh += val[(y + i - 1) * twd + (x + j - 1)] * kern;
        }
      nval[c] = val[c] + fp.decode_sharpening * sh;
    }
    __syncwarp();
    unsigned long long rcode = 0;
    float black_score = 0, white_score = 0;
    float black_score_count = 1, white_score_count = 1;
    for (int i = 0; i < nbits; i++) {
      const int bity = fam.bit_y[i], bitx = fam.bit_x[i];
      rcode = (rcode << 1);
      double v = nval[(bity - min_coord) * twd + bitx - min_coord];
      if (v > 0) {
        white_score = (float)((double)white_score + v);
        white_score_count++;
        rcode |= 1;
      } else {
        black_score = (float)((double)black_score - v);
        black_score_count++;
      }
    }
    __syncwarp();
    // quick_decode_codeword: rotation by rotation, lowest code index within max_hamming wins
    int e_id = 65535, e_ham = 255, e_rot = 0;
    {
      unsigned long long rc = rcode;
      for (int ridx = 0; ridx < 4; ridx++) {
        int best_k = 0x7fffffff, best_h = 255;
        for (int k = lane; k < fam.ncodes; k += 32) {
          int hd = __popcll(rc ^ fam.codes[k]);
          if (hd <= fp.max_hamming && k < best_k) {
            best_k = k;
            best_h = hd;
          }
        }
        for (int of = 16; of > 0; of >>= 1) {
          int ok = __shfl_xor_sync(0xffffffffu, best_k, of), oh = __shfl_xor_sync(0xffffffffu, best_h, of);
          if (ok < best_k) {
            best_k = ok;
            best_h = oh;
          }
        }
        if (best_k != 0x7fffffff) {
          e_id = best_k;
          e_ham = best_h;
          e_rot = ridx;
          break;
        }
        rc = rotate90_dev(rc, nbits);
      }
    }
    const float decision_margin = fminf(white_score / white_score_count, black_score / black_score_count);
    if (decision_margin >= 0 && e_ham < 255 && lane == 0) {
      const double kCos[4] = {1.0, 6.123233995736766e-17, -1.0, -1.8369701987210297e-16};
      const double kSin[4] = {0.0, 1.0, 1.2246467991473532e-16, -1.0};
      const double c = kCos[e_rot], s = kSin[e_rot];
      const double Rm[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
      Cand cd;
      cd.key = key;
      cd.family = fam.index;
      cd.id = e_id;
      cd.hamming = e_ham;
      cd.decision_margin = decision_margin;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) cd.H[i * 3 + j] = H[i * 3 + 0] * Rm[0 * 3 + j] + H[i * 3 + 1] * Rm[1 * 3 + j] + H[i * 3 + 2] * Rm[2 * 3 + j];
      hproject(cd.H, 0, 0, &cd.c[0], &cd.c[1]);
      for (int i = 0; i < 4; i++) {
        int tcx = (i == 1 || i == 2) ? 1 : -1;
        int tcy = (i < 2) ? 1 : -1;
        hproject(cd.H, tcx, tcy, &cd.p[i][0], &cd.p[i][1]);
      }
      uint32_t slot = atomicAdd(&cand_count[frame], 1u);
      if (slot < g.cand_cap)
        cands[(size_t)frame * g.cand_cap + slot] = cd;
      else
        atomicOr(&counters[CNT_STATUS], (uint32_t)ST_CANDS_FULL);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sparse host path (b200AprilTagsDetectBatchHost with pinned, device-mapped frames and integer quad_decimate f >= 2).
// Quad detection only ever reads every f-th source row (image_u8_decimate is a point subsample), so only those rows are
// staged by DMA.  The full-resolution pixels that refine_edges and the decoder read lie around the (few) fitted quads:
//   k_mark_quads   : rows x segments within (quad_decimate + 3) px of a quad's edges              -> need1
//   k_fetch_rows   : copies the marked segments of the missing rows from the caller's host frames (zero-copy reads over
//                    PCIe, 16 bytes per thread, coalesced) into the staged frame
//   k_refine<true> : refine_edges + homography; marks the rows of every pixel the decoder will sample -> need2
//   k_fetch_rows   : need2 & ~need1
//   k_decode_bits  : the decode of k_decode, from the stored homographies
// Every access of refine_edges is within range + 1 (+1 for the integer truncation) pixels of a point on a quad edge, hence
// inside the expanded bounding box; the decoder's accesses are enumerated exactly (border_sample / bit_sample).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mark_row(const Geo &g, unsigned long long *__restrict__ need, uint32_t frame, int y, int xa, int xb) {
  if (y < 0 || y >= g.H || (y % g.row_step) == 0) return;  // rows that are multiples of row_step were staged by DMA
  xa = max(xa, 0);
  xb = min(xb, g.W - 1);
  if (xa > xb) return;
  const int sa = xa >> g.seg_shift, sb = xb >> g.seg_shift;
  const unsigned long long m = (sb >= 63 ? ~0ull : ((2ull << sb) - 1ull)) & ~((1ull << sa) - 1ull);
  atomicOr(&need[(size_t)frame * g.H + y], m);
}

// One warp per quad, edge by edge, lanes over the rows of the edge's band.  refine_edges samples at (x0, y0) + (n +- 1) * unit
// normal with (x0, y0) on the edge and |n| <= quad_decimate + 1, truncated to int: every sample lies within
// m = quad_decimate + 3 pixels (per axis) of its edge point.  Row y therefore needs the x-range of the edge points with
// |y0 - y| <= m, widened by m (+1 against rounding of the range arithmetic).
__global__ void __launch_bounds__(256) k_mark_quads(Geo g, FitParams fp, const QuadRec *__restrict__ quads,
                                                    const uint32_t *__restrict__ counters, unsigned long long *__restrict__ need1) {
  const int lane = threadIdx.x & 31;
  const uint32_t nq = min(counters[CNT_QUADS], g.quad_cap);
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const float m = fp.quad_decimate + 3.0f;
  for (uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < nq; qi += nwarps) {
    const QuadRec q0 = quads[qi];
    float p[4][2];
    rescale_quad(fp, q0, p);
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
      const float xa = p[e][0], ya = p[e][1], xb = p[(e + 1) & 3][0], yb = p[(e + 1) & 3][1];
      const float dx = xb - xa, dy = yb - ya;
      const int r0 = max(0, (int)floorf(fminf(ya, yb) - m)), r1 = min(g.H - 1, (int)ceilf(fmaxf(ya, yb) + m));
      for (int y = r0 + lane; y <= r1; y += 32) {
        float lo = fminf(xa, xb), hi = fmaxf(xa, xb);
        if (fabsf(dy) > 1e-3f) {
          float t0 = ((float)y - m - 1.0f - ya) / dy, t1 = ((float)y + m + 1.0f - ya) / dy;
          if (t0 > t1) {
            const float t = t0;
            t0 = t1;
            t1 = t;
          }
          t0 = fmaxf(t0, 0.0f);
          t1 = fminf(t1, 1.0f);
          if (t0 > t1) continue;
          const float x0 = xa + t0 * dx, x1 = xa + t1 * dx;
          lo = fminf(x0, x1);
          hi = fmaxf(x0, x1);
        }
        mark_row(g, need1, q0.frame, y, (int)floorf(lo - m - 1.0f), (int)ceilf(hi + m + 1.0f));
      }
    }
  }
}

// One warp per (frame, row): copies the 16-byte chunks of row y whose segment is in need & ~have (most rows need nothing and
// cost one word load).  grid (ceil(H / 8), frames), 8 warps per CTA.
__global__ void __launch_bounds__(256) k_fetch_rows(Geo g, const FrameDesc *__restrict__ src, const FrameDesc *__restrict__ dst,
                                                    const unsigned long long *__restrict__ need,
                                                    const unsigned long long *__restrict__ have, uint32_t *__restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5), fr = blockIdx.y;
  if (y >= g.H || (y % g.row_step) == 0) return;
  unsigned long long w = need[(size_t)fr * g.H + y];
  if (have) w &= ~have[(size_t)fr * g.H + y];
  if (w == 0) return;
  const FrameDesc s = src[fr], d = dst[fr];
  // both frames are 16-byte aligned with 16-byte-multiple pitches (checked on the host), so a chunk never leaves its row
  const uint8_t *srow = s.ptr + (size_t)y * s.pitch;
  uint8_t *drow = const_cast<uint8_t *>(d.ptr) + (size_t)y * d.pitch;
  const int row_bytes = g.W * g.bpp;
  const int nchunks = (row_bytes + 15) >> 4;
  uint32_t copied = 0;
  for (int c0 = 0; c0 < nchunks; c0 += 128) {
    uint4 v[4];
    bool take[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {  // four independent PCIe reads in flight per lane
      const int c = c0 + u * 32 + lane;
      const int b0 = c * 16;
      take[u] = false;
      if (c < nchunks) {
        const int pa = b0 / g.bpp, pb = min(g.W - 1, (b0 + 15) / g.bpp);  // first / last pixel with a byte in the chunk
        take[u] = ((w >> (pa >> g.seg_shift)) | (w >> (pb >> g.seg_shift))) & 1ull;
      }
      if (take[u]) v[u] = *reinterpret_cast<const uint4 *>(srow + b0);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      if (take[u]) {
        *reinterpret_cast<uint4 *>(drow + (size_t)(c0 + u * 32 + lane) * 16) = v[u];
        copied++;
      }
    }
  }
  copied = __reduce_add_sync(0xffffffffu, copied);
  if (lane == 0 && copied) atomicAdd(&counters[CNT_FETCHED], copied);
}

// what follows refine_edges for one quad (one warp): refined corners out, homography, and on the sparse host path the rows of
// every pixel the decoder will sample
template <bool MARK>
__device__ __forceinline__ void refine_finish(const Geo &g, const FitParams &fp, const DecodeFams &fams, const QuadRec &q0, uint32_t qi,
                                              const float p[4][2], bool reversed, int lane, QuadRec *__restrict__ quads_refined,
                                              double *__restrict__ quad_H, unsigned long long *__restrict__ need2) {
  if (lane == 0) {
    QuadRec qr = q0;
    for (int j = 0; j < 4; j++) {
      qr.p[j][0] = p[j][0];
      qr.p[j][1] = p[j][1];
    }
    quads_refined[qi] = qr;
  }
  double H[9];
  const bool ok = quad_homography_warp(p, H);
  double *hq = quad_H + (size_t)qi * 10;
  if (lane < 9) hq[lane] = H[lane];
  if (lane == 9) hq[9] = ok ? 1.0 : 0.0;
  if (MARK && ok) {
    for (int fi = 0; fi < fp.nfam; fi++) {
      const DevFamily &fam = fams.f[fi];
      if ((fam.reversed_border != 0) != reversed) continue;
      const int wab = fam.width_at_border;
      for (int j = lane; j < 8 * wab; j += 32) {
        const BorderSample b = border_sample(H, j, wab, g.W, g.H);
        if (b.valid) mark_row(g, need2, q0.frame, b.iy, b.ix, b.ix);
      }
      for (int i = lane; i < fam.nbits; i += 32) {
        const BitSample b = bit_sample(H, fam.bit_x[i], fam.bit_y[i], wab, g.W, g.H);
        if (!b.valid) continue;
        mark_row(g, need2, q0.frame, b.y1, b.x1, b.x2);
        mark_row(g, need2, q0.frame, b.y2, b.x1, b.x2);
      }
    }
  }
}

template <bool MARK, int MINB>
__global__ void __launch_bounds__(DT, MINB) k_refine(Geo g, FitParams fp, DecodeFams fams, const FrameDesc *__restrict__ frames,
                                               const QuadRec *__restrict__ quads, QuadRec *__restrict__ quads_refined,
                                               double *__restrict__ quad_H, const uint32_t *__restrict__ counters,
                                               unsigned long long *__restrict__ need2) {
  const int lane = threadIdx.x & 31;
  const uint32_t nq = min(counters[CNT_QUADS], g.quad_cap);
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const int bpp = g.bpp;
  const int o1 = bpp > 1 ? 1 : 0, o2 = bpp > 1 ? 2 : 0;
  const bool is_bgr = (g.enc == B200AT_ENC_BGR8 || g.enc == B200AT_ENC_BGRA8);
  for (uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < nq; qi += nwarps) {
    const QuadRec q0 = quads[qi];
    const FrameDesc fd = frames[q0.frame];
    const bool reversed = q0.reversed_border != 0;
    float p[4][2];
    rescale_quad(fp, q0, p);
    if (fp.refine_edges) refine_edges_warp(fp, fd, g.W, g.H, bpp, o1, o2, is_bgr, reversed, lane, p);
    refine_finish<MARK>(g, fp, fams, q0, qi, p, reversed, lane, quads_refined, quad_H, need2);
  }
}

// Latency form of k_refine for small batches (a single frame has a few dozen quads: the warp-per-quad kernel leaves the GPU
// empty and walks the four edges one after the other): one CTA of four warps per quad, one edge per warp.
template <bool MARK>
__global__ void __launch_bounds__(128) k_refine_cta(Geo g, FitParams fp, DecodeFams fams, const FrameDesc *__restrict__ frames,
                                                    const QuadRec *__restrict__ quads, QuadRec *__restrict__ quads_refined,
                                                    double *__restrict__ quad_H, const uint32_t *__restrict__ counters,
                                                    unsigned long long *__restrict__ need2) {
  __shared__ double s_lines[4][4];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t nq = min(counters[CNT_QUADS], g.quad_cap);
  const int bpp = g.bpp;
  const int o1 = bpp > 1 ? 1 : 0, o2 = bpp > 1 ? 2 : 0;
  const bool is_bgr = (g.enc == B200AT_ENC_BGR8 || g.enc == B200AT_ENC_BGRA8);
  for (uint32_t qi = blockIdx.x; qi < nq; qi += gridDim.x) {
    const QuadRec q0 = quads[qi];
    const FrameDesc fd = frames[q0.frame];
    const bool reversed = q0.reversed_border != 0;
    float p[4][2];
    rescale_quad(fp, q0, p);
    if (fp.refine_edges) {
      double line[4];
      refine_edge_line(fp, fd, g.W, g.H, bpp, o1, o2, is_bgr, reversed, lane, p, wid, line);
      if (lane == 0) {
        s_lines[wid][0] = line[0];
        s_lines[wid][1] = line[1];
        s_lines[wid][2] = line[2];
        s_lines[wid][3] = line[3];
      }
      __syncthreads();
      double lines[4][4];
#pragma unroll
      for (int e = 0; e < 4; e++)
#pragma unroll
        for (int k = 0; k < 4; k++) lines[e][k] = s_lines[e][k];
      refine_corners(lines, p);
      __syncthreads();  // (s_lines is rewritten for the next quad)
    }
    if (wid == 0) refine_finish<MARK>(g, fp, fams, q0, qi, p, reversed, lane, quads_refined, quad_H, need2);
  }
}

template <int MINB>
__global__ void __launch_bounds__(DT, MINB) k_decode_bits(Geo g, FitParams fp, DecodeFams fams, const FrameDesc *__restrict__ frames,
                                                    const QuadRec *__restrict__ quads, const double *__restrict__ quad_H,
                                                    Cand *__restrict__ cands, uint32_t *__restrict__ cand_count,
                                                    uint32_t *__restrict__ counters) {
  __shared__ double s_val[DT / 32][MAXTW * MAXTW];
  __shared__ double s_new[DT / 32][MAXTW * MAXTW];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t nq = min(counters[CNT_QUADS], g.quad_cap);
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; qi < nq; qi += nwarps) {
    const double *hq = quad_H + (size_t)qi * 10;
    if (hq[9] == 0.0) continue;
    double H[9];
#pragma unroll
    for (int k = 0; k < 9; k++) H[k] = hq[k];
    const QuadRec q0 = quads[qi];
    const FrameDesc fd = frames[q0.frame];
    decode_quad_warp(g, fp, fams, fd, H, q0.reversed_border != 0, q0.key, q0.frame, lane, s_val[wid], s_new[wid], cands, cand_count,
                     counters);
  }
}

constexpr int kRefineCtaFrames = 4;  // batches of at most this many frames refine with one CTA per quad (k_refine_cta)
constexpr int kDecodeCtasPerSm = 6;  // persistent CTAs per SM = register budget of k_refine / k_decode_bits (measured: 4 / 5 / 6 CTAs -> decode 1.70 / 1.69 / 1.66 ms)

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// sparse host path, first half: mark the rows refine_edges will read and fetch them from the caller's frames.  May run on a
// side stream while the next sub-batch's quad detection occupies the compute stream (capi.cu).
int launch_sparse_fetch1(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  const dim3 gf((g.H + 7) / 8, nframes);
  // (measurement switch: B200AT_SPARSE_NOFETCH=1 skips the on-demand fetches -- WRONG results, shows what they cost)
  static const bool nofetch = getenv("B200AT_SPARSE_NOFETCH") != nullptr;
  k_mark_quads<<<sm_count() * 2, 256, 0, s>>>(g, ws.fp, ws.quads, ws.counters, ws.need1);
  if (!nofetch) k_fetch_rows<<<gf, 256, 0, s>>>(g, ws.src_frames, ws.frames, ws.need1, nullptr, ws.counters);
  return 2;
}

// sparse host path, second half: refine (marks the decoder's sample rows), fetch what is still missing, decode
int launch_sparse_back(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  static const bool nofetch = getenv("B200AT_SPARSE_NOFETCH") != nullptr;
  cudaMemsetAsync(ws.cand_count, 0, (size_t)nframes * sizeof(uint32_t), s);
  DecodeFams df;
  for (int i = 0; i < kMaxFamilies; i++) df.f[i] = ws.fams[i];
  const int ctas = sm_count() * kDecodeCtasPerSm;
  const dim3 gf((g.H + 7) / 8, nframes);
  if (nframes <= kRefineCtaFrames)
    k_refine_cta<true><<<sm_count() * 4, 128, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quads_refined, ws.quad_H, ws.counters, ws.need2);
  else
    k_refine<true, kDecodeCtasPerSm><<<ctas, DT, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quads_refined, ws.quad_H, ws.counters, ws.need2);
  if (!nofetch) k_fetch_rows<<<gf, 256, 0, s>>>(g, ws.src_frames, ws.frames, ws.need2, ws.need1, ws.counters);
  k_decode_bits<kDecodeCtasPerSm><<<ctas, DT, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quad_H, ws.cands, ws.cand_count, ws.counters);
  return 4;
}

int launch_decode(const Workspace &ws, int nframes, cudaStream_t s) {
  const Geo &g = ws.g;
  if (g.row_step != 0)  // sparse host path: the caller (capi.cu) zeroed need1 / need2 when it staged the frames
    return launch_sparse_fetch1(ws, nframes, s) + launch_sparse_back(ws, nframes, s);
  cudaMemsetAsync(ws.cand_count, 0, (size_t)nframes * sizeof(uint32_t), s);
  DecodeFams df;
  for (int i = 0; i < kMaxFamilies; i++) df.f[i] = ws.fams[i];
  // two kernels instead of one fused kernel that needs 167 registers (measured: 2.13 -> 1.69 ms)
  const int ctas = sm_count() * kDecodeCtasPerSm;
  if (nframes <= kRefineCtaFrames)
    k_refine_cta<false><<<sm_count() * 4, 128, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quads_refined, ws.quad_H, ws.counters, nullptr);
  else
    k_refine<false, kDecodeCtasPerSm><<<ctas, DT, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quads_refined, ws.quad_H, ws.counters, nullptr);
  k_decode_bits<kDecodeCtasPerSm><<<ctas, DT, 0, s>>>(g, ws.fp, df, ws.frames, ws.quads, ws.quad_H, ws.cands, ws.cand_count, ws.counters);
  return 3;
}

}  // namespace b200at
