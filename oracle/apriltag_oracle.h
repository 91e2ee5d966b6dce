/*
 * apriltag_oracle.h -- C API of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * This is a CPU restatement of the AprilRobotics `apriltag` 3.x detector chain
 * (`apriltag_detector_detect` + `estimate_tag_pose`) that BASELINE.json's north_star names as the
 * parity oracle.  Neither the AprilRobotics sources nor the closed cuAprilTags / VPI libraries the
 * reference node calls (/root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:450,491,229,291)
 * exist in this container, so the algorithm is restated from the published upstream behaviour
 * (upstream file/function names are cited at every function).
 *
 * PARITY STATUS: pinned at the *output* level by the reference's own golden vector
 * (isaac_ros_apriltag/test/isaac_ros_apriltag_pol_test.py:117-175, fixture re-synthesised) and by the
 * upstream code-table heads; the *intermediate* stages (threshold image, labels, clusters, quads) are
 * "parity unpinned" -- nothing in the reference pins them.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything in oracle/.  The product path (isaac_ros_apriltag_b200/) never links or loads it.
 */
#ifndef APRILTAG_ORACLE_H_
#define APRILTAG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ATO_FAM_36H11 = 0, ATO_FAM_25H9 = 1, ATO_FAM_16H5 = 2, ATO_FAM_36H10 = 3, ATO_NUM_FAMILIES = 4,
       ATO_FAM_CUSTOM0 = 4, ATO_FAM_CUSTOM1 = 5, ATO_MAX_FAMILIES = 6 /* slots 4, 5: ato_register_family */ };

/* apriltag_detector_create() defaults (upstream apriltag.c) */
typedef struct {
  float quad_decimate;      /* 2.0 */
  float quad_sigma;         /* 0.0 */
  int refine_edges;         /* 1 */
  double decode_sharpening; /* 0.25 */
  int min_cluster_pixels;   /* 5 */
  int max_nmaxima;          /* 10 */
  float critical_rad;       /* 10 deg in rad */
  float max_line_fit_mse;   /* 10 */
  int min_white_black_diff; /* 5 */
  int tile_size;            /* 4 (upstream hard-codes; the reference node exposes it, apriltag_node.cpp:566) */
  int max_hamming;          /* 2 (apriltag_detector_add_family default bits_corrected) */
  uint32_t family_mask;     /* bit i = family i registered, in index order */
} ato_params_t;

typedef struct {
  int family;
  int id;
  int hamming;
  float decision_margin;
  double H[9]; /* row-major 3x3, tag [-1,1]^2 -> pixels */
  double c[2];
  double p[4][2]; /* AprilRobotics order: (-1,1),(1,1),(1,-1),(-1,-1) in tag coords */
} ato_detection_t;

typedef struct {
  double R[9]; /* row-major */
  double t[3];
  double err;
} ato_pose_t;

typedef struct {
  float p[4][2];
  int reversed_border;
  uint64_t key; /* canonical cluster key (max_rep<<32 | min_rep), rep = min pixel index */
} ato_quad_t;

/* per-stage wall times of the last detect call, seconds (mirrors upstream timeprofile stamps) */
typedef struct {
  double decimate, blur, threshold, unionfind, clusters, fit_quads, decode, reconcile, total;
} ato_times_t;

void ato_default_params(ato_params_t *p);
void *ato_create(const ato_params_t *p);
void ato_destroy(void *h);

/* Full chain on a mono8 image.  Returns number of detections written (<= max_out), or -1. */
int ato_detect(void *h, const uint8_t *gray, int width, int height, int stride, ato_detection_t *out,
               int max_out);
void ato_get_times(void *h, ato_times_t *t);

/* Intermediates of the last ato_detect call (for stage-by-stage parity tests). */
void ato_get_quad_dims(void *h, int *w, int *hh);                  /* decimated image size */
void ato_get_quad_image(void *h, uint8_t *out);                    /* decimated (+blurred) gray, w*h */
void ato_get_threshold(void *h, uint8_t *out);                     /* {0,127,255}, w*h */
void ato_get_tile_minmax(void *h, uint8_t *mn, uint8_t *mx);       /* after 3x3 dilate/erode; tw*th */
void ato_get_labels(void *h, uint32_t *label, uint32_t *size);     /* canonical label = min idx; size[] per pixel = set size */
int ato_num_clusters(void *h);                                     /* clusters with >= 24 pts and <= perimeter bound */
int ato_get_cluster(void *h, int i, uint64_t *key, uint32_t *packed_pts, int max_pts); /* returns npts; pts sorted as fed to line fit */
int ato_num_points_total(void *h);                                 /* E: all emitted boundary points */
int ato_get_quads(void *h, ato_quad_t *out, int max_out, int which); /* which: 0 = fit (decimated coords), 1 = rescaled+refined */

/* apriltag_pose.c: estimate_tag_pose (homography init + orthogonal iteration + ambiguity). */
void ato_estimate_pose(const ato_detection_t *det, double fx, double fy, double cx, double cy, double tagsize,
                       ato_pose_t *best, ato_pose_t *p1, ato_pose_t *p2);

/* colour -> gray used on both sides of the boundary (OpenCV fixed-point BT.601: (R*4899+G*9617+B*1868+8192)>>14).
 * enc: 0 mono8, 1 rgb8, 2 bgr8, 3 rgba8, 4 bgra8 */
void ato_to_gray(const uint8_t *src, int enc, int width, int height, int stride, uint8_t *dst);

/* Frame-parallel batch (one single-thread detector per worker): the CPU baseline "all cores" mode.
 * frames: n contiguous mono8 frames (stride = width).  counts[n]; out[n*max_out]. Returns 0. */
int ato_detect_batch(const ato_params_t *p, const uint8_t *frames, int n, int width, int height, int nthreads,
                     ato_detection_t *out, int *counts, int max_out, ato_times_t *sum_times);

/* same, frames in any of the five encodings (the colour->gray conversion is done per frame inside the workers) */
int ato_detect_batch_enc(const ato_params_t *p, const uint8_t *frames, int enc, int n, int width, int height, int nthreads,
                         ato_detection_t *out, int *counts, int max_out, ato_times_t *sum_times);

/* upstream rotate90 / codebook access for known-answer tests */
uint64_t ato_rotate90(uint64_t w, int nbits);
int ato_family_info(int fam, int *nbits, int *ncodes, int *width_at_border, int *total_width);
uint64_t ato_family_code(int fam, int idx);
/* registers a family in slot ATO_FAM_CUSTOM0 / 1 (upstream apriltag_family_t fields; bit coordinates relative to the border's first
 * cell, signed); detectors created afterwards with that family_mask bit use it.  0 on success. */
int ato_register_family(int slot, const char *name, int nbits, int ncodes, int width_at_border, int total_width, int reversed_border,
                        const signed char *bit_x, const signed char *bit_y, const uint64_t *codes);

#ifdef __cplusplus
}
#endif
#endif
