/*
 * b200_apriltags.h -- C ABI of the B200-native AprilTag detector library (libb200apriltags.so).
 *
 * Part 1 mirrors, symbol for symbol, the three entry points the reference node binds from the closed
 * cuAprilTags library, so the library is a link-time drop-in for
 *   /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:450-452  (nvCreateAprilTagsDetector)
 *   /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:491-493  (cuAprilTagsDetect)
 *   /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:556      (cuAprilTagsDestroy)
 * with the structs those call sites fill and read (:401-406, :409-418, :447, :481-486, :509-516).
 *
 * Part 2 (b200AprilTags*) is the extended surface this build adds: batches of frames per launch, every
 * family/encoding the reference's VPI path accepts (apriltag_node.cpp:47-58, :76-82), host-buffer entry
 * points, AprilRobotics detector knobs, per-stage timing and intermediate-buffer read-back for parity tests.
 *
 * All functions return 0 on success, non-zero error codes otherwise; no C++ exception crosses this ABI.
 * Plain pointers and sizes only.
 *
 * Threading: a handle is not thread-safe (one call in flight per handle, like the reference's single callback group,
 * apriltag_node.cpp:605-610); independent handles -- e.g. one per GPU via b200AprilTagsOptions_t::device -- may be used
 * from different threads or from one thread.  The caller's current CUDA device is left unchanged by every call.
 */
#ifndef B200_APRILTAGS_H_
#define B200_APRILTAGS_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__) || defined(B200_APRILTAGS_USE_CUDA_HEADERS)
#include <cuda_runtime_api.h>
#include <vector_types.h>
#else
/* layout-compatible stand-ins so plain C / ctypes / cgo callers need no CUDA headers */
#ifndef __VECTOR_TYPES_H__
/* CUDA's float2 is __align__(8): the stand-in must be too, or cuAprilTagsID_t shrinks from 88 to 84 bytes */
#if defined(__cplusplus)
typedef struct alignas(8) { float x, y; } float2;
#elif defined(_MSC_VER)
typedef __declspec(align(8)) struct { float x, y; } float2;
#else
typedef struct { float x, y; } __attribute__((aligned(8))) float2;
#endif
typedef struct { unsigned char x, y, z; } uchar3;
#endif
#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st *cudaStream_t;
#endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 1: cuAprilTags-shaped surface (what apriltag_node.cpp's CUAprilTagImpl binds)
 * ---------------------------------------------------------------------------------------------- */

typedef struct cuAprilTagsHandle_st *cuAprilTagsHandle;

/* apriltag_node.cpp:409-418 reads orientation (column major) and translation, :509-516 id and corners. */
typedef struct {
  float2 corners[4];    /* message order: reverse of AprilRobotics p[0..3] */
  uint16_t id;
  uint8_t hamming_error;
  float orientation[9]; /* 3x3 rotation, COLUMN major, camera optical frame */
  float translation[3]; /* in units of tag_dim */
} cuAprilTagsID_t;         /* 88 bytes, 8-byte aligned (float2), orientation at offset 36 -- asserted in capi.cu and the node core */

/* apriltag_node.cpp:481-486 */
typedef struct {
  uchar3 *dev_ptr; /* device pointer, pitch-linear rgb8/bgr8 */
  size_t pitch;    /* bytes */
  uint16_t width;
  uint16_t height;
} cuAprilTagsImageInput_t;

/* apriltag_node.cpp:447 brace-initialises {fx, fy, cx, cy} */
typedef struct {
  float fx, fy, cx, cy;
} cuAprilTagsCameraIntrinsics_t;

/* apriltag_node.cpp:401-406 */
typedef enum {
  NVAT_TAG36H11 = 0,
  NVAT_ENUM_SIZE = 0x7fffffff
} cuAprilTagsFamily;

/* 0 on success.  Detector is sized once for (img_width, img_height) (apriltag_node.cpp:450-457). */
int nvCreateAprilTagsDetector(cuAprilTagsHandle *hApriltags, const uint32_t img_width, const uint32_t img_height,
                              const uint32_t tile_size, const cuAprilTagsFamily tag_family,
                              const cuAprilTagsCameraIntrinsics_t *cam, float tag_dim);

/* Synchronous: tags_out (HOST, caller-allocated, max_tags entries) is valid on return
 * (apriltag_node.cpp:490-503).  The bgr8/rgb8 distinction is not visible at this ABI (the reference passes
 * both through the same uchar3 pointer); the luma weights assume the default set by
 * b200AprilTagsSetInputEncoding (bgr8). */
uint32_t cuAprilTagsDetect(cuAprilTagsHandle hApriltags, const cuAprilTagsImageInput_t *img_input,
                           cuAprilTagsID_t *tags_out, uint32_t *num_tags, const uint32_t max_tags,
                           cudaStream_t input_stream);

int cuAprilTagsDestroy(cuAprilTagsHandle hApriltags);

/* ------------------------------------------------------------------------------------------------
 * Part 2: extended surface
 * ---------------------------------------------------------------------------------------------- */

enum {
  B200AT_OK = 0,
  B200AT_ERR_INVALID_ARG = 1,
  B200AT_ERR_UNSUPPORTED = 2, /* family / tile size / decimation not available on this backend */
  B200AT_ERR_CUDA = 3,
  B200AT_ERR_NOMEM = 4,
  B200AT_ERR_OVERFLOW = 5,    /* a bounded device buffer overflowed; results are truncated, see status word */
  B200AT_ERR_NO_DEVICE = 6
};

/* family indices (bit i of family_mask); same order as the oracle */
enum { B200AT_FAM_36H11 = 0, B200AT_FAM_25H9 = 1, B200AT_FAM_16H5 = 2, B200AT_FAM_36H10 = 3, B200AT_NUM_FAMILIES = 4,
       B200AT_FAM_CUSTOM0 = 4, B200AT_FAM_CUSTOM1 = 5, B200AT_MAX_FAMILIES = 6 /* slots 4, 5: b200AprilTagsRegisterFamily */ };

/* A tag family supplied by the caller: the fields of AprilRobotics' apriltag_family_t (tagStandard41h12.c, tagCircle21h7.c, ...).
 * This is how the families the reference's VPI path lists (apriltag_node.cpp:47-58) but whose code tables are not built in are
 * used: copy the table from the upstream tag*.c file.  bit_x / bit_y are relative to the border's first cell and may be negative
 * or >= width_at_border (data bits outside the border); reversed_border = the border is white on a black surround. */
typedef struct {
  uint32_t struct_size;      /* sizeof(b200AprilTagsFamilyDesc_t) */
  uint32_t nbits;            /* <= 52 */
  uint32_t ncodes;
  uint32_t width_at_border;
  uint32_t total_width;      /* <= 12 */
  uint32_t reversed_border;
  const int8_t *bit_x;       /* [nbits] */
  const int8_t *bit_y;
  const uint64_t *codes;     /* [ncodes] */
} b200AprilTagsFamilyDesc_t;

/* sensor_msgs encodings the reference's VPI path accepts (apriltag_node.cpp:76-82) */
enum { B200AT_ENC_MONO8 = 0, B200AT_ENC_RGB8 = 1, B200AT_ENC_BGR8 = 2, B200AT_ENC_RGBA8 = 3, B200AT_ENC_BGRA8 = 4 };

typedef struct {
  uint32_t struct_size;       /* sizeof(b200AprilTagsOptions_t), for forward compatibility */
  uint32_t family_mask;       /* default 1<<B200AT_FAM_36H11 */
  uint32_t max_batch;         /* frames per launch the workspace is sized for; default 1 */
  uint32_t max_tags;          /* per-frame output capacity; default 64 (node param max_tags, apriltag_node.cpp:564) */
  uint32_t tile_size;         /* default 4 (node param tile_size, :566) */
  float quad_decimate;        /* default 2.0  -- AprilRobotics apriltag_detector_create defaults below */
  float quad_sigma;           /* default 0.0 */
  int32_t refine_edges;       /* default 1 */
  double decode_sharpening;   /* default 0.25 */
  int32_t min_white_black_diff; /* default 5 */
  int32_t max_nmaxima;        /* default 10 */
  float critical_rad;         /* default 10 deg */
  float max_line_fit_mse;     /* default 10 */
  int32_t max_hamming;        /* default 2 */
  int32_t input_encoding;     /* default B200AT_ENC_BGR8 */
  int32_t device;             /* CUDA device ordinal, default -1 = current */
  /* bounded-buffer sizing, 0 = automatic from resolution */
  uint32_t hash_slots_per_frame;
  uint32_t points_per_frame;
  uint32_t clusters_per_frame;
  uint32_t quads_per_frame;
} b200AprilTagsOptions_t;

/* Extended per-detection record (AprilRobotics apriltag_detection_t + pose). */
typedef struct {
  int32_t family;         /* B200AT_FAM_* */
  int32_t id;
  int32_t hamming;
  float decision_margin;
  double H[9];            /* row major, tag [-1,1]^2 -> pixels */
  double c[2];
  double p[4][2];         /* AprilRobotics order */
  double R[9];            /* row major rotation, camera optical frame */
  double t[3];
  double pose_err;
} b200AprilTagsDetection_t;

typedef struct {
  const void *ptr;  /* device (or host, for the *Host entry points) pointer to the top-left pixel */
  size_t pitch;     /* bytes per row */
} b200AprilTagsFrame_t;

void b200AprilTagsDefaultOptions(b200AprilTagsOptions_t *opt);

/* Registers (copies) a caller-supplied family in slot B200AT_FAM_CUSTOM0 / B200AT_FAM_CUSTOM1, process wide; handles created
 * AFTERWARDS with that bit in family_mask decode it (existing handles keep the table they were created with). */
int b200AprilTagsRegisterFamily(int32_t slot, const b200AprilTagsFamilyDesc_t *desc);

int b200AprilTagsCreate(cuAprilTagsHandle *h, uint32_t img_width, uint32_t img_height,
                        const cuAprilTagsCameraIntrinsics_t *cam, float tag_dim, const b200AprilTagsOptions_t *opt);

int b200AprilTagsSetInputEncoding(cuAprilTagsHandle h, int32_t encoding);

/* Fused pre-stage (the reference's "AprilTag Graph" puts a rectify node and a resize node in front of the detector,
 * README.md:16-29, launch/isaac_ros_apriltag_usb_cam.launch.py:43-63): with a rectification set, the frames handed to the
 * DEVICE-pointer entry points (cuAprilTagsDetect, b200AprilTagsDetectBatch / EnqueueBatch) are RAW camera frames of
 * src_width x src_height in the handle's input encoding; one kernel undistorts, rectifies, resizes and converts them to gray into
 * an internal img_width x img_height image (the size the handle was created for), and detection runs on that.  The map is
 * OpenCV's initUndistortRectifyMap: for every output pixel (u, v): [x y w] = R^T * P^-1 * [u v 1]; distortion
 * (k1, k2, p1, p2, k3, k4, k5, k6: plumb_bob / rational_polynomial of sensor_msgs/CameraInfo) ; source = K * distorted point;
 * bilinear interpolation of the gray values of the four source pixels, constant 0 outside the source.  A resize is a P with scaled
 * focal lengths / principal point.  Detections (corners, centre, pose) are in the rectified output image, as in the reference's
 * graph; pass P's intrinsics to the create call.  NULL disables.  The host-buffer entry points return B200AT_ERR_UNSUPPORTED
 * while a rectification is set. */
typedef struct {
  uint32_t struct_size;   /* sizeof(b200AprilTagsRectify_t) */
  uint32_t src_width, src_height;
  double K[9];            /* raw camera matrix, row major */
  double D[8];            /* k1 k2 p1 p2 k3 k4 k5 k6 (unused = 0) */
  double R[9];            /* rectification rotation, row major (identity for a monocular camera) */
  double P[9];            /* camera matrix of the rectified (and resized) output, row major (the 3x3 part of CameraInfo.p) */
} b200AprilTagsRectify_t;
int b200AprilTagsSetRectification(cuAprilTagsHandle h, const b200AprilTagsRectify_t *rect);

/* Batch detect on DEVICE frames.  dets_out: HOST [n_frames][max_tags] (may be NULL), ids_out: HOST
 * [n_frames][max_tags] cuAprilTagsID_t (may be NULL), counts: HOST [n_frames].  Synchronous. */
int b200AprilTagsDetectBatch(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n_frames,
                             b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out, uint32_t *counts,
                             cudaStream_t stream);

/* Same, frames in HOST memory (pinned for full copy bandwidth): the H2D copies are issued inside, chunked and
 * overlapped with compute on an internal second stream.  n_frames may exceed max_batch. */
int b200AprilTagsDetectBatchHost(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n_frames,
                                 b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out, uint32_t *counts);

/* Asynchronous halves of DetectBatchHost: Enqueue validates every frame, queues the copies and kernels of the whole batch and
 * returns; Collect waits for the OLDEST batch in flight and unpacks it.  Up to TWO batches may be in flight per handle, so the
 * host->device copies and quad detection of batch k+1 overlap the decode / pose / device->host copy of batch k:
 *     Enqueue(b0); Enqueue(b1); Collect(b0); Enqueue(b2); Collect(b1); ...
 * The caller's frames must stay valid and unmodified until their batch has been collected.  Enqueue returns
 * B200AT_ERR_INVALID_ARG when two batches are already in flight; on any error nothing of the failed call is left running. */
int b200AprilTagsEnqueueBatchHost(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n_frames);
int b200AprilTagsCollectBatchHost(cuAprilTagsHandle h, b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out,
                                  uint32_t *counts);

/* Asynchronous halves of DetectBatch for pipelined callers: Enqueue launches the kernels and the D2H copy of the results on
 * `stream` and returns; Collect waits for the OLDEST batch in flight and unpacks it into host arrays.  Up to TWO batches may be
 * in flight per handle, both on the same stream (the second one is queued behind the first: the host side of a call -- frame
 * table, graph launch, unpacking -- then overlaps the previous batch's kernels instead of leaving the GPU idle between calls).
 * B200AT_ERR_INVALID_ARG when two batches are already in flight, when the stream differs from the one in flight, or with stage
 * timing enabled and a batch in flight.  b200AprilTagsReadBuffer shows the workspace of the batch queued LAST. */
int b200AprilTagsEnqueueBatch(cuAprilTagsHandle h, const b200AprilTagsFrame_t *frames, uint32_t n_frames,
                              cudaStream_t stream);
int b200AprilTagsCollectBatch(cuAprilTagsHandle h, b200AprilTagsDetection_t *dets_out, cuAprilTagsID_t *ids_out,
                              uint32_t *counts);

/* Status word of the last batch: bit0 hash table full, bit1 point pool full, bit2 cluster list full,
 * bit3 quad list full, bit4 candidate list full, bit5 output truncated to max_tags. */
int b200AprilTagsLastStatus(cuAprilTagsHandle h, uint32_t *status);

/* Per-stage device times (ms, CUDA events on the launch stream) of the last batch, when enabled. */
enum {
  B200AT_STAGE_PREPROCESS = 0, B200AT_STAGE_THRESHOLD, B200AT_STAGE_CCL, B200AT_STAGE_CLUSTER, B200AT_STAGE_QUADFIT,
  B200AT_STAGE_DECODE, B200AT_STAGE_FINALIZE, B200AT_STAGE_D2H, B200AT_NUM_STAGES
};
int b200AprilTagsEnableStageTiming(cuAprilTagsHandle h, int enable);
int b200AprilTagsGetStageTimes(cuAprilTagsHandle h, float *ms /* [B200AT_NUM_STAGES] */);
/* launches issued by the last Enqueue (for the bench's gpu_launches claim) and counters of the last batch */
int b200AprilTagsGetCounters(cuAprilTagsHandle h, uint64_t *counters /* [8]: launches, points, clusters, quads, candidates, detections, host->device bytes of the last DetectBatchHost, 1 if it staged rows sparsely */);

/* Intermediate buffers of the last batch, copied to HOST memory (parity tests). */
enum {
  B200AT_BUF_DECIMATED = 0, /* u8  [Hd][Wd] */
  B200AT_BUF_TILE_MIN,      /* u8  [th][tw] (after 3x3 erode) */
  B200AT_BUF_TILE_MAX,      /* u8  [th][tw] (after 3x3 dilate) */
  B200AT_BUF_THRESHOLD,     /* u8  [Hd][Wd] */
  B200AT_BUF_LABELS,        /* u32 [Hd][Wd] min-index representative */
  B200AT_BUF_SIZES,         /* u32 [Hd][Wd] component size stored at the representative's index */
  B200AT_BUF_CLUSTERS,      /* records {u64 key; u32 offset; u32 count; u32 frame; u32 pad}, all frames */
  B200AT_BUF_POINTS,        /* u64 sort keys per kept point (slope bits | y | x), cluster-contiguous, sorted */
  B200AT_BUF_QUADS,         /* records b200AprilTagsQuadRec_t, all frames */
  B200AT_BUF_QUADS_REFINED, /* same records after rescale + refine_edges */
  B200AT_BUF_POINTS_RAW,    /* u32 packed points as emitted: x | y<<14 | gx code<<28 | gy code<<30, cluster-contiguous */
  B200AT_BUF_RECTIFIED      /* u8  [H][W] gray output of the rectify / resize pre-stage (only with a rectification set) */
};
typedef struct {
  uint64_t key;
  float p[4][2];
  uint32_t frame;
  uint32_t reversed_border;
} b200AprilTagsQuadRec_t;
typedef struct {
  uint64_t key;
  uint32_t offset, count, frame, pad;
} b200AprilTagsClusterRec_t;
int b200AprilTagsGetDims(cuAprilTagsHandle h, uint32_t *wd, uint32_t *hd, uint32_t *tw, uint32_t *th);
/* returns number of elements available via *n_elems; copies min(cap_bytes, available) bytes */
int b200AprilTagsReadBuffer(cuAprilTagsHandle h, int which, uint32_t frame, void *dst, size_t cap_bytes, size_t *n_elems);

const char *b200AprilTagsVersion(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_APRILTAGS_H_ */
