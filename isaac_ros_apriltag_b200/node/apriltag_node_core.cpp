// apriltag_node_core.cpp -- see apriltag_node_core.hpp.  Structure follows the reference's strategy pattern
// (AprilTagImpl / CUAprilTagImpl / VPIAprilTagImpl, apriltag_node.cpp:93-130, 133-387, 389-559); every block
// cites the lines it mirrors.  Host C++ only; all detection work happens behind the C ABI.
#include "apriltag_node_core.hpp"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>

// the library writes cuAprilTagsID_t at the stride the reference node (CUDA headers: float2 is 8-byte aligned) reads it
static_assert(sizeof(cuAprilTagsID_t) == 88 && alignof(cuAprilTagsID_t) == 8 && offsetof(cuAprilTagsID_t, orientation) == 36,
              "cuAprilTagsID_t layout differs from libb200apriltags.so");

namespace nvidia {
namespace isaac_ros {
namespace apriltag {

namespace {

// apriltag_node.cpp:47-58 (the reference lists tag36h11 twice; a map keeps one)
enum FamilyId { FAM_INVALID = -1, FAM_36H11, FAM_16H5, FAM_25H9, FAM_36H10, FAM_CIRCLE21H7, FAM_CIRCLE49H12, FAM_CUSTOM48H12,
                FAM_STANDARD41H12, FAM_STANDARD52H13 };
const std::map<std::string, FamilyId> g_str_to_family = {
    {"tag36h11", FAM_36H11},           {"tag16h5", FAM_16H5},           {"tag25h9", FAM_25H9},
    {"tag36h10", FAM_36H10},           {"circle21h7", FAM_CIRCLE21H7},  {"circle49h12", FAM_CIRCLE49H12},
    {"custom48h12", FAM_CUSTOM48H12},  {"standard41h12", FAM_STANDARD41H12}, {"standard52h13", FAM_STANDARD52H13}};

FamilyId ToFamily(const std::string &s) {
  auto it = g_str_to_family.find(s);
  return it == g_str_to_family.end() ? FAM_INVALID : it->second;
}
std::string ToString(FamilyId f) {
  for (auto &kv : g_str_to_family)
    if (kv.second == f) return kv.first;
  return "";
}
// family -> C-ABI family bit, -1 when this build has no code table for it (the five families whose tables are not
// derivable offline: circle21h7, circle49h12, custom48h12, standard41h12, standard52h13)
int ToB200Family(FamilyId f) {
  switch (f) {
    case FAM_36H11: return B200AT_FAM_36H11;
    case FAM_25H9: return B200AT_FAM_25H9;
    case FAM_16H5: return B200AT_FAM_16H5;
    case FAM_36H10: return B200AT_FAM_36H10;
    default: return -1;
  }
}
// The five families the reference's VPI strategy lists but whose code tables this build cannot derive offline
// (circle21h7, circle49h12, custom48h12, standard41h12, standard52h13) are read from $B200AT_FAMILY_PATH/<family>.txt when the file
// exists -- the fields of upstream's tag<Family>.c as whitespace-separated text:
//   nbits ncodes width_at_border total_width reversed_border   bit_x[nbits]   bit_y[nbits]   codes[ncodes] (hex)
// -- and registered in slot B200AT_FAM_CUSTOM0.  Returns the family bit index, or -1.
int LoadFamilyFile(const std::string &name) {
  const char *dir = std::getenv("B200AT_FAMILY_PATH");
  if (!dir) return -1;
  // (C stdio on purpose: this library is loaded into processes that carry other C++ runtimes -- e.g. Python extension modules --
  // and iostream's locale facets do not survive two of them)
  FILE *f = std::fopen((std::string(dir) + "/" + name + ".txt").c_str(), "r");
  if (!f) return -1;
  struct Closer {
    FILE *f;
    ~Closer() { std::fclose(f); }
  } closer{f};
  long nbits = 0, ncodes = 0, wab = 0, tw = 0, rev = 0;
  if (std::fscanf(f, "%ld %ld %ld %ld %ld", &nbits, &ncodes, &wab, &tw, &rev) != 5 || nbits < 1 || nbits > 64 || ncodes < 1 || ncodes > (1 << 20)) return -1;
  std::vector<int8_t> bx(nbits), by(nbits);
  std::vector<uint64_t> codes(ncodes);
  for (long i = 0; i < nbits; i++) {
    long v;
    if (std::fscanf(f, "%ld", &v) != 1) return -1;
    bx[i] = (int8_t)v;
  }
  for (long i = 0; i < nbits; i++) {
    long v;
    if (std::fscanf(f, "%ld", &v) != 1) return -1;
    by[i] = (int8_t)v;
  }
  for (long i = 0; i < ncodes; i++) {
    char tok[64];
    if (std::fscanf(f, "%63s", tok) != 1) return -1;
    codes[i] = std::strtoull(tok, nullptr, 16);
  }
  b200AprilTagsFamilyDesc_t d;
  d.struct_size = sizeof(d);
  d.nbits = (uint32_t)nbits;
  d.ncodes = (uint32_t)ncodes;
  d.width_at_border = (uint32_t)wab;
  d.total_width = (uint32_t)tw;
  d.reversed_border = rev ? 1u : 0u;
  d.bit_x = bx.data();
  d.bit_y = by.data();
  d.codes = codes.data();
  return b200AprilTagsRegisterFamily(B200AT_FAM_CUSTOM0, &d) == 0 ? (int)B200AT_FAM_CUSTOM0 : -1;
}

// apriltag_node.cpp:76-82
int ToB200Encoding(const std::string &enc) {
  if (enc == "rgb8") return B200AT_ENC_RGB8;
  if (enc == "bgr8") return B200AT_ENC_BGR8;
  if (enc == "rgba8") return B200AT_ENC_RGBA8;
  if (enc == "bgra8") return B200AT_ENC_BGRA8;
  if (enc == "mono8") return B200AT_ENC_MONO8;
  return -1;
}

}  // namespace

uint32_t ParseBackends(const std::string &s) {
  uint32_t mask = 0;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    tok.erase(std::remove_if(tok.begin(), tok.end(), [](unsigned char c) { return std::isspace(c); }), tok.end());
    std::transform(tok.begin(), tok.end(), tok.begin(), [](unsigned char c) { return (char)std::toupper(c); });
    if (tok == "CPU") mask |= BACKEND_CPU;
    else if (tok == "CUDA") mask |= BACKEND_CUDA;
    else if (tok == "PVA") mask |= BACKEND_PVA;
    else if (tok == "VIC") mask |= BACKEND_VIC;
    else if (!tok.empty()) mask |= BACKEND_INVALID;
  }
  return mask;
}

Quaternion RotationToQuaternion(const float *m, bool col_major, bool normalize) {
  auto R = [&](int r, int c) -> float { return col_major ? m[c * 3 + r] : m[r * 3 + c]; };
  float w, x, y, z;
  float t = R(0, 0) + R(1, 1) + R(2, 2);
  if (t > 0.0f) {
    t = std::sqrt(t + 1.0f);
    w = 0.5f * t;
    t = 0.5f / t;
    x = (R(2, 1) - R(1, 2)) * t;
    y = (R(0, 2) - R(2, 0)) * t;
    z = (R(1, 0) - R(0, 1)) * t;
  } else {
    int i = 0;
    if (R(1, 1) > R(0, 0)) i = 1;
    if (R(2, 2) > R(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0f);
    float q[3];
    q[i] = 0.5f * t;
    t = 0.5f / t;
    w = (R(k, j) - R(j, k)) * t;
    q[j] = (R(j, i) + R(i, j)) * t;
    q[k] = (R(k, i) + R(i, k)) * t;
    x = q[0];
    y = q[1];
    z = q[2];
  }
  if (normalize) {
    float n = std::sqrt(w * w + x * x + y * y + z * z);
    if (n > 0) {
      w /= n;
      x /= n;
      y /= n;
      z /= n;
    }
  }
  Quaternion o;
  o.w = w;
  o.x = x;
  o.y = y;
  o.z = z;
  return o;
}

// ---- apriltag_node.cpp:93-130 ----
struct AprilTagNodeCore::AprilTagImpl {
  virtual ~AprilTagImpl() = default;
  cudaStream_t stream_ = nullptr;  // created in Initialize, destroyed in the strategy's destructor iff initialised (:460, :552-558)
  bool IsInitialized() const { return initialized_; }
  virtual std::unordered_set<int> SupportedTagFamilies() const = 0;
  virtual void Initialize(const AprilTagNodeCore &node, const ImageView &, const CameraInfo &) {
    initialized_ = true;
    tag_family_str_ = node.params_.tag_family;
    tag_family_ = ToFamily(tag_family_str_);
  }
  virtual void OnCameraFrame(AprilTagNodeCore &node, const ImageView &image, const CameraInfo &camera_info) = 0;
  bool initialized_{false};
  FamilyId tag_family_{FAM_INVALID};
  std::string tag_family_str_{};
};

// ---- apriltag_node.cpp:389-559: the default backend, behind the cuAprilTags-shaped entry points ----
struct AprilTagNodeCore::CUAprilTagImpl : AprilTagNodeCore::AprilTagImpl {
  cuAprilTagsHandle detector_ = nullptr;
  cuAprilTagsCameraIntrinsics_t cam_intrinsics_{};

  std::unordered_set<int> SupportedTagFamilies() const override { return {FAM_36H11}; }  // :429-432

  void Initialize(const AprilTagNodeCore &node, const ImageView &image, const CameraInfo &camera_info) override {
    AprilTagImpl::Initialize(node, image, camera_info);
    const double *k = camera_info.k.data();  // :442-447
    cam_intrinsics_ = {static_cast<float>(k[0]), static_cast<float>(k[4]), static_cast<float>(k[2]), static_cast<float>(k[5])};
    const int error = nvCreateAprilTagsDetector(&detector_, camera_info.width, camera_info.height, node.params_.tile_size,
                                                NVAT_TAG36H11, &cam_intrinsics_, static_cast<float>(node.params_.size));
    if (error != 0) {  // :453-457
      initialized_ = false;
      throw std::runtime_error("Failed to create cuAprilTags detector (error code " + std::to_string(error) + ")");
    }
    cudaStreamCreate(&stream_);  // :460 "Create stream for detection"
  }

  void OnCameraFrame(AprilTagNodeCore &node, const ImageView &image, const CameraInfo &camera_info) override {
    if (image.encoding != "rgb8" && image.encoding != "bgr8") {  // :469-476
      node.Log(1, "Unsupported image encoding: " + image.encoding + " (only 'rgb8' or 'bgr8' supported)");
      throw std::runtime_error("cuAprilTags detector only supports 'rgb8' or 'bgr8' image input");
    }
    b200AprilTagsSetInputEncoding(detector_, ToB200Encoding(image.encoding));
    cuAprilTagsImageInput_t input_image;  // :481-486
    input_image.width = static_cast<uint16_t>(image.width);
    input_image.height = static_cast<uint16_t>(image.height);
    input_image.dev_ptr = const_cast<uchar3 *>(reinterpret_cast<const uchar3 *>(image.dev_ptr));
    input_image.pitch = image.step;
    uint32_t num_detections = 0;
    std::vector<cuAprilTagsID_t> tags(node.params_.max_tags);  // :490
    const int error = (int)cuAprilTagsDetect(detector_, &input_image, tags.data(), &num_detections, node.params_.max_tags, stream_);  // :491-493
    if (error != 0) {  // :494-497 log and drop the frame
      node.Log(1, "Failed to run AprilTags detector (error code " + std::to_string(error) + ")");
      return;
    }
    AprilTagDetectionArray msg_detections;
    msg_detections.header = camera_info.header;
    std::vector<TransformStamped> tfs;
    for (uint32_t i = 0; i < num_detections; i++) {
      const cuAprilTagsID_t &detection = tags[i];
      AprilTagDetection msg_detection;
      msg_detection.family = tag_family_str_;
      msg_detection.id = detection.id;
      for (int c = 0; c < 4; c++) {  // :512-517 corners 1:1
        msg_detection.corners[c].x = detection.corners[c].x;
        msg_detection.corners[c].y = detection.corners[c].y;
      }
      // :520-530 centre = intersection of the diagonals 0-2 and 1-3 (parametric form: no division by zero for
      // vertical diagonals, same point otherwise)
      {
        const float x0 = detection.corners[0].x, y0 = detection.corners[0].y, x2 = detection.corners[2].x, y2 = detection.corners[2].y;
        const float x1 = detection.corners[1].x, y1 = detection.corners[1].y, x3 = detection.corners[3].x, y3 = detection.corners[3].y;
        const float d1x = x2 - x0, d1y = y2 - y0, d2x = x3 - x1, d2y = y3 - y1;
        const float den = d1x * d2y - d1y * d2x;
        const float s = ((x1 - x0) * d2y - (y1 - y0) * d2x) / den;
        msg_detection.center.x = x0 + s * d1x;
        msg_detection.center.y = y0 + s * d1y;
      }
      TransformStamped tf;  // :533-538
      tf.header = camera_info.header;
      tf.child_frame_id = tag_family_str_ + ":" + std::to_string(detection.id);
      tf.transform.translation.x = detection.translation[0];
      tf.transform.translation.y = detection.translation[1];
      tf.transform.translation.z = detection.translation[2];
      tf.transform.rotation = RotationToQuaternion(detection.orientation, /*col_major=*/true, /*normalize=*/false);  // :409-427
      tfs.push_back(tf);
      msg_detection.pose.position = tf.transform.translation;
      msg_detection.pose.orientation = tf.transform.rotation;
      msg_detections.detections.push_back(msg_detection);
    }
    node.Publish(msg_detections, tfs);  // :548-549
  }

  ~CUAprilTagImpl() override {  // :552-558
    if (stream_) cudaStreamDestroy(stream_);
    if (detector_) cuAprilTagsDestroy(detector_);
  }
};

// ---- apriltag_node.cpp:133-387: the "any other backends" strategy (VPI in the reference): all families with a
// code table, all five encodings, centre from the library, normalised quaternion ----
struct AprilTagNodeCore::VPIAprilTagImpl : AprilTagNodeCore::AprilTagImpl {
  cuAprilTagsHandle detector_ = nullptr;

  std::unordered_set<int> SupportedTagFamilies() const override {  // :182-191
    return {FAM_16H5, FAM_25H9, FAM_36H10, FAM_36H11, FAM_CIRCLE21H7, FAM_CIRCLE49H12, FAM_CUSTOM48H12, FAM_STANDARD41H12,
            FAM_STANDARD52H13};
  }

  void Initialize(const AprilTagNodeCore &node, const ImageView &image, const CameraInfo &camera_info) override {
    AprilTagImpl::Initialize(node, image, camera_info);
    int fam = ToB200Family(tag_family_);
    if (fam < 0) fam = LoadFamilyFile(tag_family_str_);
    if (fam < 0) {
      initialized_ = false;
      throw std::runtime_error("Failed to create AprilTag detector: no code table for family '" + tag_family_str_ +
                               "' in this build (supply one as $B200AT_FAMILY_PATH/" + tag_family_str_ + ".txt)");
    }
    const int enc = ToB200Encoding(image.encoding);
    if (enc < 0) {
      initialized_ = false;
      throw std::runtime_error("Unsupported image encoding: " + image.encoding);
    }
    b200AprilTagsOptions_t opt;
    b200AprilTagsDefaultOptions(&opt);
    opt.family_mask = 1u << fam;
    opt.max_tags = (uint32_t)node.params_.max_tags;
    opt.tile_size = node.params_.tile_size;
    opt.input_encoding = enc;
    const double *k = camera_info.k.data();  // :215-225 (the skew term k[1] is not used by the pose stage)
    cuAprilTagsCameraIntrinsics_t cam = {static_cast<float>(k[0]), static_cast<float>(k[4]), static_cast<float>(k[2]),
                                         static_cast<float>(k[5])};
    const int error = b200AprilTagsCreate(&detector_, camera_info.width, camera_info.height, &cam, (float)node.params_.size, &opt);
    if (error != 0) {
      initialized_ = false;
      throw std::runtime_error("Failed to create AprilTag detector (error code " + std::to_string(error) + ")");
    }
    cudaStreamCreate(&stream_);  // :211 (the VPI strategy's cuda_stream_)
    node.Log(0, "AprilTag detector: " + std::to_string(camera_info.width) + "x" + std::to_string(camera_info.height) + " " + tag_family_str_);
  }

  void OnCameraFrame(AprilTagNodeCore &node, const ImageView &image, const CameraInfo &camera_info) override {
    const int enc = ToB200Encoding(image.encoding);
    if (enc < 0) throw std::runtime_error("Unsupported image encoding: " + image.encoding);
    b200AprilTagsSetInputEncoding(detector_, enc);
    b200AprilTagsFrame_t fr{image.dev_ptr, image.step};
    std::vector<b200AprilTagsDetection_t> dets(node.params_.max_tags);
    uint32_t n = 0;
    const int error = b200AprilTagsDetectBatch(detector_, &fr, 1, dets.data(), nullptr, &n, stream_);
    if (error != 0 && error != B200AT_ERR_OVERFLOW) {
      node.Log(1, "Failed to run AprilTags detector (error code " + std::to_string(error) + ")");
      return;
    }
    AprilTagDetectionArray msg_detections;
    msg_detections.header = camera_info.header;
    std::vector<TransformStamped> tfs;
    for (uint32_t i = 0; i < n; i++) {
      const b200AprilTagsDetection_t &d = dets[i];
      AprilTagDetection m;
      m.family = tag_family_str_;
      m.id = d.id;
      // :337-344: the library's corner order is reversed into message order (dest = 3 - idx)
      for (int c = 0; c < 4; c++) {
        m.corners[3 - c].x = (float)d.p[c][0];
        m.corners[3 - c].y = (float)d.p[c][1];
      }
      m.center.x = (float)d.c[0];  // :347-348
      m.center.y = (float)d.c[1];
      TransformStamped tf;
      tf.header = camera_info.header;
      tf.child_frame_id = tag_family_str_ + ":" + std::to_string(d.id);  // :354
      float R[9];
      for (int k = 0; k < 9; k++) R[k] = (float)d.R[k];
      tf.transform.rotation = RotationToQuaternion(R, /*col_major=*/false, /*normalize=*/true);  // :147-180
      tf.transform.translation.x = (float)d.t[0];
      tf.transform.translation.y = (float)d.t[1];
      tf.transform.translation.z = (float)d.t[2];
      tfs.push_back(tf);
      m.pose.position = tf.transform.translation;
      m.pose.orientation = tf.transform.rotation;
      msg_detections.detections.push_back(m);
    }
    node.Publish(msg_detections, tfs);
  }

  ~VPIAprilTagImpl() override {
    if (stream_) cudaStreamDestroy(stream_);
    if (detector_) cuAprilTagsDestroy(detector_);
  }
};

// ---- apriltag_node.cpp:562-611 ----
AprilTagNodeCore::AprilTagNodeCore(const NodeParams &params, DetectionsSink det, TfSink tf, LogSink log)
    : params_(params), backends_(params.backends_mask ? params.backends_mask : ParseBackends(params.backends)), det_sink_(std::move(det)), tf_sink_(std::move(tf)),
      log_sink_(std::move(log)) {
  if (backends_ == BACKEND_CUDA) {  // :576-582
    Log(0, "Using cuAprilTag implementation.");
    impl_ = std::make_unique<CUAprilTagImpl>();
  } else {
    Log(0, "Using VPI implementation.");
    impl_ = std::make_unique<VPIAprilTagImpl>();
  }
  auto supported = impl_->SupportedTagFamilies();  // :584-599
  if (supported.find(ToFamily(params_.tag_family)) == supported.end()) {
    std::ostringstream os;
    os << "Tag family not supported by specified backend: '" << params_.tag_family << "'" << std::endl;
    os << "'tag_family' parameter must be one of:" << std::endl;
    for (int f : supported) os << ToString((FamilyId)f) << std::endl;
    Log(2, "Tag family not supported by specified backend: '" + params_.tag_family + "'");
    throw std::runtime_error(os.str());
  }
}

AprilTagNodeCore::~AprilTagNodeCore() = default;

bool AprilTagNodeCore::UsingCuAprilTagImpl() const { return backends_ == BACKEND_CUDA; }

cudaStream_t AprilTagNodeCore::cuda_stream() const { return impl_ ? impl_->stream_ : nullptr; }

void AprilTagNodeCore::CameraImageCallback(const ImageView &image, const CameraInfo &camera_info) {  // :613-623
  if (!impl_->IsInitialized()) impl_->Initialize(*this, image, camera_info);
  impl_->OnCameraFrame(*this, image, camera_info);
}

void AprilTagNodeCore::Publish(const AprilTagDetectionArray &d, const std::vector<TransformStamped> &t) {
  last_detections_ = d;
  last_tfs_ = t;
  if (det_sink_) det_sink_(d);
  if (tf_sink_) tf_sink_(t);
}

void AprilTagNodeCore::Log(int level, const std::string &m) const {
  if (log_sink_) log_sink_(level, m);
}

}  // namespace apriltag
}  // namespace isaac_ros
}  // namespace nvidia
