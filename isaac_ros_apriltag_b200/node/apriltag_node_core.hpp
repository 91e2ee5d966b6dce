// apriltag_node_core.hpp -- AprilTagNodeCore: ROS-free mirror of nvidia::isaac_ros::apriltag::AprilTagNode
// (/root/reference/isaac_ros_apriltag/include/isaac_ros_apriltag/apriltag_node.hpp:48-91,
//  /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:93-130, 389-559, 562-623).
//
// Same parameters (max_tags=64, size=0.22, tile_size=4, tag_family="tag36h11", backends="CUDA"), same backend
// selection rule, same family validation and error strings, same per-frame marshalling (corner order, centre,
// quaternion, tf child frame "<family>:<id>").  ROS types are replaced by plain structs with the same field
// names; the rclcpp wrapper (apriltag_node_ros.cpp) only converts messages.  The detector behind both
// backend strategies is libb200apriltags.so (include/b200_apriltags.h); there is no CPU detector in the product.
#pragma once
#include <array>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

// compiled against the CUDA headers like the reference node (float2 / uchar3 / cudaStream_t are the real types)
#define B200_APRILTAGS_USE_CUDA_HEADERS 1
#include "../../include/b200_apriltags.h"

namespace nvidia {
namespace isaac_ros {
namespace apriltag {

// --- message mirrors (field names follow the ROS IDL the reference publishes) ---
struct Header {
  int32_t stamp_sec = 0;
  uint32_t stamp_nanosec = 0;
  std::string frame_id;
};
struct Point2 {
  double x = 0, y = 0;
};
struct Vector3 {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Transform {
  Vector3 translation;
  Quaternion rotation;
};
struct TransformStamped {
  Header header;
  std::string child_frame_id;
  Transform transform;
};
struct Pose {
  Vector3 position;
  Quaternion orientation;
};
struct AprilTagDetection {  // isaac_ros_apriltag_interfaces/msg/AprilTagDetection
  std::string family;
  int32_t id = 0;
  Point2 center;
  std::array<Point2, 4> corners;
  Pose pose;  // msg.pose.pose.pose
};
struct AprilTagDetectionArray {
  Header header;
  std::vector<AprilTagDetection> detections;
};
struct CameraInfo {  // sensor_msgs/msg/CameraInfo (fields the node reads)
  Header header;
  uint32_t width = 0, height = 0;
  std::array<double, 9> k{};
};
// View of a NitrosImage (apriltag_node.cpp:237-245, 480-486): encoding, geometry and a DEVICE pointer.
struct ImageView {
  std::string encoding;
  uint32_t width = 0, height = 0, step = 0;
  const void *dev_ptr = nullptr;
};

// backends bitmask, as isaac_ros_vpi_utils::DeclareVPIBackendParameter produces (VPI_BACKEND_* values)
enum : uint32_t { BACKEND_CPU = 1u << 0, BACKEND_CUDA = 1u << 1, BACKEND_PVA = 1u << 2, BACKEND_VIC = 1u << 3, BACKEND_INVALID = 1u << 31 };
uint32_t ParseBackends(const std::string &s);  // "CUDA", "CPU", "PVA", "CPU,CUDA", ...

struct NodeParams {
  int max_tags = 64;                    // apriltag_node.cpp:564
  double size = 0.22;                   // :565
  uint32_t tile_size = 4;               // :566
  std::string tag_family = "tag36h11";  // :567
  std::string backends = "CUDA";        // :568 (the string DeclareVPIBackendParameter parses) ...
  uint32_t backends_mask = 0;           // ... or, when non-zero, the VPIBackend flags it returned (the rclcpp wrapper passes these)
};

// 3x3 rotation -> quaternion exactly as Eigen::Quaternion<float>(Matrix3f) does (trace method with the
// largest-diagonal fallback); col_major selects the cuAprilTags layout (apriltag_node.cpp:409-427).
Quaternion RotationToQuaternion(const float *m, bool col_major, bool normalize);

class AprilTagNodeCore {
 public:
  using DetectionsSink = std::function<void(const AprilTagDetectionArray &)>;
  using TfSink = std::function<void(const std::vector<TransformStamped> &)>;
  using LogSink = std::function<void(int level, const std::string &)>;  // 0 info, 1 error, 2 fatal

  explicit AprilTagNodeCore(const NodeParams &params, DetectionsSink det = nullptr, TfSink tf = nullptr, LogSink log = nullptr);
  ~AprilTagNodeCore();
  AprilTagNodeCore(const AprilTagNodeCore &) = delete;
  AprilTagNodeCore &operator=(const AprilTagNodeCore &) = delete;

  // apriltag_node.cpp:613-623: lazy Initialize on the first frame, then OnCameraFrame.
  void CameraImageCallback(const ImageView &image, const CameraInfo &camera_info);

  const NodeParams &params() const { return params_; }
  bool UsingCuAprilTagImpl() const;
  // The strategy's own CUDA stream (apriltag_node.cpp:460, :211): created at the first frame like the reference's stream_, nullptr
  // before.  The rclcpp wrapper hands it to NitrosImage::get_read_handle so the image is ordered against the detector's work.
  cudaStream_t cuda_stream() const;
  // last published messages (also delivered to the sinks)
  const AprilTagDetectionArray &last_detections() const { return last_detections_; }
  const std::vector<TransformStamped> &last_transforms() const { return last_tfs_; }

  struct AprilTagImpl;
  struct CUAprilTagImpl;
  struct VPIAprilTagImpl;

 private:
  friend struct AprilTagImpl;
  friend struct CUAprilTagImpl;
  friend struct VPIAprilTagImpl;
  void Publish(const AprilTagDetectionArray &d, const std::vector<TransformStamped> &t);
  void Log(int level, const std::string &m) const;
  const NodeParams params_;
  const uint32_t backends_;
  DetectionsSink det_sink_;
  TfSink tf_sink_;
  LogSink log_sink_;
  AprilTagDetectionArray last_detections_;
  std::vector<TransformStamped> last_tfs_;
  std::unique_ptr<AprilTagImpl> impl_;
};

}  // namespace apriltag
}  // namespace isaac_ros
}  // namespace nvidia
