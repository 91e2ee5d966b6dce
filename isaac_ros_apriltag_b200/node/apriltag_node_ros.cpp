// apriltag_node_ros.cpp -- rclcpp component wrapper around the node core: the plugin surface of
// /root/reference/isaac_ros_apriltag/src/apriltag_node.cpp:562-633 (plugin "nvidia::isaac_ros::apriltag::AprilTagNode",
// subscriptions `image` + `camera_info` with ExactTime(3), publisher `tag_detections` QoS 1, /tf).
//
// NOT BUILT IN THIS IMAGE: ROS 2 (rclcpp, message_filters, tf2_ros, isaac_ros_nitros, isaac_ros_apriltag_interfaces)
// is absent, so this translation unit is compiled only where B200_APRILTAG_WITH_ROS is defined by an ament build
// (see INTEGRATION.md).  It only converts messages; every behaviour lives in apriltag_node_core.cpp, which IS built
// and tested here.
#ifdef B200_APRILTAG_WITH_ROS
#include "apriltag_node_core.hpp"
#include "geometry_msgs/msg/transform_stamped.hpp"
#include "isaac_ros_apriltag_interfaces/msg/april_tag_detection_array.hpp"
#include "isaac_ros_nitros/types/nitros_type_message_filter_traits.hpp"
#include "isaac_ros_nitros_image_type/nitros_image.hpp"
#include "isaac_ros_vpi_utils/vpi_utilities.hpp"
#include "message_filters/subscriber.h"
#include "message_filters/sync_policies/exact_time.h"
#include "message_filters/synchronizer.h"
#include "rclcpp/rclcpp.hpp"
#include "sensor_msgs/msg/camera_info.hpp"
#include "tf2_ros/transform_broadcaster.h"

namespace nvidia {
namespace isaac_ros {
namespace apriltag {

namespace core = ::nvidia::isaac_ros::apriltag;

class AprilTagNode : public rclcpp::Node {
 public:
  explicit AprilTagNode(const rclcpp::NodeOptions &options = rclcpp::NodeOptions())
      : rclcpp::Node("apriltag_node", options),
        camera_image_sync_{ExactPolicy{3}, image_sub_, camera_info_sub_},
        detections_pub_{create_publisher<isaac_ros_apriltag_interfaces::msg::AprilTagDetectionArray>("tag_detections", rclcpp::QoS(1))} {
    core::NodeParams p;
    p.max_tags = declare_parameter<int>("max_tags", 64);
    p.size = declare_parameter<double>("size", 0.22);
    p.tile_size = declare_parameter<uint16_t>("tile_size", 4);
    p.tag_family = declare_parameter<std::string>("tag_family", "tag36h11");
    // apriltag_node.cpp:568: the `backends` parameter is declared and parsed by isaac_ros_vpi_utils, which returns VPIBackend flags
    p.backends_mask = static_cast<uint32_t>(nvidia::isaac_ros::vpi_utils::DeclareVPIBackendParameter(this, VPI_BACKEND_CUDA));
    core_ = std::make_unique<core::AprilTagNodeCore>(
        p, [this](const core::AprilTagDetectionArray &a) { PublishDetections(a); },
        [this](const std::vector<core::TransformStamped> &t) { PublishTf(t); },
        [this](int lvl, const std::string &m) {
          if (lvl == 0) RCLCPP_INFO(get_logger(), "%s", m.c_str());
          else if (lvl == 1) RCLCPP_ERROR(get_logger(), "%s", m.c_str());
          else RCLCPP_FATAL(get_logger(), "%s", m.c_str());
        });
    tf_broadcaster_ = std::make_unique<tf2_ros::TransformBroadcaster>(this);
    camera_image_sync_.registerCallback(std::bind(&AprilTagNode::Callback, this, std::placeholders::_1, std::placeholders::_2));
    image_sub_.subscribe(this, "image");
    camera_info_sub_.subscribe(this, "camera_info");
  }

 private:
  void Callback(const nvidia::isaac_ros::nitros::NitrosImage::ConstSharedPtr &img,
                const sensor_msgs::msg::CameraInfo::ConstSharedPtr &ci) {
    // kept alive across the call, on the strategy's own stream (apriltag_node.cpp:479-480; nullptr only before the first frame,
    // when the strategy has not created its stream yet)
    auto read_handle = img->get_read_handle(core_->cuda_stream());
    core::ImageView v{img->encoding, img->width, img->height, img->step, read_handle.get_ptr()};
    core::CameraInfo c;
    c.header.stamp_sec = ci->header.stamp.sec;
    c.header.stamp_nanosec = ci->header.stamp.nanosec;
    c.header.frame_id = ci->header.frame_id;
    c.width = ci->width;
    c.height = ci->height;
    for (int i = 0; i < 9; i++) c.k[i] = ci->k[i];
    header_ = ci->header;
    core_->CameraImageCallback(v, c);
  }
  void PublishDetections(const core::AprilTagDetectionArray &a) {
    isaac_ros_apriltag_interfaces::msg::AprilTagDetectionArray m;
    m.header = header_;
    for (const auto &d : a.detections) {
      isaac_ros_apriltag_interfaces::msg::AprilTagDetection o;
      o.family = d.family;
      o.id = d.id;
      o.center.x = d.center.x;
      o.center.y = d.center.y;
      for (int k = 0; k < 4; k++) {
        o.corners.data()[k].x = d.corners[k].x;
        o.corners.data()[k].y = d.corners[k].y;
      }
      o.pose.pose.pose.position.x = d.pose.position.x;
      o.pose.pose.pose.position.y = d.pose.position.y;
      o.pose.pose.pose.position.z = d.pose.position.z;
      o.pose.pose.pose.orientation.x = d.pose.orientation.x;
      o.pose.pose.pose.orientation.y = d.pose.orientation.y;
      o.pose.pose.pose.orientation.z = d.pose.orientation.z;
      o.pose.pose.pose.orientation.w = d.pose.orientation.w;
      m.detections.push_back(o);
    }
    detections_pub_->publish(m);
  }
  void PublishTf(const std::vector<core::TransformStamped> &tfs) {
    std::vector<geometry_msgs::msg::TransformStamped> out;
    for (const auto &t : tfs) {
      geometry_msgs::msg::TransformStamped o;
      o.header = header_;
      o.child_frame_id = t.child_frame_id;
      o.transform.translation.x = t.transform.translation.x;
      o.transform.translation.y = t.transform.translation.y;
      o.transform.translation.z = t.transform.translation.z;
      o.transform.rotation.x = t.transform.rotation.x;
      o.transform.rotation.y = t.transform.rotation.y;
      o.transform.rotation.z = t.transform.rotation.z;
      o.transform.rotation.w = t.transform.rotation.w;
      out.push_back(o);
    }
    tf_broadcaster_->sendTransform(out);
  }
  message_filters::Subscriber<nvidia::isaac_ros::nitros::NitrosImage> image_sub_;
  message_filters::Subscriber<sensor_msgs::msg::CameraInfo> camera_info_sub_;
  using ExactPolicy = message_filters::sync_policies::ExactTime<nvidia::isaac_ros::nitros::NitrosImage, sensor_msgs::msg::CameraInfo>;
  message_filters::Synchronizer<ExactPolicy> camera_image_sync_;
  rclcpp::Publisher<isaac_ros_apriltag_interfaces::msg::AprilTagDetectionArray>::SharedPtr detections_pub_;
  std::unique_ptr<tf2_ros::TransformBroadcaster> tf_broadcaster_;
  std::unique_ptr<core::AprilTagNodeCore> core_;
  std_msgs::msg::Header header_;
};

}  // namespace apriltag
}  // namespace isaac_ros
}  // namespace nvidia

#include "rclcpp_components/register_node_macro.hpp"
// the plugin name every launch file of the reference asks for (launch/isaac_ros_apriltag.launch.py:26, CMakeLists.txt:46-48,
// apriltag_node.cpp:633)
RCLCPP_COMPONENTS_REGISTER_NODE(nvidia::isaac_ros::apriltag::AprilTagNode)
#endif  // B200_APRILTAG_WITH_ROS
