"""isaac_ros_apriltag_b200 -- B200-native AprilTag detection hot path behind the isaac_ros_apriltag
plugin boundary.  See DESIGN.md."""
__version__ = "0.1.0"
