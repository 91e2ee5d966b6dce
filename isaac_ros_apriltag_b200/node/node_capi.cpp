// node_capi.cpp -- flat C entry points over AprilTagNodeCore (node core) so tests and non-C++ hosts can drive the
// plugin surface without ROS.  Exceptions are converted to error codes + message (none crosses the ABI).
#include <cstring>
#include <string>

#include "apriltag_node_core.hpp"

using namespace nvidia::isaac_ros::apriltag;

extern "C" {

struct b200NodeDetectionMsg {
  char family[32];
  int32_t id;
  double center[2];
  double corners[4][2];
  double position[3];
  double orientation_xyzw[4];
  char child_frame_id[48];
};

static void set_err(char *err, size_t n, const std::string &m) {
  if (err && n) {
    std::strncpy(err, m.c_str(), n - 1);
    err[n - 1] = 0;
  }
}

// returns 0 on success; 1 when the constructor throws (message in err), like the reference's gtests expect
int b200NodeCreate(void **node, const char *tag_family, const char *backends, double size, int max_tags, int tile_size,
                   char *err, size_t errlen) {
  if (!node) return 2;
  *node = nullptr;
  try {
    NodeParams p;
    if (tag_family) p.tag_family = tag_family;
    if (backends) p.backends = backends;
    p.size = size;
    p.max_tags = max_tags;
    p.tile_size = (uint32_t)tile_size;
    *node = new AprilTagNodeCore(p);
    return 0;
  } catch (const std::exception &e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

void b200NodeDestroy(void *node) { delete static_cast<AprilTagNodeCore *>(node); }

int b200NodeUsingCuAprilTagImpl(void *node) { return static_cast<AprilTagNodeCore *>(node)->UsingCuAprilTagImpl() ? 1 : 0; }

// one synchronised (image, camera_info) pair; out receives the published AprilTagDetectionArray
int b200NodeOnFrame(void *node, const char *encoding, uint32_t width, uint32_t height, uint32_t step, const void *dev_ptr,
                    const double *K9, uint32_t ci_width, uint32_t ci_height, const char *frame_id, b200NodeDetectionMsg *out,
                    int max_out, int *n_out, char *err, size_t errlen) {
  try {
    AprilTagNodeCore *n = static_cast<AprilTagNodeCore *>(node);
    ImageView im;
    im.encoding = encoding ? encoding : "";
    im.width = width;
    im.height = height;
    im.step = step;
    im.dev_ptr = dev_ptr;
    CameraInfo ci;
    ci.width = ci_width;
    ci.height = ci_height;
    ci.header.frame_id = frame_id ? frame_id : "";
    for (int i = 0; i < 9; i++) ci.k[i] = K9[i];
    n->CameraImageCallback(im, ci);
    const auto &msg = n->last_detections();
    const auto &tfs = n->last_transforms();
    int cnt = 0;
    for (size_t i = 0; i < msg.detections.size() && cnt < max_out; i++, cnt++) {
      const auto &d = msg.detections[i];
      b200NodeDetectionMsg &o = out[cnt];
      std::memset(&o, 0, sizeof(o));
      std::strncpy(o.family, d.family.c_str(), sizeof(o.family) - 1);
      o.id = d.id;
      o.center[0] = d.center.x;
      o.center[1] = d.center.y;
      for (int c = 0; c < 4; c++) {
        o.corners[c][0] = d.corners[c].x;
        o.corners[c][1] = d.corners[c].y;
      }
      o.position[0] = d.pose.position.x;
      o.position[1] = d.pose.position.y;
      o.position[2] = d.pose.position.z;
      o.orientation_xyzw[0] = d.pose.orientation.x;
      o.orientation_xyzw[1] = d.pose.orientation.y;
      o.orientation_xyzw[2] = d.pose.orientation.z;
      o.orientation_xyzw[3] = d.pose.orientation.w;
      if (i < tfs.size()) std::strncpy(o.child_frame_id, tfs[i].child_frame_id.c_str(), sizeof(o.child_frame_id) - 1);
    }
    if (n_out) *n_out = cnt;
    return 0;
  } catch (const std::exception &e) {
    set_err(err, errlen, e.what());
    return 1;
  }
}

// exposed for unit tests of the marshalling helpers
void b200NodeRotationToQuaternion(const float *m9, int col_major, int normalize, double *xyzw) {
  Quaternion q = RotationToQuaternion(m9, col_major != 0, normalize != 0);
  xyzw[0] = q.x;
  xyzw[1] = q.y;
  xyzw[2] = q.z;
  xyzw[3] = q.w;
}
uint32_t b200NodeParseBackends(const char *s) { return ParseBackends(s ? s : ""); }

}  // extern "C"
