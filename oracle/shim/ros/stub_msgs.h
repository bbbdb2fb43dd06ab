// oracle/shim/ros/stub_msgs.h -- TEST INFRASTRUCTURE ONLY.
// Plain-struct stand-ins for the handful of ROS message / tf2 types that the
// reference's hot-path include closure mentions.  They let
// the unmodified reference headers parse and link without ROS; none of them
// takes part in the arithmetic that is being checked.
#ifndef ERGODIC_SHIM_ROS_STUB_MSGS_H
#define ERGODIC_SHIM_ROS_STUB_MSGS_H
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace ros
{
struct Duration
{
  double sec;
  Duration(double s = 0.0) : sec(s) {}
};
struct Time
{
  double sec = 0.0;
};
}  // namespace ros

namespace std_msgs
{
struct Header
{
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
struct ColorRGBA
{
  float r = 0, g = 0, b = 0, a = 0;
};
}  // namespace std_msgs

namespace geometry_msgs
{
struct Point
{
  double x = 0, y = 0, z = 0;
};
struct Vector3
{
  double x = 0, y = 0, z = 0;
};
struct Quaternion
{
  double x = 0, y = 0, z = 0, w = 1;
};
struct Pose
{
  Point position;
  Quaternion orientation;
};
struct PoseStamped
{
  std_msgs::Header header;
  Pose pose;
};
}  // namespace geometry_msgs

namespace nav_msgs
{
struct Path
{
  std_msgs::Header header;
  std::vector<geometry_msgs::PoseStamped> poses;
};
struct MapMetaData
{
  ros::Time map_load_time;
  float resolution = 0;
  uint32_t width = 0, height = 0;
  geometry_msgs::Pose origin;
};
struct OccupancyGrid
{
  typedef std::shared_ptr<OccupancyGrid> Ptr;
  typedef std::shared_ptr<const OccupancyGrid> ConstPtr;
  std_msgs::Header header;
  MapMetaData info;
  std::vector<int8_t> data;
};
}  // namespace nav_msgs

namespace visualization_msgs
{
struct Marker
{
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  std_msgs::Header header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  std_msgs::ColorRGBA color;
  ros::Duration lifetime;
};
struct MarkerArray
{
  std::vector<Marker> markers;
};
}  // namespace visualization_msgs

namespace tf2
{
class Quaternion
{
public:
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  void setRPY(double roll, double pitch, double yaw)
  {
    const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5);
    const double cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
    const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
    x_ = sr * cp * cy - cr * sp * sy;
    y_ = cr * sp * cy + sr * cp * sy;
    z_ = cr * cp * sy - sr * sp * cy;
    w_ = cr * cp * cy + sr * sp * sy;
  }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double w() const { return w_; }

private:
  double x_, y_, z_, w_;
};
}  // namespace tf2
#endif
