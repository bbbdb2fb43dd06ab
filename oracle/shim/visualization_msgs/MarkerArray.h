// test-only ROS header stub, see ros/stub_msgs.h
#include <ros/stub_msgs.h>
