// TEST INFRASTRUCTURE -- header-only TEST DOUBLES of the handful of ROS types the reference's
// nodes touch (roscpp, tf, message structs).  Written for this repo; nothing here is ROS
// source.  Purpose: let oracle/Makefile compile the reference's own .cpp files UNMODIFIED,
// from where they lie under /root/reference, into oracle/_ref/ so that the CPU oracle (and
// through it the CUDA path) is pinned against the reference's real code instead of only a
// restatement of it.  The doubles carry data and nothing else:
//   * messages are plain structs with the fields of the .msg definitions;
//   * NodeHandle::getParam reads a thread-local name -> double table the harness fills;
//   * Publisher::publish appends a copy of the message to a thread-local capture list;
//   * ros::Time::now() is a deterministic tick counter (the harness turns the reference's
//     1 s wall-clock guard, local_planner.cpp:33-39, into an expansion / node cap with it);
//   * tf lookups return transforms the harness registered.
// All arithmetic of the path stays in the reference's own translation unit + libm + libstdc++.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace ros_doubles {

struct State {
  std::map<std::string, double> params;
  // clock: now() returns `ticks`, then advances it by one; toSec() = ticks * sec_per_tick.
  int64_t ticks = 0;
  double ticks_per_sec = 1e18;      // "never times out" until the harness says otherwise
  std::function<bool()> force_timeout;  // optional: when it returns true, now() jumps far ahead
  // manual clock (controller harness): now() returns `manual_sec` and does not tick
  bool manual_clock = false;
  double manual_sec = 0;
  bool quiet = true;
  std::function<void()> spin_script;    // what ros::spin() runs (the harness delivers messages from it)
};
inline State& state() { thread_local State s; return s; }

template <class T>
inline std::vector<T>& captured() { thread_local std::vector<T> v; return v; }

// topic -> callbacks registered through NodeHandle::subscribe<M>
template <class M>
inline std::map<std::string, std::vector<std::function<void(const std::shared_ptr<const M>&)>>>& subscribers() {
  thread_local std::map<std::string, std::vector<std::function<void(const std::shared_ptr<const M>&)>>> m;
  return m;
}
// what a publisher on `topic` would cause: every subscriber's callback runs, in order
template <class M>
inline int deliver(const std::string& topic, const M& msg) {
  auto it = subscribers<M>().find(topic);
  if (it == subscribers<M>().end()) return 0;
  auto ptr = std::make_shared<const M>(msg);
  for (auto& cb : it->second) cb(ptr);
  return static_cast<int>(it->second.size());
}

}  // namespace ros_doubles

namespace boost {
template <class S>
using function = std::function<S>;
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
// boost::lexical_cast<T>(string): the whole string must parse (stream semantics), else it throws
template <class T>
inline T lexical_cast(const std::string& s) {
  std::istringstream in(s);
  T v;
  if (!(in >> v) || !in.eof()) throw bad_lexical_cast();
  return v;
}
}

namespace ros {

struct Duration {
  double sec = 0;
  Duration() {}
  explicit Duration(double s) : sec(s) {}
  double toSec() const { return sec; }
  bool sleep() const { return true; }
};

struct Time {
  int64_t ticks = 0;
  bool forced = false;
  bool manual = false;
  double sec = 0;                   // manual clock only
  Time() {}
  explicit Time(double) {}
  static Time now() {
    ros_doubles::State& s = ros_doubles::state();
    Time t;
    if (s.manual_clock) {
      t.manual = true;
      t.sec = s.manual_sec;
      return t;
    }
    t.ticks = s.ticks++;
    t.forced = s.force_timeout && s.force_timeout();
    return t;
  }
  double toSec() const { return manual ? sec : static_cast<double>(ticks) / ros_doubles::state().ticks_per_sec; }
};
inline Duration operator-(const Time& a, const Time& b) {
  if (a.manual || b.manual) return Duration(a.sec - b.sec);
  if (a.forced) return Duration(1e9);
  return Duration(static_cast<double>(a.ticks - b.ticks) / ros_doubles::state().ticks_per_sec);
}

struct Subscriber {};
struct Publisher {
  template <class M>
  void publish(const M& m) const { ros_doubles::captured<M>().push_back(m); }
};

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class M, class T>
  Subscriber subscribe(const std::string& topic, uint32_t, void (T::*fp)(const std::shared_ptr<const M>&), T* obj) {
    ros_doubles::subscribers<M>()[topic].push_back([obj, fp](const std::shared_ptr<const M>& m) { (obj->*fp)(m); });
    return Subscriber();
  }
  template <class M>
  Subscriber subscribe(const std::string& topic, uint32_t, const std::function<void(const std::shared_ptr<const M>&)>& cb) {
    ros_doubles::subscribers<M>()[topic].push_back(cb);
    return Subscriber();
  }
  template <class M>
  Publisher advertise(const std::string&, uint32_t, bool = false) { return Publisher(); }
  // names are looked up by their last path component as well, so "~name" and "/node/name" agree
  bool getParam(const std::string& name, double& out) const {
    const auto& p = ros_doubles::state().params;
    auto it = p.find(name);
    if (it == p.end()) {
      const size_t k = name.rfind('/');
      if (k != std::string::npos) it = p.find(name.substr(k + 1));
    }
    if (it == p.end()) return false;
    out = it->second;
    return true;
  }
  bool getParam(const std::string& name, int& out) const {
    double d;
    if (!getParam(name, d)) return false;
    out = static_cast<int>(d);
    return true;
  }
  template <class T>
  bool param(const std::string& name, T& out, const T& fallback) const {
    if (getParam(name, out)) return true;
    out = fallback;
    return false;
  }
  bool getParam(const std::string& name, std::string& out) const {
    (void)name; out.clear(); return false;
  }
};

inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void spin() { if (ros_doubles::state().spin_script) ros_doubles::state().spin_script(); }
inline void spinOnce() {}

}  // namespace ros

#define ROS_DOUBLES_LOG(tag, expr)                                   \
  do {                                                               \
    if (!ros_doubles::state().quiet) {                               \
      std::ostringstream ros_doubles_ss;                             \
      ros_doubles_ss << expr;                                        \
      std::fprintf(stderr, "[%s] %s\n", tag, ros_doubles_ss.str().c_str()); \
    }                                                                \
  } while (0)
#define ROS_INFO_STREAM(expr) ROS_DOUBLES_LOG("INFO", expr)
#define ROS_WARN_STREAM(expr) ROS_DOUBLES_LOG("WARN", expr)
#define ROS_ERROR_STREAM(expr) ROS_DOUBLES_LOG("ERROR", expr)
#define ROS_INFO(...) do { if (!ros_doubles::state().quiet) { std::fprintf(stderr, __VA_ARGS__); std::fputc('\n', stderr); } } while (0)
#define ROS_WARN(...) ROS_INFO(__VA_ARGS__)
#define ROS_ERROR(...) ROS_INFO(__VA_ARGS__)

// ---- message structs (fields of the .msg files; float64 -> double, int8[] -> vector<int8_t>)
namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Point32 { float x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 0; };
struct Pose { Point position; Quaternion orientation; };
struct Twist { Vector3 linear, angular; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}
namespace nav_msgs {
struct MapMetaData { ros::Time map_load_time; float resolution = 0; uint32_t width = 0, height = 0; geometry_msgs::Pose origin; };
struct OccupancyGrid {
  std_msgs::Header header; MapMetaData info; std::vector<int8_t> data;
  typedef std::shared_ptr<const OccupancyGrid> ConstPtr;
};
struct Path {
  std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses;
  typedef std::shared_ptr<const Path> ConstPtr;
};
}
namespace gazebo_msgs {
struct ModelStates {
  std::vector<std::string> name; std::vector<geometry_msgs::Pose> pose; std::vector<geometry_msgs::Twist> twist;
  typedef std::shared_ptr<const ModelStates> ConstPtr;
};
struct ModelState { std::string model_name; geometry_msgs::Pose pose; geometry_msgs::Twist twist; std::string reference_frame; };
}
namespace prius_msgs {
struct LocalNav {   // prius_msgs/msg/LocalNav.msg
  std_msgs::Header header; double x = 0, y = 0, speed = 0, max_speed = 0, max_accel = 0;
  typedef std::shared_ptr<const LocalNav> ConstPtr;
};
struct Control {   // prius_msgs/msg/Control.msg
  enum { NO_COMMAND = 0, NEUTRAL = 1, FORWARD = 2, REVERSE = 3 };
  std_msgs::Header header; double throttle = 0, brake = 0, steer = 0; uint8_t shift_gears = 0;
  typedef std::shared_ptr<const Control> ConstPtr;
};
struct MotionPlanning {   // prius_msgs/msg/MotionPlanning.msg
  std_msgs::Header header; std::vector<double> yaw_rates, durations, speeds; double max_speed = 0, max_accel = 0;
  typedef std::shared_ptr<const MotionPlanning> ConstPtr;
};
}
namespace sensor_msgs {
struct LaserScan {
  std_msgs::Header header;
  float angle_min = 0, angle_max = 0, angle_increment = 0, time_increment = 0, scan_time = 0, range_min = 0, range_max = 0;
  std::vector<float> ranges, intensities;
  typedef std::shared_ptr<const LaserScan> ConstPtr;
};
struct ChannelFloat32 { std::string name; std::vector<float> values; };
struct PointCloud {
  std_msgs::Header header; std::vector<geometry_msgs::Point32> points; std::vector<ChannelFloat32> channels;
  typedef std::shared_ptr<const PointCloud> ConstPtr;
};
struct Range { std_msgs::Header header; float range = 0; typedef std::shared_ptr<const Range> ConstPtr; };
}

// ---- dynamic_reconfigure + the generated config of prius_controller (cfg/prius_controller.cfg: name, default)
namespace prius_controller {
struct PriusControllerConfig { double kp_s = 7.5, ki_s = 0.05, kd_s = 0.15, kp_p = 0.25, ki_p = 0.075, kd_p = 0.055; };
}
namespace dynamic_reconfigure {
template <class Config>
class Server {
 public:
  typedef boost::function<void(Config&, uint32_t)> CallbackType;
  // like the real server, setCallback calls back at once with the current (default) configuration
  void setCallback(const CallbackType& cb) { cb_ = cb; if (cb_) cb_(config, ~uint32_t(0)); }
  void update(const Config& c) { config = c; if (cb_) cb_(config, 0); }
  Config config;
 private:
  CallbackType cb_;
};
}
// boost::bind(&Class::method, obj, _1, _2) -- the one shape the controller uses
namespace boost {
namespace doubles_placeholders { struct P1 {}; struct P2 {}; }
template <class C, class A, class B>
inline std::function<void(A, B)> bind(void (C::*m)(A, B), C* obj, doubles_placeholders::P1, doubles_placeholders::P2) {
  return [obj, m](A a, B b) { (obj->*m)(a, b); };
}
}
namespace { const boost::doubles_placeholders::P1 _1{}; const boost::doubles_placeholders::P2 _2{}; }
