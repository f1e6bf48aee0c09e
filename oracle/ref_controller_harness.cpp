// TEST INFRASTRUCTURE -- C harness around the REFERENCE's own controller,
//   /root/reference/catkin_ws/src/prius_controller/src/controller.cpp   (compiled UNMODIFIED, included below)
// against the ROS test doubles of oracle/ros_doubles/.  Built by oracle/Makefile into
// oracle/_ref/libref_controller.so; used by tests/ to pin the profile follower of row N4
// (autonomouscar_b200/rollout.py: ProfileFollower / PriusController).  Never linked into the product.
//
// The node's main (controller_node.cpp:14-31) is: spin until the first /mp arrived, then
// calculateAndPublishControls() at 100 Hz.  The harness exposes those steps one by one with a manual clock.
// The reference leaves m_speed / m_yaw_rate uninitialised (controller.hpp:55-56); the object is built in
// zeroed storage so that calculateGear's first read (controller.cpp:67) is defined.  Stepping past the end of
// the profile indexes the vectors out of bounds (:148, :47): refk_step refuses and reports -1.
#include <cstdlib>
#include <cstring>
#include <new>

#include "ros_doubles/ros_doubles.h"   // (and every std header it pulls in) before the access switch

#define private public
#include REF_CTRL_CPP
#undef private

namespace {
struct Node {
  ros::NodeHandle nh, pnh{"~"};
  Prius::PriusController* ctrl = nullptr;
  void* mem = nullptr;
};
}

extern "C" {

void refk_set_time(double t) {
  ros_doubles::state().manual_clock = true;
  ros_doubles::state().manual_sec = t;
}

void* refk_create(double t0) {
  refk_set_time(t0);
  ros_doubles::subscribers<prius_msgs::MotionPlanning>().clear();
  ros_doubles::subscribers<gazebo_msgs::ModelStates>().clear();
  ros_doubles::captured<prius_msgs::Control>().clear();
  Node* n = new Node();
  n->mem = std::calloc(1, sizeof(Prius::PriusController));
  n->ctrl = new (n->mem) Prius::PriusController(n->nh, n->pnh);
  return n;
}

void refk_destroy(void* h) {
  Node* n = static_cast<Node*>(h);
  if (!n) return;
  n->ctrl->~PriusController();
  std::free(n->mem);
  delete n;
  ros_doubles::state().manual_clock = false;
}

// /gazebo/model_states with one model called "prius" (and a decoy in front of it)
void refk_model_state(void* h, double speed_x, double yaw_rate_z) {
  (void)h;
  gazebo_msgs::ModelStates m;
  m.name = {"ground_plane", "prius"};
  m.pose.resize(2);
  m.twist.resize(2);
  m.twist[0].linear.x = -99;
  m.twist[1].linear.x = speed_x;
  m.twist[1].angular.z = yaw_rate_z;
  ros_doubles::deliver("/gazebo/model_states", m);
}

// /mp: motionPlanningCallback (controller.cpp:176-183) -- runs one control pass itself
void refk_profile(void* h, int n, const double* yaw_rates, const double* durations, const double* speeds,
                  double max_speed, double max_accel) {
  (void)h;
  prius_msgs::MotionPlanning m;
  m.yaw_rates.assign(yaw_rates, yaw_rates + n);
  m.durations.assign(durations, durations + n);
  m.speeds.assign(speeds, speeds + n);
  m.max_speed = max_speed;
  m.max_accel = max_accel;
  ros_doubles::deliver("/mp", m);
}

// one pass of the 100 Hz loop; -1 when the reference would index past the end of the profile
int refk_step(void* h) {
  Node* n = static_cast<Node*>(h);
  if (!n->ctrl->m_received_msg) return -2;
  const int size = static_cast<int>(n->ctrl->m_mp_control.yaw_rates.size()), it = n->ctrl->m_control_it;
  if (it >= size) return -1;
  // the pass that moves the cursor onto size already reads yaw_rates[size] (:144-153 then :47)
  if (it == size - 1 && (ros::Time::now() - n->ctrl->m_start_time).toSec() >= n->ctrl->m_mp_control.durations[it] / 1000) return -1;
  n->ctrl->calculateAndPublishControls();
  return 0;
}

// out[0..5] = m_control_it, throttle, brake, steer, shift_gears, messages published so far
void refk_state(void* h, double* out) {
  Node* n = static_cast<Node*>(h);
  out[0] = n->ctrl->m_control_it;
  out[1] = n->ctrl->m_control.throttle;
  out[2] = n->ctrl->m_control.brake;
  out[3] = n->ctrl->m_control.steer;
  out[4] = n->ctrl->m_control.shift_gears;
  out[5] = static_cast<double>(ros_doubles::captured<prius_msgs::Control>().size());
}

}  // extern "C"
