// local_planner_b200_node.cpp -- drop-in replacement of the reference executable
// local_planner_node (catkin_ws/src/local_planner/src/local_planner_node.cpp:1-21) for a
// ROS Kinetic box with a B200.  Same node name, topics, queue sizes and parameters as
// Prius::LocalPlanner::LocalPlanner (local_planner.cpp:5-18, 390-423); the work goes through
// prius_b200::LocalPlanner -> libprius_planner_b200.so.
//
// NOT BUILT IN THIS REPOSITORY'S CONTAINER (no ROS, tf, nav_msgs, gazebo_msgs, prius_msgs).
// It is a field-by-field translation between ROS messages and the structs of
// autonomouscar_b200/host/local_planner_b200.hpp; INTEGRATION.md has the CMake lines.
#include <gazebo_msgs/ModelStates.h>
#include <geometry_msgs/PoseStamped.h>
#include <nav_msgs/OccupancyGrid.h>
#include <nav_msgs/Path.h>
#include <prius_msgs/LocalNav.h>
#include <prius_msgs/MotionPlanning.h>
#include <ros/ros.h>
#include <tf/transform_listener.h>

#include "local_planner_b200.hpp"

namespace {

pp_params read_params(ros::NodeHandle& pnh) {
  pp_params p;
  pp_default_params(&p);
  double res = p.costmap_res, width_m = 75, height_m = 75;
  pnh.getParam("/local_costmap_node/local_costmap_res", res);        // LP:394
  pnh.getParam("/local_costmap_node/local_costmap_height", height_m);// LP:395
  pnh.getParam("/local_costmap_node/local_costmap_width", width_m);  // LP:396
  p.costmap_res = res;
  p.costmap_height = static_cast<int>(height_m / res);               // LP:397
  p.costmap_width = static_cast<int>(width_m / res);                 // LP:398
  pnh.getParam("collision_buffer_distance", p.collision_buffer_distance);  // LP:415-422
  pnh.getParam("time_step_ms", p.time_step_ms);
  pnh.getParam("velocity_res", p.velocity_res);
  pnh.getParam("heading_res", p.heading_res);
  pnh.getParam("heading_diff", p.heading_diff);
  pnh.getParam("goal_pos_tolerance", p.goal_pos_tolerance);
  pnh.getParam("goal_speed_tolerance", p.goal_speed_tolerance);
  pnh.getParam("search_res", p.search_res);
  // extensions of this build (absent = the reference's behaviour)
  pnh.param("collision_mode", p.collision_mode, static_cast<int>(PP_COLLISION_AS_SHIPPED));
  pnh.param("max_expansions", p.max_expansions, p.max_expansions);
  pnh.param("max_nodes", p.max_nodes, p.max_nodes);   // library default: 60000 ~ what the reference reaches in its 1 s budget
  pnh.param("car_length", p.car_length, 4.7);   // LP:331-334 derive these from the URDF through tf
  pnh.param("car_width", p.car_width, 1.586);
  return p;
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "local_planner_node");
  ros::NodeHandle nh, pnh("~");
  prius_b200::LocalPlanner planner(read_params(pnh), 0);
  tf::TransformListener listener;

  ros::Publisher mp_pub = nh.advertise<prius_msgs::MotionPlanning>("/mp", 10);       // LP:10
  ros::Publisher goal_pub = nh.advertise<geometry_msgs::PoseStamped>("/goal", 100);  // LP:11
  ros::Publisher path_pub = nh.advertise<nav_msgs::Path>("/local_path", 10);         // LP:12

  planner.on_error = [](const std::string& m) { ROS_ERROR_STREAM(m); };
  planner.on_goal = [&](double x, double y) {                                        // LP:633-641
    geometry_msgs::PoseStamped g;
    g.header.frame_id = "base_link";
    g.header.stamp = ros::Time::now();
    g.pose.position.x = x;
    g.pose.position.y = y;
    goal_pub.publish(g);
  };
  planner.on_path = [&](const prius_b200::Path& p) {                                 // LP:528-539
    nav_msgs::Path out;
    out.header.frame_id = "base_link";
    out.header.stamp = ros::Time::now();
    for (size_t k = 0; k < p.x.size(); ++k) {
      geometry_msgs::PoseStamped pose;
      pose.header = out.header;
      pose.pose.position.x = p.x[k];
      pose.pose.position.y = p.y[k];
      out.poses.push_back(pose);
    }
    path_pub.publish(out);
  };
  planner.on_mp = [&](const prius_b200::MotionPlanning& m) {                         // LP:544-582
    prius_msgs::MotionPlanning out;
    out.header.stamp = ros::Time::now();
    out.max_accel = m.max_accel;
    out.max_speed = m.max_speed;
    out.yaw_rates = m.yaw_rates;
    out.durations = m.durations;
    out.speeds = m.speeds;
    mp_pub.publish(out);
  };

  boost::function<void(const nav_msgs::OccupancyGrid::ConstPtr&)> costmap_cb =
      [&](const nav_msgs::OccupancyGrid::ConstPtr& msg) {                           // LP:658-661
        prius_b200::OccupancyGrid g;
        g.resolution = msg->info.resolution;
        g.width = msg->info.width;
        g.height = msg->info.height;
        g.data.assign(msg->data.begin(), msg->data.end());
        planner.costmapCallback(g);
      };
  boost::function<void(const gazebo_msgs::ModelStates::ConstPtr&)> states_cb =
      [&](const gazebo_msgs::ModelStates::ConstPtr& msg) {                          // LP:693-703
        for (size_t i = 0; i < msg->name.size(); ++i) {
          if (msg->name[i] == "prius") {
            prius_b200::Twist t;
            t.linear_x = msg->twist[i].linear.x;
            t.linear_y = msg->twist[i].linear.y;
            t.linear_z = msg->twist[i].linear.z;
            planner.gazeboStatesCallback(t);
            break;
          }
        }
      };
  boost::function<void(const prius_msgs::LocalNav::ConstPtr&)> nav_cb =
      [&](const prius_msgs::LocalNav::ConstPtr& msg) {                              // LP:663-691
        tf::StampedTransform tr;
        while (!listener.canTransform("/base_link", "/map", ros::Time(0))) continue;
        try {
          listener.lookupTransform("/base_link", "/map", ros::Time(0), tr);
        } catch (tf::TransformException& ex) {
          ROS_ERROR("%s", ex.what());
          return;
        }
        prius_b200::LocalNav nav;
        nav.x = msg->x; nav.y = msg->y; nav.speed = msg->speed;
        nav.max_speed = msg->max_speed; nav.max_accel = msg->max_accel;
        prius_b200::Transform2D t;
        t.tx = tr.getOrigin().x();
        t.ty = tr.getOrigin().y();
        t.yaw = tf::getYaw(tr.getRotation());
        planner.localNavCallback(nav, t);
      };
  ros::Subscriber costmap_sub = nh.subscribe<nav_msgs::OccupancyGrid>("/local_costmap", 10, costmap_cb);     // LP:7
  ros::Subscriber nav_sub = nh.subscribe<prius_msgs::LocalNav>("/local_nav_waypoints", 1, nav_cb);           // LP:8
  ros::Subscriber state_sub = nh.subscribe<gazebo_msgs::ModelStates>("/gazebo/model_states", 100, states_cb);// LP:9
  ros::spin();   // single-threaded, like local_planner_node.cpp:14-17
  return 0;
}
