// local_costmap_b200_node.cpp -- drop-in replacement of the reference executable
// local_costmap_node (catkin_ws/src/local_costmap/src/local_costmap_node.cpp) for a ROS box with
// a B200: same node name, topics, queue sizes and parameters as Prius::LocalCostmap::LocalCostmap
// (local_costmap.cpp:5-29); the grid is built by prius_b200::LocalCostmap -> libprius_planner_b200.so.
//
// Not linkable against roscpp in this repository's container (no ROS); it is compiled unchanged
// against the ROS test doubles and checked against the reference node in tests/test_gpu_ros_shim.py.
#include <nav_msgs/OccupancyGrid.h>
#include <ros/ros.h>
#include <sensor_msgs/LaserScan.h>
#include <sensor_msgs/PointCloud.h>
#include <tf/transform_listener.h>

#include "local_costmap_b200.hpp"

namespace {

pc_params read_params(ros::NodeHandle& pnh) {
  pc_params p;
  pc_default_params(&p);
  double res = p.res, h_m = 75, w_m = 75, gh_m = 3000, gw_m = 3000;
  pnh.getParam("local_costmap_res", res);                                   // LC:14
  pnh.getParam("local_costmap_height", h_m);                                // LC:15
  pnh.getParam("local_costmap_width", w_m);                                 // LC:16
  pnh.getParam("z_ground_buffer", p.z_ground_buffer);                       // LC:17
  pnh.getParam("obstacle_range", p.obstacle_range);                         // LC:18
  pnh.getParam("max_inflation_radius", p.max_inflation_radius);             // LC:19
  pnh.getParam("/global_costmap_node/global_costmap_height", gh_m);         // LC:23
  pnh.getParam("/global_costmap_node/global_costmap_width", gw_m);          // LC:24
  p.res = res;
  p.height_cells = h_m / res;                                               // LC:20-21
  p.width_cells = w_m / res;
  p.global_height_cells = gh_m / res;                                       // LC:25-26
  p.global_width_cells = gw_m / res;
  return p;
}

void mat12(const tf::Transform& t, double* m) {
  for (int r = 0; r < 3; ++r) {
    m[4 * r + 0] = t.getBasis()[r].x(); m[4 * r + 1] = t.getBasis()[r].y(); m[4 * r + 2] = t.getBasis()[r].z();
  }
  m[3] = t.getOrigin().x(); m[7] = t.getOrigin().y(); m[11] = t.getOrigin().z();
}

// the look-ups of mapPlanarScan (LC:89, 99, 108, 141), mapCenterScan (LC:164), mapRoadVectors (LC:217, 248)
bool read_tf(tf::TransformListener& l, prius_b200::CostmapTf* out) {
  const char* frames[2] = {"front_right_laser_link", "front_left_laser_link"};
  prius_b200::PlanarTf* scans[2] = {&out->right, &out->left};
  try {
    for (int k = 0; k < 2; ++k) {
      tf::StampedTransform t, tz;
      l.lookupTransform(frames[k], "/center_laser_link", ros::Time(0), t);
      l.lookupTransform("/map", frames[k], ros::Time(0), tz);
      double pitch, yaw;
      tf::Matrix3x3(t.getRotation()).getRPY(scans[k]->roll, pitch, yaw);
      scans[k]->tf_x = t.getOrigin().getX();
      scans[k]->tf_y = t.getOrigin().getY();
      scans[k]->z_map = tz.getOrigin().getZ();
      scans[k]->inv_x = t.inverse().getOrigin().getX();
      scans[k]->inv_y = t.inverse().getOrigin().getY();
    }
    tf::StampedTransform c, m;
    l.lookupTransform("/center_laser_link", "/map", ros::Time(0), c);
    out->center_z = c.getOrigin().getZ();
    l.lookupTransform("/map", "/base_link", ros::Time(0), m);
    mat12(m, out->map_from_base);
    mat12(m.inverse(), out->base_from_map);
  } catch (tf::TransformException& ex) {
    ROS_ERROR("%s", ex.what());
    return false;
  }
  return true;
}

prius_b200::LaserScan to_scan(const sensor_msgs::LaserScan& m) {
  prius_b200::LaserScan s;
  s.angle_min = m.angle_min; s.angle_increment = m.angle_increment; s.range_max = m.range_max;
  s.ranges.assign(m.ranges.begin(), m.ranges.end());
  return s;
}

}  // namespace

int main(int argc, char** argv) {
  ros::init(argc, argv, "local_costmap_node");
  ros::NodeHandle nh, pnh("~");
  const pc_params params = read_params(pnh);
  prius_b200::LocalCostmap costmap(params, 0);
  tf::TransformListener listener;
  ros::Publisher occ_grid_pub = nh.advertise<nav_msgs::OccupancyGrid>("/local_costmap", 100);   // LC:11

  // setupCostmap (LC:36-68): header and origin of the published grid
  nav_msgs::OccupancyGrid out;
  tf::StampedTransform laser;
  while (!listener.canTransform("/center_laser_link", "/base_link", ros::Time(0))) continue;
  try {
    listener.lookupTransform("/base_link", "/center_laser_link", ros::Time(0), laser);
  } catch (tf::TransformException& ex) {
    ROS_ERROR("%s", ex.what());
  }
  out.header.frame_id = "base_link";
  out.info.resolution = params.res;
  out.info.origin.position.x = -params.width_cells * params.res / 2 + laser.getOrigin().getX();
  out.info.origin.position.y = -params.height_cells * params.res / 2 + laser.getOrigin().getY();
  out.info.origin.position.z = 0;
  out.info.origin.orientation.w = laser.getRotation().getW();
  out.info.origin.orientation.x = laser.getRotation().getX();
  out.info.origin.orientation.y = laser.getRotation().getY();
  out.info.origin.orientation.z = laser.getRotation().getZ();
  out.info.height = params.height_cells;
  out.info.width = params.width_cells;

  costmap.on_error = [](const std::string& m) { ROS_ERROR_STREAM(m); };
  costmap.on_grid = [&](const prius_b200::CostmapGrid& g) {                                      // LC:360-364
    out.header.stamp = ros::Time::now();
    out.data.assign(g.data.begin(), g.data.end());
    occ_grid_pub.publish(out);
  };

  boost::function<void(const sensor_msgs::PointCloud::ConstPtr&)> center_cb =
      [&](const sensor_msgs::PointCloud::ConstPtr& msg) {                                       // LC:366-370
        prius_b200::CostmapTf t;
        if (!read_tf(listener, &t)) return;
        prius_b200::PointCloud pc;
        pc.xyz.reserve(3 * msg->points.size());
        for (const auto& p : msg->points) { pc.xyz.push_back(p.x); pc.xyz.push_back(p.y); pc.xyz.push_back(p.z); }
        costmap.centerLaserCallback(pc, t);
      };
  boost::function<void(const sensor_msgs::LaserScan::ConstPtr&)> right_cb =
      [&](const sensor_msgs::LaserScan::ConstPtr& msg) { costmap.rightLaserCallback(to_scan(*msg)); };   // LC:372-375
  boost::function<void(const sensor_msgs::LaserScan::ConstPtr&)> left_cb =
      [&](const sensor_msgs::LaserScan::ConstPtr& msg) { costmap.leftLaserCallback(to_scan(*msg)); };    // LC:377-380
  boost::function<void(const nav_msgs::OccupancyGrid::ConstPtr&)> global_cb =
      [&](const nav_msgs::OccupancyGrid::ConstPtr& msg) { costmap.globalCostmapCallback(msg->data); };  // LC:382-385
  ros::Subscriber s0 = nh.subscribe<sensor_msgs::PointCloud>("/prius/center_laser/scan", 100, center_cb);      // LC:7
  ros::Subscriber s1 = nh.subscribe<sensor_msgs::LaserScan>("/prius/front_right_laser/scan", 100, right_cb);   // LC:8
  ros::Subscriber s2 = nh.subscribe<sensor_msgs::LaserScan>("/prius/front_left_laser/scan", 100, left_cb);     // LC:9
  ros::Subscriber s3 = nh.subscribe<nav_msgs::OccupancyGrid>("/global_costmap", 1, global_cb);                 // LC:10
  ros::spin();
  return 0;
}
