// Runs ros/local_costmap_b200_node.cpp UNCHANGED against the ROS test doubles on a GPU box.
// scenario lines:  param <name> <value> | tf <target> <source> <x> <y> <z> <roll> <pitch> <yaw>
//                  scan right|left <file> <n> <angle_min> <angle_inc> <range_max>
//                  global <file> <len> | cloud <file> <n_points> <out_file>
// After every `cloud` the published /local_costmap data is written to <out_file> (raw int8) and
// "published <count> <frame> <width> <height> <res %a> <origin x %a> <origin y %a>" is printed.
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#define main shim_main
#include "../../ros/local_costmap_b200_node.cpp"
#undef main

namespace {
struct Cmd { std::string op; std::vector<std::string> a; };
template <class T>
std::vector<T> read_raw(const std::string& path, size_t n) {
  std::vector<T> v(n);
  std::ifstream f(path, std::ios::binary);
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  return v;
}
void apply_tf(const Cmd& c) {
  tf::Quaternion q;
  q.setRPY(std::stod(c.a[5]), std::stod(c.a[6]), std::stod(c.a[7]));
  tf::register_transform(c.a[0], c.a[1], tf::Transform(q, tf::Vector3(std::stod(c.a[2]), std::stod(c.a[3]), std::stod(c.a[4]))));
}
}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1]);
  std::vector<Cmd> script;
  for (std::string line; std::getline(in, line);) {
    std::istringstream ls(line);
    Cmd c;
    if (!(ls >> c.op)) continue;
    for (std::string w; ls >> w;) c.a.push_back(w);
    script.push_back(c);
  }
  for (const Cmd& c : script) {
    if (c.op == "param") ros_doubles::state().params[c.a[0]] = std::stod(c.a[1]);
    if (c.op == "tf") apply_tf(c);
  }
  ros_doubles::state().quiet = false;
  ros_doubles::state().spin_script = [&]() {
    for (const Cmd& c : script) {
      if (c.op == "tf") {
        apply_tf(c);
      } else if (c.op == "scan") {
        sensor_msgs::LaserScan m;
        m.ranges = read_raw<float>(c.a[1], std::stoul(c.a[2]));
        m.angle_min = std::stof(c.a[3]); m.angle_increment = std::stof(c.a[4]); m.range_max = std::stof(c.a[5]);
        ros_doubles::deliver(c.a[0] == "right" ? "/prius/front_right_laser/scan" : "/prius/front_left_laser/scan", m);
      } else if (c.op == "global") {
        nav_msgs::OccupancyGrid g;
        g.data = read_raw<int8_t>(c.a[0], std::stoul(c.a[1]));
        ros_doubles::deliver("/global_costmap", g);
      } else if (c.op == "cloud") {
        sensor_msgs::PointCloud pc;
        const std::vector<float> xyz = read_raw<float>(c.a[0], 3 * std::stoul(c.a[1]));
        pc.points.resize(xyz.size() / 3);
        for (size_t k = 0; k < pc.points.size(); ++k) { pc.points[k].x = xyz[3 * k]; pc.points[k].y = xyz[3 * k + 1]; pc.points[k].z = xyz[3 * k + 2]; }
        auto& pub = ros_doubles::captured<nav_msgs::OccupancyGrid>();
        pub.clear();
        ros_doubles::deliver("/prius/center_laser/scan", pc);
        if (pub.empty()) { std::printf("published 0\n"); continue; }
        const auto& g = pub.back();
        std::ofstream o(c.a[2], std::ios::binary);
        o.write(reinterpret_cast<const char*>(g.data.data()), g.data.size());
        std::printf("published %zu %s %u %u %a %a %a\n", pub.size(), g.header.frame_id.c_str(), g.info.width, g.info.height,
                    static_cast<double>(g.info.resolution), g.info.origin.position.x, g.info.origin.position.y);
      }
    }
  };
  char* args[] = {argv[0], nullptr};
  return shim_main(1, args);
}
