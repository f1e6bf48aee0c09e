// TEST INFRASTRUCTURE -- test double of the slice of tf the reference's nodes use
// (see ../ros_doubles.h).  The linear algebra is the textbook rigid-transform arithmetic in
// double precision, written here from the formulas; lookups return what the harness registered.
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>

#include "../ros_doubles.h"

namespace tf {

typedef double tfScalar;

class Vector3 {
 public:
  Vector3() : v_{0, 0, 0, 0} {}
  Vector3(double x, double y, double z) : v_{x, y, z, 0} {}
  double getX() const { return v_[0]; }
  double getY() const { return v_[1]; }
  double getZ() const { return v_[2]; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  double w() const { return v_[3]; }
  double dot(const Vector3& o) const { return v_[0] * o.v_[0] + v_[1] * o.v_[1] + v_[2] * o.v_[2]; }
  Vector3 operator-() const { return Vector3(-v_[0], -v_[1], -v_[2]); }
 private:
  double v_[4];
};
typedef Vector3 Point;

class Quaternion {
 public:
  Quaternion() : q_{0, 0, 0, 1} {}
  Quaternion(double x, double y, double z, double w) : q_{x, y, z, w} {}
  double getX() const { return q_[0]; }
  double getY() const { return q_[1]; }
  double getZ() const { return q_[2]; }
  double getW() const { return q_[3]; }
  double x() const { return q_[0]; }
  double y() const { return q_[1]; }
  double z() const { return q_[2]; }
  double w() const { return q_[3]; }
  void setX(double v) { q_[0] = v; }
  void setY(double v) { q_[1] = v; }
  void setZ(double v) { q_[2] = v; }
  void setW(double v) { q_[3] = v; }
  double length2() const { return q_[0] * q_[0] + q_[1] * q_[1] + q_[2] * q_[2] + q_[3] * q_[3]; }
  // fixed-axis roll (x), pitch (y), yaw (z)
  void setRPY(double roll, double pitch, double yaw) {
    const double cy = std::cos(yaw / 2), sy = std::sin(yaw / 2);
    const double cp = std::cos(pitch / 2), sp = std::sin(pitch / 2);
    const double cr = std::cos(roll / 2), sr = std::sin(roll / 2);
    q_[0] = sr * cp * cy - cr * sp * sy;
    q_[1] = cr * sp * cy + sr * cp * sy;
    q_[2] = cr * cp * sy - sr * sp * cy;
    q_[3] = cr * cp * cy + sr * sp * sy;
  }
 private:
  double q_[4];
};

class Matrix3x3 {
 public:
  Matrix3x3() { setIdentity(); }
  explicit Matrix3x3(const Quaternion& q) { setRotation(q); }
  Matrix3x3(double xx, double xy, double xz, double yx, double yy, double yz, double zx, double zy, double zz) {
    r_[0] = Vector3(xx, xy, xz); r_[1] = Vector3(yx, yy, yz); r_[2] = Vector3(zx, zy, zz);
  }
  void setIdentity() { r_[0] = Vector3(1, 0, 0); r_[1] = Vector3(0, 1, 0); r_[2] = Vector3(0, 0, 1); }
  void setRotation(const Quaternion& q) {
    const double d = q.length2();
    const double s = 2.0 / d;
    const double xs = q.x() * s, ys = q.y() * s, zs = q.z() * s;
    const double wx = q.w() * xs, wy = q.w() * ys, wz = q.w() * zs;
    const double xx = q.x() * xs, xy = q.x() * ys, xz = q.x() * zs;
    const double yy = q.y() * ys, yz = q.y() * zs, zz = q.z() * zs;
    r_[0] = Vector3(1.0 - (yy + zz), xy - wz, xz + wy);
    r_[1] = Vector3(xy + wz, 1.0 - (xx + zz), yz - wx);
    r_[2] = Vector3(xz - wy, yz + wx, 1.0 - (xx + yy));
  }
  const Vector3& operator[](int i) const { return r_[i]; }
  Matrix3x3 transpose() const {
    return Matrix3x3(r_[0].x(), r_[1].x(), r_[2].x(), r_[0].y(), r_[1].y(), r_[2].y(), r_[0].z(), r_[1].z(), r_[2].z());
  }
  Vector3 operator*(const Vector3& v) const { return Vector3(r_[0].dot(v), r_[1].dot(v), r_[2].dot(v)); }
  void getRPY(double& roll, double& pitch, double& yaw) const {
    if (std::fabs(r_[2].x()) >= 1) {
      yaw = 0;
      const double delta = std::atan2(r_[2].y(), r_[2].z());
      if (r_[2].x() < 0) { pitch = M_PI / 2.0; roll = delta; }
      else { pitch = -M_PI / 2.0; roll = delta; }
    } else {
      pitch = -std::asin(r_[2].x());
      roll = std::atan2(r_[2].y() / std::cos(pitch), r_[2].z() / std::cos(pitch));
      yaw = std::atan2(r_[1].x() / std::cos(pitch), r_[0].x() / std::cos(pitch));
    }
  }
  Quaternion getRotation() const {
    const double trace = r_[0].x() + r_[1].y() + r_[2].z();
    double t[4];
    if (trace > 0.0) {
      double s = std::sqrt(trace + 1.0);
      t[3] = s * 0.5;
      s = 0.5 / s;
      t[0] = (r_[2].y() - r_[1].z()) * s;
      t[1] = (r_[0].z() - r_[2].x()) * s;
      t[2] = (r_[1].x() - r_[0].y()) * s;
    } else {
      const double m[3][3] = {{r_[0].x(), r_[0].y(), r_[0].z()}, {r_[1].x(), r_[1].y(), r_[1].z()}, {r_[2].x(), r_[2].y(), r_[2].z()}};
      const int i = m[0][0] < m[1][1] ? (m[1][1] < m[2][2] ? 2 : 1) : (m[0][0] < m[2][2] ? 2 : 0);
      const int j = (i + 1) % 3, k = (i + 2) % 3;
      double s = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
      t[i] = s * 0.5;
      s = 0.5 / s;
      t[3] = (m[k][j] - m[j][k]) * s;
      t[j] = (m[j][i] + m[i][j]) * s;
      t[k] = (m[k][i] + m[i][k]) * s;
    }
    return Quaternion(t[0], t[1], t[2], t[3]);
  }
 private:
  Vector3 r_[3];
};

class Transform {
 public:
  Transform() {}
  Transform(const Matrix3x3& b, const Vector3& o) : basis_(b), origin_(o) {}
  Transform(const Quaternion& q, const Vector3& o) : basis_(q), origin_(o) {}
  const Vector3& getOrigin() const { return origin_; }
  const Matrix3x3& getBasis() const { return basis_; }
  Quaternion getRotation() const { return basis_.getRotation(); }
  void setOrigin(const Vector3& o) { origin_ = o; }
  void setRotation(const Quaternion& q) { basis_.setRotation(q); }
  Vector3 operator*(const Vector3& x) const {
    return Vector3(basis_[0].dot(x) + origin_.x(), basis_[1].dot(x) + origin_.y(), basis_[2].dot(x) + origin_.z());
  }
  Transform inverse() const {
    const Matrix3x3 inv = basis_.transpose();
    return Transform(inv, inv * -origin_);
  }
 private:
  Matrix3x3 basis_;
  Vector3 origin_;
};

class StampedTransform : public Transform {
 public:
  StampedTransform() {}
  StampedTransform(const Transform& t) : Transform(t) {}
  ros::Time stamp_;
  std::string frame_id_, child_frame_id_;
};

class TransformException : public std::runtime_error {
 public:
  explicit TransformException(const std::string& w) : std::runtime_error(w) {}
};

inline std::string strip_slash(const std::string& f) { return (!f.empty() && f[0] == '/') ? f.substr(1) : f; }
// harness-side registry: (target, source) -> transform
inline std::map<std::string, Transform>& registry() { thread_local std::map<std::string, Transform> r; return r; }
inline void register_transform(const std::string& target, const std::string& source, const Transform& t) {
  registry()[strip_slash(target) + "<-" + strip_slash(source)] = t;
}

// yaw of a rotation about z (what tf::getYaw returns for planar motion)
inline double getYaw(const Quaternion& q) {
  double roll, pitch, yaw;
  Matrix3x3(q).getRPY(roll, pitch, yaw);
  return yaw;
}

class TransformListener {
 public:
  bool canTransform(const std::string&, const std::string&, const ros::Time&) const { return true; }
  void lookupTransform(const std::string& target, const std::string& source, const ros::Time&, StampedTransform& out) const {
    auto it = registry().find(strip_slash(target) + "<-" + strip_slash(source));
    if (it == registry().end()) throw TransformException("no transform " + target + " <- " + source);
    out = StampedTransform(it->second);
  }
};

}  // namespace tf
