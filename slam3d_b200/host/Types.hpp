// Types.hpp — the few slam3d core types the scan-matching path touches, re-declared without Eigen/Boost/PCL so the
// host mirror builds in this environment (none of those libraries is installed).
// Mirrors (names, argument meaning, error behaviour): slam3d/core/Types.hpp:48-56 (Transform, Covariance),
// :108-135 (Measurement), :145-187 (Constraint, SE3Constraint); slam3d/core/Sensor.hpp:44-72 (BadMeasurementType, NoMatch);
// slam3d/sensor/pcl/PointCloudSensor.hpp:43-100 (PointType, PointCloud, PointCloudMeasurement).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <exception>
#include <memory>
#include <string>
#include <vector>

namespace slam3d_b200 {

typedef double ScalarType;

// Eigen::Isometry3d stand-in: 4x4, column-major like Eigen's matrix().data().
struct Transform {
  std::array<double, 16> m;
  Transform() { m.fill(0.0); m[0] = m[5] = m[10] = m[15] = 1.0; }
  static Transform Identity() { return Transform(); }
  double& operator()(int r, int c) { return m[c * 4 + r]; }
  double operator()(int r, int c) const { return m[c * 4 + r]; }
  const double* data() const { return m.data(); }
  Transform operator*(const Transform& o) const {
    Transform r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += (*this)(i, k) * o(k, j); r(i, j) = s; }
    return r;
  }
  Transform inverse() const {  // rigid inverse, like Eigen::Isometry3d::inverse()
    Transform r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i);
    for (int i = 0; i < 3; ++i) { double s = 0; for (int j = 0; j < 3; ++j) s += r(i, j) * (*this)(j, 3); r(i, 3) = -s; }
    return r;
  }
  double translationNorm() const { return std::sqrt(m[12] * m[12] + m[13] * m[13] + m[14] * m[14]); }
};

template <unsigned N>
struct Covariance {
  std::array<double, N * N> m;
  Covariance() { m.fill(0.0); }
  static Covariance Identity() { Covariance c; for (unsigned i = 0; i < N; ++i) c.m[i * N + i] = 1.0; return c; }
  double operator()(unsigned r, unsigned c) const { return m[c * N + r]; }
  Covariance operator*(double s) const { Covariance c(*this); for (double& v : c.m) v *= s; return c; }
  Covariance inverseDiagonal() const { Covariance c; for (unsigned i = 0; i < N; ++i) c.m[i * N + i] = 1.0 / m[i * N + i]; return c; }
};

class BadMeasurementType : public std::exception {  // core/Sensor.hpp:44-53
 public:
  const char* what() const throw() override { return "Measurement type does not match sensor type!"; }
};

class NoMatch : public std::exception {  // core/Sensor.hpp:61-72
 public:
  explicit NoMatch(const std::string& msg) : message(msg) {}
  const char* what() const throw() override { return message.c_str(); }
  std::string message;
};

enum LOG_LEVEL { DEBUG, INFO, WARNING, ERROR, FATAL };
class Logger {  // core/Logger.hpp:47-107 (interface only)
 public:
  virtual ~Logger() {}
  virtual void message(LOG_LEVEL, const std::string&) {}
};

class Measurement {  // core/Types.hpp:108-135
 public:
  typedef std::shared_ptr<Measurement> Ptr;
  Measurement(const std::string& r, const std::string& s, const Transform& p) : mRobotName(r), mSensorName(s), mSensorPose(p), mInverseSensorPose(p.inverse()) {}
  virtual ~Measurement() {}
  std::string getRobotName() const { return mRobotName; }
  std::string getSensorName() const { return mSensorName; }
  Transform getSensorPose() const { return mSensorPose; }
  Transform getInverseSensorPose() const { return mInverseSensorPose; }
  virtual const char* getTypeName() const = 0;
 protected:
  std::string mRobotName, mSensorName;
  Transform mSensorPose, mInverseSensorPose;
};

enum ConstraintType { TENTATIVE, SE3, GRAVITY, POSITION, ORIENTATION, POSE };
class Constraint {  // core/Types.hpp:145-165
 public:
  typedef std::shared_ptr<Constraint> Ptr;
  explicit Constraint(const std::string& sensor) : mSensorName(sensor) {}
  virtual ~Constraint() {}
  virtual ConstraintType getType() = 0;
  const std::string& getSensorName() const { return mSensorName; }
 protected:
  std::string mSensorName;
};
class SE3Constraint : public Constraint {  // core/Types.hpp:168-187
 public:
  typedef std::shared_ptr<SE3Constraint> Ptr;
  SE3Constraint(const std::string& s, const Transform& t, const Covariance<6>& i) : Constraint(s), mRelativePose(t), mInformation(i) {}
  ConstraintType getType() override { return SE3; }
  const Transform& getRelativePose() const { return mRelativePose; }
  const Covariance<6>& getInformation() const { return mInformation; }
 protected:
  Transform mRelativePose;
  Covariance<6> mInformation;
};

struct alignas(16) PointType {  // pcl::PointXYZ memory: x, y, z, padding = 1.0f
  float x, y, z, pad;
  PointType() : x(0), y(0), z(0), pad(1.0f) {}
  PointType(float a, float b, float c) : x(a), y(b), z(c), pad(1.0f) {}
};
struct PointCloud {  // pcl::PointCloud<pcl::PointXYZ> stand-in
  typedef std::shared_ptr<PointCloud> Ptr;
  typedef std::shared_ptr<const PointCloud> ConstPtr;
  std::vector<PointType> points;
  size_t size() const { return points.size(); }
  void push_back(const PointType& p) { points.push_back(p); }
};

class PointCloudMeasurement : public Measurement {  // sensor/pcl/PointCloudSensor.hpp:50-100
 public:
  typedef std::shared_ptr<PointCloudMeasurement> Ptr;
  PointCloudMeasurement(const PointCloud::Ptr& cloud, const std::string& r, const std::string& s, const Transform& p) : Measurement(r, s, p), mPointCloud(cloud) {}
  const PointCloud::Ptr getPointCloud() const { return mPointCloud; }
  const char* getTypeName() const override { return "slam3d::PointCloudMeasurement"; }
 protected:
  PointCloud::Ptr mPointCloud;
};

}  // namespace slam3d_b200
