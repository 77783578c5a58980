// MiniHost.hpp — a dependency-free stand-in for the CALLER of the hot path: the link-to-previous policy of
// slam3d::ScanSensor::addMeasurement(m, odom) (slam3d/core/ScanSensor.cpp:94-135) and the loop-closure call of
// ScanSensor::link (:137-168), with a recording "graph" instead of BoostGraph/g2o (not installed here).
// It exists so tests can drive createConstraint in the reference's call order, with its exception handling
// (NoMatch -> warning, vertex stays unlinked) and from two threads at once (ScanSensor.cpp:209-210).
#pragma once

#include <mutex>
#include <string>
#include <vector>

#include "PointCloudSensor.hpp"

namespace slam3d_b200 {

struct RecordedEdge { unsigned source, target; Transform relative; Covariance<6> information; bool loop; };

class MiniHost {
 public:
  explicit MiniHost(PointCloudSensor* s) : mSensor(s), mLastVertex(0), mHasVertex(false), mLinkPrevious(true) {}

  // ScanSensor::addMeasurement(m, odom)  :94-135 (checkMinDistance omitted: every scan becomes a vertex)
  bool addMeasurement(const Measurement::Ptr& m, const Transform& odom) {
    if (!mHasVertex) {  // :96-101
      mMeasurements.push_back(m); mLastVertex = 0; mHasVertex = true; mLastOdometry = odom;
      return true;
    }
    Transform lastTransform = mLastOdometry.inverse() * odom;  // :104
    const unsigned newVertex = static_cast<unsigned>(mMeasurements.size());
    Measurement::Ptr source = mMeasurements[mLastVertex];
    mMeasurements.push_back(m);  // :107
    if (mLinkPrevious) {
      try {
        Constraint::Ptr c = mSensor->createConstraint(source, m, lastTransform, false);  // :113
        record(mLastVertex, newVertex, c, false);                                        // :114
      } catch (std::exception& e) {  // :124-127
        std::lock_guard<std::mutex> g(mMutex);
        warnings.push_back(std::string("Could not link Measurement to previous: ") + e.what());
      }
    }
    mLastOdometry = odom; mLastVertex = newVertex;  // :129-131
    return true;
  }

  // ScanSensor::link(source_id, target_id, guess)  :144-168 (patch range 0: the patch is the measurement itself)
  void link(unsigned source_id, unsigned target_id, const Transform& guess) {
    try {
      Constraint::Ptr se3 = mSensor->createConstraint(mMeasurements[source_id], mMeasurements[target_id], guess, true);  // :156
      record(source_id, target_id, se3, true);
    } catch (NoMatch& e) {  // :159-166
      std::lock_guard<std::mutex> g(mMutex);
      warnings.push_back(std::string("Failed to link vertex: ") + e.what());
    }
  }

  std::vector<RecordedEdge> edges;
  std::vector<std::string> warnings;

 private:
  void record(unsigned s, unsigned t, const Constraint::Ptr& c, bool loop) {
    SE3Constraint::Ptr se3 = std::dynamic_pointer_cast<SE3Constraint>(c);
    std::lock_guard<std::mutex> g(mMutex);
    edges.push_back({s, t, se3->getRelativePose(), se3->getInformation(), loop});
  }
  PointCloudSensor* mSensor;
  std::vector<Measurement::Ptr> mMeasurements;
  unsigned mLastVertex;
  bool mHasVertex, mLinkPrevious;
  Transform mLastOdometry;
  std::mutex mMutex;
};

}  // namespace slam3d_b200
