// MiniHost.hpp — a dependency-free stand-in for the CALLER of the hot path: the link-to-previous policy of
// slam3d::ScanSensor::addMeasurement(m, odom) (slam3d/core/ScanSensor.cpp:94-135) and the loop-closure call of
// ScanSensor::link (:137-168), with a recording "graph" instead of BoostGraph/g2o (not installed here).
// It exists so tests can drive createConstraint in the reference's call order, with its exception handling
// (NoMatch -> warning, vertex stays unlinked) and from two threads at once (ScanSensor.cpp:209-210).
// linkLastToNeighbors() reproduces ScanSensor::linkToNeighbors (:170-202) on the recorded graph: vertices within
// mNeighborRadius of the new vertex's corrected pose (Graph::getNearbyVertices, core/Graph.cpp:240-261), newest first, no
// existing edge, hop distance (BoostGraph::calculateGraphDistance, unit edge weights) >= mMinLoopLength, at most
// mMaxNeighorLinks links, each through link(index, vertex) with the graph's current relative pose as the guess (:137-142).
#pragma once

#include <cmath>
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
      mCorrected.push_back(Transform::Identity()); mAdjacency.emplace_back();  // Mapper::addMeasurement: vertex at mStartPose
      return true;
    }
    Transform lastTransform = mLastOdometry.inverse() * odom;  // :104
    const unsigned newVertex = static_cast<unsigned>(mMeasurements.size());
    Measurement::Ptr source = mMeasurements[mLastVertex];
    mMeasurements.push_back(m);  // :107
    mCorrected.push_back(mCorrected[mLastVertex]); mAdjacency.emplace_back();  // Mapper.cpp:89: new vertex at the mapper's current pose
    if (mLinkPrevious) {
      try {
        Constraint::Ptr c = mSensor->createConstraint(source, m, lastTransform, false);  // :113
        record(mLastVertex, newVertex, c, false);                                        // :114
        SE3Constraint::Ptr se3 = std::dynamic_pointer_cast<SE3Constraint>(c);            // :117-123
        if (se3) mCorrected[newVertex] = mCorrected[mLastVertex] * se3->getRelativePose();
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

  // ScanSensor::linkLastToNeighbors(false) -> linkToNeighbors(mLastVertex)  :170-213 (patch building range 0)
  void linkLastToNeighbors() {
    if (mMaxNeighorLinks < 1 || !mHasVertex) return;
    const unsigned vertex = mLastVertex;
    std::vector<unsigned> neighbors;  // Graph::getNearbyVertices: index order, d < radius
    for (unsigned v = 0; v < mCorrected.size(); ++v) {
      double d2 = 0;
      for (int a = 0; a < 3; ++a) { const double d = mCorrected[v](a, 3) - mCorrected[vertex](a, 3); d2 += d * d; }
      if (std::sqrt(d2) < mNeighborRadius) neighbors.push_back(v);
    }
    int count = 0;
    for (auto i = neighbors.rbegin(); i != neighbors.rend() && count < mMaxNeighorLinks; ++i) {
      const unsigned index = *i;
      if (index == vertex) continue;
      bool has_edge = false;
      for (unsigned o : mAdjacency[vertex]) if (o == index) has_edge = true;
      if (has_edge) continue;
      const float dist = graphDistance(index, vertex);
      if (dist <= 0.f /* mPatchBuildingRange * 2 */ || dist < (float)mMinLoopLength) continue;
      ++count;
      link(index, vertex, mCorrected[index].inverse() * mCorrected[vertex]);  // :137-142: guess = Graph::getTransform(source, target)
    }
  }

  void setNeighborRadius(float r, int max_links) { mNeighborRadius = r; mMaxNeighorLinks = max_links; }  // ScanSensor.hpp setters
  void setMinLoopLength(unsigned l) { mMinLoopLength = l; }
  const Transform& correctedPose(unsigned v) const { return mCorrected[v]; }

  std::vector<RecordedEdge> edges;
  std::vector<std::string> warnings;

 private:
  void record(unsigned s, unsigned t, const Constraint::Ptr& c, bool loop) {
    SE3Constraint::Ptr se3 = std::dynamic_pointer_cast<SE3Constraint>(c);
    std::lock_guard<std::mutex> g(mMutex);
    edges.push_back({s, t, se3->getRelativePose(), se3->getInformation(), loop});
    mAdjacency[s].push_back(t); mAdjacency[t].push_back(s);
  }
  // hop count between two vertices (dijkstra with unit weights == BFS); unreachable = max float like boost's distance map
  float graphDistance(unsigned from, unsigned to) const {
    std::vector<int> dist(mAdjacency.size(), -1);
    std::vector<unsigned> queue{from};
    dist[from] = 0;
    for (size_t h = 0; h < queue.size(); ++h) {
      const unsigned v = queue[h];
      if (v == to) return (float)dist[v];
      for (unsigned o : mAdjacency[v]) if (dist[o] < 0) { dist[o] = dist[v] + 1; queue.push_back(o); }
    }
    return 3.4028235e38f;
  }
  std::vector<Transform> mCorrected;                // VertexObject::correctedPose (no optimiser here: poses are chained edges)
  std::vector<std::vector<unsigned>> mAdjacency;
  float mNeighborRadius = 1.0f;                     // ScanSensor.cpp:37-39 defaults
  int mMaxNeighorLinks = 1;
  unsigned mMinLoopLength = 10;
  PointCloudSensor* mSensor;
  std::vector<Measurement::Ptr> mMeasurements;
  unsigned mLastVertex;
  bool mHasVertex, mLinkPrevious;
  Transform mLastOdometry;
  std::mutex mMutex;
};

}  // namespace slam3d_b200
