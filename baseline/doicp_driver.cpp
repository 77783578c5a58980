// doicp_driver.cpp — the PCL pin for the oracle (last row).  NOT part of the product and not built in
// this repository's environment (PCL is not installed here): it is the ready-to-run piece for the day a PCL >= 1.8.1 exists
// (slam3d-dependencies.cmake:24).  It runs, on the reference's own test clouds test/cloud1-4.bin, exactly what
// slam3d::PointCloudSensor runs per registration —
//     downsample() twice (PointCloudSensor.cpp:125-131 -> :190-201, pcl::VoxelGrid with a float leaf),
//     doICP<pcl::GeneralizedIterativeClosestPoint<PointXYZ, PointXYZ>> (:52-82): the setter sequence, the source/target swap,
//     align(result, guess.cast<float>()), getFitnessScore(max_correspondence_distance), the accept test —
// and writes tests/golden/pcl_<version>.json in the layout of tests/golden/golden.json, plus what the parity gates of
// BASELINE.json name: the voxel leaf index of every input point (hash), the k-NN / 1-NN correspondence indices (hash), final
// pose, fitness, iteration count.  tests/test_pcl_golden.py loads such a file when present and reports PCL-vs-oracle and
// PCL-vs-GPU deltas; without one it says "parity unpinned" as its skip reason.
//
// Build:   cmake -S baseline -B baseline/_build && cmake --build baseline/_build
// Run:     baseline/_build/doicp_driver <dir with cloud1.bin .. cloud4.bin> tests/golden
//
// The body of run_doicp() is written against PCL's public API only; it mirrors the call sequence of the reference
// (file:line cited per statement) without copying its source.
#include <pcl/filters/voxel_grid.h>
#include <pcl/kdtree/kdtree_flann.h>
#include <pcl/pcl_config.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <pcl/registration/gicp.h>

#include <Eigen/Geometry>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

typedef pcl::PointXYZ PointType;                 // PointCloudSensor.hpp:43
typedef pcl::PointCloud<PointType> PointCloud;   // PointCloudSensor.hpp:44

struct Params {                                  // RegistrationParameters.hpp:36-97 defaults
  double point_cloud_density = 0.2, max_fitness_score = 2.0, max_translation = 1.0, max_rotation = 1.0;
  double euclidean_fitness_epsilon = 1.0, transformation_epsilon = 1e-5, max_correspondence_distance = 2.5;
  int maximum_iterations = 50;
  double rotation_epsilon = 2e-3;
  int correspondence_randomness = 20, maximum_optimizer_iterations = 20;
};

static PointCloud::Ptr load_bin(const std::string& path) {  // KITTI layout: float32 x, y, z, intensity
  std::ifstream f(path, std::ios::binary);
  PointCloud::Ptr c(new PointCloud);
  float v[4];
  while (f.read(reinterpret_cast<char*>(v), sizeof v)) c->push_back(PointType(v[0], v[1], v[2]));
  return c;
}

static PointCloud::Ptr downsample(PointCloud::Ptr in, double leaf) {  // PointCloudSensor.cpp:190-201
  PointCloud::Ptr out(new PointCloud);
  if (in->size() > 0) {
    pcl::VoxelGrid<PointType> grid;
    grid.setLeafSize(leaf, leaf, leaf);
    grid.setInputCloud(in);
    grid.filter(*out);
  }
  return out;
}

// FNV-1a over raw bytes: the same hash tests/test_pcl_golden.py computes over the oracle's / the GPU's arrays
static uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

struct Outcome {
  bool converged = false;
  double fitness = 0;
  int status = 0;  // s3d_status: 0 ok, 1 too few points, 2 not converged / fitness, 3 too far from guess
  Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
  size_t n_source = 0, n_target = 0;
};

// align() + doICP<GICP>   PointCloudSensor.cpp:119-174, :52-82
static Outcome run_doicp(PointCloud::Ptr source, PointCloud::Ptr target, const Eigen::Isometry3d& guess, const Params& cfg) {
  Outcome o;
  PointCloud::Ptr fs = source, ft = target;
  if (cfg.point_cloud_density > 0) { fs = downsample(source, cfg.point_cloud_density); ft = downsample(target, cfg.point_cloud_density); }  // :125-131
  o.n_source = fs->size(); o.n_target = ft->size();
  if (ft->size() < 100 || fs->size() < 100) { o.status = 1; return o; }  // :134-135
  pcl::GeneralizedIterativeClosestPoint<PointType, PointType> icp;       // :141-144
  icp.setMaxCorrespondenceDistance(cfg.max_correspondence_distance);     // :58-64
  icp.setMaximumIterations(cfg.maximum_iterations);
  icp.setTransformationEpsilon(cfg.transformation_epsilon);
  icp.setEuclideanFitnessEpsilon(cfg.euclidean_fitness_epsilon);
  icp.setCorrespondenceRandomness(cfg.correspondence_randomness);
  icp.setMaximumOptimizerIterations(cfg.maximum_optimizer_iterations);
  icp.setRotationEpsilon(cfg.rotation_epsilon);
  PointCloud result;
  icp.setInputSource(ft);                                                // :68-69  (source and target are swapped)
  icp.setInputTarget(fs);
  icp.align(result, guess.matrix().cast<float>());                       // :70
  o.fitness = icp.getFitnessScore(cfg.max_correspondence_distance);      // :73
  o.converged = icp.hasConverged();
  o.T = Eigen::Isometry3d(Eigen::Isometry3f(icp.getFinalTransformation())).matrix();  // :80
  if (!o.converged || o.fitness > cfg.max_fitness_score) { o.status = 2; return o; }  // :74-77
  const Eigen::Isometry3d diff = guess.inverse() * Eigen::Isometry3d(o.T);            // :167-172
  if (diff.translation().norm() > cfg.max_translation || Eigen::AngleAxisd(diff.linear()).angle() > cfg.max_rotation) o.status = 3;
  return o;
}

static std::string hex(uint64_t v) { std::ostringstream s; s << std::hex << std::setw(16) << std::setfill('0') << v; return s.str(); }

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s <dir with cloud1.bin..cloud4.bin> <output dir>\n", argv[0]); return 2; }
  std::vector<PointCloud::Ptr> clouds;
  for (int i = 1; i <= 4; ++i) clouds.push_back(load_bin(std::string(argv[1]) + "/cloud" + std::to_string(i) + ".bin"));
  std::ostringstream js;
  // GICP's inner optimiser: BFGS (pcl/registration/bfgs.h) up to 1.13, Newton from 1.14 on (useBFGS() switches back); the tests
  // put the oracle into the same mode before they compare (oracle.set_gicp_optimizer)
#if PCL_VERSION_COMPARE(>=, 1, 14, 0)
  const char* inner_optimizer = "newton";
#else
  const char* inner_optimizer = "bfgs";
#endif
  js << std::setprecision(17) << "{\n  \"pcl_version\": \"" << PCL_VERSION_PRETTY << "\",\n  \"inner_optimizer\": \"" << inner_optimizer << "\",\n  \"clouds\": [";
  for (size_t i = 0; i < clouds.size(); ++i) js << (i ? ", " : "") << clouds[i]->size();
  js << "],\n  \"voxel\": {";
  // (a) VoxelGrid: number of output points and a hash of the output cloud (x, y, z floats in output order) per leaf size
  const double leaves[] = {0.05, 0.1, 0.2, 0.5, 1.0};
  for (size_t l = 0; l < 5; ++l) {
    js << (l ? ", " : "") << "\"" << leaves[l] << "\": [";
    for (size_t i = 0; i < clouds.size(); ++i) {
      PointCloud::Ptr out = downsample(clouds[i], leaves[l]);
      std::vector<float> xyz;
      for (const PointType& p : out->points) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
      js << (i ? ", " : "") << "{\"n_out\": " << out->size() << ", \"xyz_fnv1a\": \"" << hex(fnv1a(xyz.data(), xyz.size() * 4)) << "\"}";
    }
    js << "]";
  }
  // (b) exact NN semantics: kNN-20 of every point of cloud1@0.1 in itself, 1-NN of cloud2@0.1 in cloud1@0.1 (index hashes)
  js << "},\n  \"knn\": {";
  {
    PointCloud::Ptr f1 = downsample(clouds[0], 0.1), f2 = downsample(clouds[1], 0.1);
    pcl::KdTreeFLANN<PointType> tree;
    tree.setInputCloud(f1);
    std::vector<int> idx(20); std::vector<float> d2(20);
    std::vector<uint32_t> all_idx; std::vector<float> all_d2;
    for (const PointType& p : f1->points) { tree.nearestKSearch(p, 20, idx, d2); for (int j = 0; j < 20; ++j) { all_idx.push_back((uint32_t)idx[j]); all_d2.push_back(d2[j]); } }
    js << "\"cloud1@0.1,k=20\": {\"index_fnv1a\": \"" << hex(fnv1a(all_idx.data(), all_idx.size() * 4)) << "\", \"dist2_fnv1a\": \"" << hex(fnv1a(all_d2.data(), all_d2.size() * 4)) << "\"}";
    all_idx.clear(); all_d2.clear();
    std::vector<int> i1(1); std::vector<float> d1(1);
    for (const PointType& p : f2->points) { tree.nearestKSearch(p, 1, i1, d1); all_idx.push_back((uint32_t)i1[0]); all_d2.push_back(d1[0]); }
    js << ", \"nn cloud2@0.1 -> cloud1@0.1\": {\"index_fnv1a\": \"" << hex(fnv1a(all_idx.data(), all_idx.size() * 4)) << "\", \"dist2_fnv1a\": \"" << hex(fnv1a(all_d2.data(), all_d2.size() * 4)) << "\"}";
  }
  // (c) consecutive-pair registration, identity guess, slam3d's defaults at 0.1 and 0.2 m (BASELINE.json configs[0])
  js << "},\n  \"align\": {";
  bool first = true;
  for (double density : {0.1, 0.2})
    for (int a = 0; a < 3; ++a) {
      Params cfg; cfg.point_cloud_density = density;
      const Outcome o = run_doicp(clouds[a], clouds[a + 1], Eigen::Isometry3d::Identity(), cfg);
      js << (first ? "" : ", ") << "\n    \"cloud" << a + 1 << "->cloud" << a + 2 << "@" << density << "\": {\"status\": " << o.status << ", \"converged\": " << (o.converged ? 1 : 0)
         << ", \"fitness\": " << o.fitness << ", \"n_source\": " << o.n_source << ", \"n_target\": " << o.n_target << ", \"T\": [";
      for (int r = 0; r < 4; ++r) { js << (r ? ", [" : "["); for (int c = 0; c < 4; ++c) js << (c ? ", " : "") << o.T(r, c); js << "]"; }
      js << "]}";
      first = false;
    }
  js << "\n  }\n}\n";
  std::string ver = PCL_VERSION_PRETTY;
  const std::string path = std::string(argv[2]) + "/pcl_" + ver + ".json";
  std::ofstream(path) << js.str();
  std::printf("wrote %s\n", path.c_str());
  return 0;
}
