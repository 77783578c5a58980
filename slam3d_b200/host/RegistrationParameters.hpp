// Host mirror of slam3d/sensor/pcl/RegistrationParameters.hpp:30-97 — same enum, same field names, same defaults,
// so application code that fills a RegistrationParameters keeps compiling.  Converted 1:1 into the C-ABI struct.
#pragma once

#include "../../include/s3d_b200.h"

namespace slam3d_b200 {

enum RegistrationAlgorithm { ICP, GICP, GICP_OMP, NDT, NDT_OMP };

struct RegistrationParameters {
  RegistrationAlgorithm registration_algorithm = GICP;
  double point_cloud_density = 0.2;
  double max_fitness_score = 2.0;
  double max_translation = 1.0;
  double max_rotation = 1.0;
  double euclidean_fitness_epsilon = 1.0;
  double transformation_epsilon = 1e-5;
  double max_correspondence_distance = 2.5;
  int maximum_iterations = 50;
  double rotation_epsilon = 2e-3;
  int correspondence_randomness = 20;
  int maximum_optimizer_iterations = 20;
  float resolution = 1.0;
  double step_size = 0.05;
  double outlier_ratio = 0.35;

  s3d_registration_parameters toC() const {
    s3d_registration_parameters c;
    c.registration_algorithm = static_cast<int32_t>(registration_algorithm);
    c.point_cloud_density = point_cloud_density; c.max_fitness_score = max_fitness_score;
    c.max_translation = max_translation; c.max_rotation = max_rotation;
    c.euclidean_fitness_epsilon = euclidean_fitness_epsilon; c.transformation_epsilon = transformation_epsilon;
    c.max_correspondence_distance = max_correspondence_distance; c.maximum_iterations = maximum_iterations;
    c.rotation_epsilon = rotation_epsilon; c.correspondence_randomness = correspondence_randomness;
    c.maximum_optimizer_iterations = maximum_optimizer_iterations; c.resolution = resolution;
    c.step_size = step_size; c.outlier_ratio = outlier_ratio;
    return c;
  }
};

}  // namespace slam3d_b200
