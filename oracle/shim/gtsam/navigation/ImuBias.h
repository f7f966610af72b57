// oracle/shim/gtsam/navigation/ImuBias.h — TEST INFRASTRUCTURE ONLY.  GTSAM is not installed in this image.  The reference's
// include/FeatureTracker.h only needs two GTSAM types to DECLARE members (initialBias, predVelocity) that none of the functions
// compiled into oracle/_ref (oracle/slice_reference.py) touches; these stand-ins let the real header parse.
#pragma once
#include <Eigen/Core>
namespace gtsam {
typedef Eigen::Vector3d Vector3;
namespace imuBias { struct ConstantBias {}; }
}  // namespace gtsam
