// TEST-ONLY stand-in for the interface the drop-in is compiled against.  A real integration includes the reference's
// own include/LaseCamCalCeres.h (its lines 11-29 declare the one struct and the four free functions below); this file
// only has to present the same names, parameter types and default arguments to the compiler, which it does through
// aliases, against the Eigen stand-in in this directory.
#pragma once
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

namespace clc_test_aliases {
using P3 = Eigen::Vector3d;
using P3List = std::vector<P3>;
using M4 = Eigen::Matrix4d;
}  // namespace clc_test_aliases

// one frame: board pose in the camera frame (rotation as a quaternion, identity by default; translation) plus the laser
// points on the board and the two end points on their fitted line
struct Oberserve {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Eigen::Quaterniond tagPose_Qca{1, 0, 0, 0};
  clc_test_aliases::P3 tagPose_tca = clc_test_aliases::P3::Zero();
  clc_test_aliases::P3List points, points_on_line;
};
using ObsList = std::vector<Oberserve>;

// closed form (returns T_lc), LM refinement (T_cl in/out), per-scan robust line, debug dump
void CamLaserCalClosedSolution(const ObsList obs, clc_test_aliases::M4& Tlc);
void CamLaserCalibration(const ObsList obs, clc_test_aliases::M4& Trc, bool use_linefitting_data = true,
                         bool use_boundary_constraint = false);
void LineFittingCeres(const clc_test_aliases::P3List Points, Eigen::Vector2d& Line);
void CalibrationTool_SavePlanePoints(const ObsList obs, const clc_test_aliases::M4 Tcl, const std::string path);
