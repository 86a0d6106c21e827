// Test-only declaration of the reference's solver interface (reference include/LaseCamCalCeres.h:11-29), written
// against the Eigen stand-in next to it.  A real integration includes the reference's own header instead; what
// matters is that the drop-in compiles against exactly these names, types and default arguments.
#pragma once
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

struct Oberserve {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Oberserve() : tagPose_Qca(1, 0, 0, 0), tagPose_tca(Eigen::Vector3d::Zero()) {}
  Eigen::Quaterniond tagPose_Qca;
  Eigen::Vector3d tagPose_tca;
  std::vector<Eigen::Vector3d> points;
  std::vector<Eigen::Vector3d> points_on_line;
};

void LineFittingCeres(const std::vector<Eigen::Vector3d> Points, Eigen::Vector2d& Line);
void CamLaserCalClosedSolution(const std::vector<Oberserve> obs, Eigen::Matrix4d& Tlc);
void CamLaserCalibration(const std::vector<Oberserve> obs, Eigen::Matrix4d& Trc, bool use_linefitting_data = true,
                         bool use_boundary_constraint = false);
void CalibrationTool_SavePlanePoints(const std::vector<Oberserve> obs, const Eigen::Matrix4d Tcl, const std::string path);
