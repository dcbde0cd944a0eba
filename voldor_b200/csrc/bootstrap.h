// Monocular bootstrap: relative pose of the first frame pair from the dense flow + closed-form depth.
// Behavioural source: reference voldor/geometry.cpp:267-285 (estimate_depth_closed_form) and :288-332
// (estimate_camera_pose_epipolar: cv::findEssentialMat(LMEDS, 0.999, 1.0) + cv::recoverPose on a 4-pixel grid).
#pragma once
#include <cstdio>

namespace vb {
namespace boot {

bool bootstrap_from_flow(const float* flow, int w, int h, const float* K9, float* R9, float* t3, float* depth);

}  // namespace boot
}  // namespace vb
