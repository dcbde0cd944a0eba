// voldor_b200 — drop-in boundary of the EM hot path (Python-binding level).
//
// Same C++-linkage signature as the reference's voldor/py_export.h:3-11, which the Cython module
// slam_py/install/pyvoldor_vo.pyx:5-12 declares as `cdef extern from "../../voldor/py_export.h"`.
// Runs one VO window: init + EM solve; outputs as in reference voldor/py_export.cpp:56-76.
#pragma once

#if defined(WIN32) || defined(_WIN32)
#define VB_EXPORT __declspec(dllexport)
#else
#define VB_EXPORT __attribute__((visibility("default")))
#endif

extern VB_EXPORT int py_voldor_wrapper(
	// inputs
	const float* flows, const float* disparity, const float* disparity_pconf,
	const float* depth_priors, const float* depth_prior_poses, const float* depth_prior_pconfs,
	const float fx, const float fy, const float cx, const float cy, const float basefocal,
	const int N, const int N_dp, const int w, const int h,
	const char* config,
	// outputs
	int& n_registered, float* poses, float* poses_covar, float* depth, float* depth_conf);
