// voldor_b200 — drop-in boundary of the EM hot path (Python-binding level).
//
// py_voldor_wrapper has the C++-linkage signature of the reference's voldor/py_export.h:3-11, which the Cython module
// slam_py/install/pyvoldor_vo.pyx:5-12 declares via `cdef extern`.  It runs one VO window (init + EM solve) and fills
// the outputs as reference voldor/py_export.cpp:56-76 does.  Parameter names and documentation are this project's.
#pragma once

#if defined(WIN32) || defined(_WIN32)
#define VB_EXPORT __declspec(dllexport)
#else
#define VB_EXPORT __attribute__((visibility("default")))
#endif

extern VB_EXPORT int py_voldor_wrapper(
    const float* flow_stack,             // [frames][rows][cols][2]
    const float* disparity_map,          // [rows][cols] or NULL (monocular)
    const float* disparity_confidence,   // [rows][cols] or NULL (= 1)
    const float* prior_stack,            // [priors][rows][cols] or NULL
    const float* prior_poses,            // [priors][6] rvec, tvec of each prior relative to the first frame
    const float* prior_confidences,      // [priors][rows][cols] or NULL (= 1)
    const float focal_x,
    const float focal_y,
    const float centre_x,
    const float centre_y,
    const float stereo_basefocal,        // baseline * focal, 0 for monocular
    const int frames,
    const int priors,
    const int cols,
    const int rows,
    const char* flags,                   // "--flag value ..." (csrc/config.h)
    int& registered_out,                 // number of frames kept; 0 = window failed
    float* poses_out,                    // [frames][6]
    float* pose_covariances_out,         // [frames][36]
    float* depth_out,                    // [rows][cols]
    float* depth_confidence_out);        // [rows][cols]
