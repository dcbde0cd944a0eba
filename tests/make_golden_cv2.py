"""Golden vectors for the two pieces of third-party arithmetic on the host side of the path: cv::Rodrigues in
both directions (reference voldor/utils.h:52-56, geometry.cpp:184,258, py_export.cpp) and cv::norm of a 3-vector.  The C++
OpenCV the reference links is absent from this image, but the cv2 wheel carries the same implementation, so the vectors
are generated with it HERE (no GPU needed):  python tests/make_golden_cv2.py  -> tests/golden/rodrigues_cv2.npz"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(7)
rvecs = np.concatenate([rng.normal(0, s, (400, 3)) for s in (1e-8, 1e-4, 0.01, 0.2, 1.0, 2.5)]).astype(np.float32)
rvecs[0] = 0
Rs = np.stack([cv2.Rodrigues(r.reshape(3, 1))[0].astype(np.float32) for r in rvecs])
# matrices that are only approximately rotations (what the pose pipeline feeds back): perturb and convert
noisy = (Rs.astype(np.float64) + rng.normal(0, 1e-4, Rs.shape)).astype(np.float32)
back = np.stack([cv2.Rodrigues(R.astype(np.float32))[0].reshape(3).astype(np.float32) for R in noisy])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rodrigues_cv2.npz")
np.savez_compressed(out, rvecs=rvecs, Rs=Rs.reshape(-1, 9), noisy_Rs=noisy.reshape(-1, 9), rvecs_of_noisy=back,
                    cv2_version=cv2.__version__)
print("written", out, rvecs.shape)
