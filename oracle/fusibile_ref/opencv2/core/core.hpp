// Stand-in for <opencv2/core/core.hpp>, ONLY so that the reference's gipuma/fusibile/camera.h parses when its kernel
// file (fusibile.cu) is compiled for the oracle (oracle/build.py build_fusibile_ref).  camera.h uses OpenCV for one
// host-side struct (`Camera`: three Mat_<float> members and Mat::eye in its constructor) that the kernel never
// touches; the kernel-side types (Camera_cu, CameraParameters_cu) are plain CUDA.  Test infrastructure, not product.
#pragma once
#define CV_32F 5
namespace cv {
struct Mat {
    static Mat eye(int, int, int) { return Mat(); }
};
template <typename T> struct Mat_ : Mat {
    Mat_() {}
    Mat_(const Mat &) {}
};
}  // namespace cv
