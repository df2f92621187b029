// TEST INFRASTRUCTURE -- stand-in for <opencv2/opencv.hpp>, which this image does not have.  include/tensor_utils.h (pulled in
// by include/gaussian_model.h:36) defines inline cv::Mat <-> tensor converters that GaussianModel never calls; they only have
// to compile.  Every member below that would do work aborts.
#pragma once
#include <cstdlib>

#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {

struct Vec3f {
    float v[3];
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

class Mat {
public:
    Mat() {}
    Mat(int r, int c, int type, void* d) : rows(r), cols(c), data((unsigned char*)d), type_(type) {}
    int channels() const { return type_ == CV_32FC3 ? 3 : 1; }
    Mat clone() const { std::abort(); }
    template <typename T> T& at(int, int) { std::abort(); }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

private:
    int type_ = CV_32FC1;
};

namespace cuda {
class GpuMat {
public:
    GpuMat() {}
    GpuMat(int r, int c, int type, void* d) : rows(r), cols(c), data((unsigned char*)d), type_(type) {}
    int channels() const { return type_ == CV_32FC3 ? 3 : 1; }
    GpuMat clone() const { std::abort(); }
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;

private:
    int type_ = CV_32FC1;
};
}  // namespace cuda

}  // namespace cv
