// TEST INFRASTRUCTURE -- stand-in for <opencv2/opencv.hpp>, which this image does not have.  include/tensor_utils.h (pulled in
// by include/gaussian_model.h:36) and include/camera.h (pulled in by include/gaussian_keyframe.h:30) define inline image
// helpers -- cv::Mat <-> tensor converters, undistortion maps -- that GaussianModel, GaussianRasterizer and GaussianRenderer
// never call; they only have to compile.  Every member below that would do work aborts.
#pragma once
#include <cstddef>
#include <cstdlib>

#define CV_32F 5
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv {

struct Vec3f {
    float v[3];
    Vec3f() : v{0, 0, 0} {}
    Vec3f(float a, float b, float c) : v{a, b, c} {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

struct Size {
    Size() {}
    template <typename A, typename B> Size(A w, B h) : width((int)w), height((int)h) {}
    int width = 0, height = 0;
};

class Mat {
public:
    Mat() {}
    Mat(int r, int c, int type, void* d) : rows(r), cols(c), data((unsigned char*)d), type_(type) {}
    Mat(Size s, int type, Vec3f) : rows(s.height), cols(s.width), type_(type) {}
    static Mat eye(int, int, int) { std::abort(); }
    int channels() const { return type_ == CV_32FC3 ? 3 : 1; }
    Mat clone() const { std::abort(); }
    template <typename T> T& at(int, int) { std::abort(); }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;

private:
    int type_ = CV_32FC1;
};

// `cv::Mat m = (cv::Mat_<float>(1, 4) << a, b, c, d);` (include/camera.h:130): the values are dropped
template <typename T>
struct MatCommaInitializer_ {
    MatCommaInitializer_ operator,(T) const { return *this; }
    operator Mat() const { return Mat(); }
};
template <typename T>
struct Mat_ {
    Mat_(int, int) {}
    MatCommaInitializer_<T> operator<<(T) const { return MatCommaInitializer_<T>(); }
};

struct _InputArray {
    _InputArray(const Mat&) {}
};
struct _OutputArray {
    _OutputArray(Mat&) {}
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1 };

inline void initUndistortRectifyMap(InputArray, InputArray, InputArray, InputArray, Size, int, OutputArray, OutputArray) { std::abort(); }
inline void remap(InputArray, OutputArray, InputArray, InputArray, int) { std::abort(); }

namespace cuda {
class GpuMat {
public:
    GpuMat() {}
    GpuMat(int r, int c, int type, void* d) : rows(r), cols(c), data((unsigned char*)d), type_(type) {}
    int channels() const { return type_ == CV_32FC3 ? 3 : 1; }
    GpuMat clone() const { std::abort(); }
    void upload(const Mat&) { std::abort(); }
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;

private:
    int type_ = CV_32FC1;
};
inline void resize(const GpuMat&, GpuMat&, Size) { std::abort(); }
}  // namespace cuda

}  // namespace cv
