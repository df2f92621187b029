// TEST INFRASTRUCTURE -- stand-in for the reference's vendored Sophus se3.hpp, which needs the real Eigen (absent here).
// GaussianModel uses exactly two things of Sophus::SE3f: the (R, t) constructor in a default argument
// (include/gaussian_model.h:95-97) and matrix() (src/gaussian_model.cpp:394).  See ../../../../Eigen/Core.
#pragma once
#include <Eigen/Core>

namespace Sophus {

template <typename S>
class SE3 {
public:
    SE3() : R_(Eigen::Matrix<S, 3, 3>::Identity()), t_() {}
    SE3(const Eigen::Matrix<S, 3, 3>& R, const Eigen::Matrix<S, 3, 1>& t) : R_(R), t_(t) {}
    Eigen::Matrix<S, 4, 4> matrix() const {
        Eigen::Matrix<S, 4, 4> m = Eigen::Matrix<S, 4, 4>::Identity();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) m(i, j) = R_(i, j);
            m(i, 3) = t_(i);
        }
        return m;
    }

private:
    Eigen::Matrix<S, 3, 3> R_;
    Eigen::Matrix<S, 3, 1> t_;
};

typedef SE3<float> SE3f;
typedef SE3<double> SE3d;

}  // namespace Sophus
