// TEST INFRASTRUCTURE -- C entry points around the reference's OWN .ply library (third_party/tinyply/tinyply.{h,cpp}, compiled
// from where it lies by oracle/build_ref.py build_ply()) -> oracle/_ref/ref_ply.so (ctypes).  Oracle for SURVEY.md 8f row 3.
//
// The two functions below issue the SAME tinyply calls as GaussianModel::savePly / loadPly
// (/root/reference/src/gaussian_model.cpp:854-1075) in the same order, without the class (which needs stand-ins for Eigen /
// OpenCV / Sophus to compile: that build is oracle/_ref/ref_model.so, tests/test_reference_model.py, CPU tensors only) -- the
// byte format (header syntax, record packing) is then the reference's own code, the property sequence is restated from:
//   savePly : add_properties_to_element("vertex", ...) for xyz, normals, f_dc_*, f_rest_*, lf_*, opacity, scale_*, rot_*
//             (gaussian_model.cpp:992-1068), write(os, /*isBinary=*/true) (:1071)
//   loadPly : request_properties_from_element("vertex", ...) for x y z | f_dc_0..2 | f_rest_0..n-1 | opacity | scale_0..2 |
//             rot_0..3 (:882-905), read() (:907), buffers copied out (:918-941).  The reference's loader does not read lf_*;
//             ref_ply_read additionally requests them when asked to (same call pattern) so the round trip covers them.
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "tinyply.h"

static std::vector<std::string> names(const char* prefix, int n) {
    std::vector<std::string> v(n);
    for (int i = 0; i < n; ++i) v[i] = std::string(prefix) + std::to_string(i);
    return v;
}

extern "C" int ref_ply_write(const char* path, int P, int n_dc, int n_rest, int n_lf, const float* xyz, const float* normals,
                             const float* f_dc, const float* f_rest, const float* lf, const float* opacity, const float* scale,
                             const float* rot) {
    try {
        std::filebuf fb;
        if (!fb.open(path, std::ios::out | std::ios::binary)) return 1;
        std::ostream os(&fb);
        tinyply::PlyFile f;
        auto add = [&](const std::vector<std::string>& keys, const float* data) {
            f.add_properties_to_element("vertex", keys, tinyply::Type::FLOAT32, (size_t)P,
                                        reinterpret_cast<uint8_t*>(const_cast<float*>(data)), tinyply::Type::INVALID, 0);
        };
        add({"x", "y", "z"}, xyz);
        add({"nx", "ny", "nz"}, normals);
        add(names("f_dc_", n_dc), f_dc);
        add(names("f_rest_", n_rest), f_rest);
        add(names("lf_", n_lf), lf);
        add({"opacity"}, opacity);
        add(names("scale_", 3), scale);
        add(names("rot_", 4), rot);
        f.write(os, true);
        fb.close();
        return 0;
    } catch (const std::exception&) {
        return 2;
    }
}

// Header only: number of vertices and whether the file is binary; -1 on failure.
extern "C" long long ref_ply_count(const char* path) {
    try {
        std::ifstream is(path, std::ios::binary);
        if (!is.is_open()) return -1;
        tinyply::PlyFile f;
        f.parse_header(is);
        for (const auto& e : f.get_elements())
            if (e.name == "vertex") return (long long)e.size;
        return -1;
    } catch (const std::exception&) {
        return -1;
    }
}

extern "C" int ref_ply_read(const char* path, int max_sh_degree, int n_lf, float* xyz, float* f_dc, float* f_rest, float* lf,
                            float* opacity, float* scale, float* rot) {
    try {
        std::ifstream is(path, std::ios::binary);
        if (!is.is_open() || is.fail()) return 1;
        is.seekg(0, std::ios::beg);
        tinyply::PlyFile f;
        f.parse_header(is);
        const int n_f_rest = ((max_sh_degree + 1) * (max_sh_degree + 1) - 1) * 3;
        std::shared_ptr<tinyply::PlyData> d_xyz, d_dc, d_rest, d_lf, d_op, d_sc, d_rot;
        d_xyz = f.request_properties_from_element("vertex", {"x", "y", "z"});
        d_dc = f.request_properties_from_element("vertex", {"f_dc_0", "f_dc_1", "f_dc_2"});
        if (n_f_rest > 0) d_rest = f.request_properties_from_element("vertex", names("f_rest_", n_f_rest));
        if (n_lf > 0 && lf) d_lf = f.request_properties_from_element("vertex", names("lf_", n_lf));
        d_op = f.request_properties_from_element("vertex", {"opacity"});
        d_sc = f.request_properties_from_element("vertex", {"scale_0", "scale_1", "scale_2"});
        d_rot = f.request_properties_from_element("vertex", {"rot_0", "rot_1", "rot_2", "rot_3"});
        f.read(is);
        auto out = [](const std::shared_ptr<tinyply::PlyData>& d, float* dst) {
            if (d && dst) std::memcpy(dst, d->buffer.get(), d->buffer.size_bytes());
        };
        out(d_xyz, xyz); out(d_dc, f_dc); out(d_rest, f_rest); out(d_lf, lf); out(d_op, opacity); out(d_sc, scale); out(d_rot, rot);
        return 0;
    } catch (const std::exception&) {
        return 2;
    }
}
