// mapper_driver.cpp -- a C++ consumer of the libtorch layers (liblgs_torch.so), written the way the reference's
// GaussianMapper::trainForOneIteration uses its own classes (src/gaussian_mapper.cpp:662-796): a GaussianModel with
// trainingSetup, per-iteration settings, GaussianRenderer::render inside mappingIterationBackward (fused loss, autograd through
// the rasterizer node, densification statistics) and the optimizer step.  No Python in the process.
//   mapper_driver in.bin out.bin
// in : int32 {P, W, H, n_iter, lf_h, lf_w}; float {FoVx, FoVy}; xyz[P,3] f_dc[P,1,3] f_rest[P,15,3] lf[P,64] opacity[P,1]
//      scaling[P,3] rotation[P,4] view[4,4] proj[4,4] campos[3] gt_image[3,H,W] gt_lf[64,lf_h,lf_w] gt_depth[1,H,W]
// out: float losses[n_iter], xyz[P,3], opacity[P,1], denom[P], max_radii2D[P], xyz learning rate of the last iteration
#include <torch/torch.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "gaussian_renderer.h"

static torch::Tensor rd(FILE* f, std::vector<int64_t> shape) {
    torch::Tensor t = torch::empty(shape, torch::kFloat32);
    const size_t n = (size_t)t.numel();
    if (n && fread(t.data_ptr<float>(), 4, n, f) != n) {
        fprintf(stderr, "short read\n");
        exit(3);
    }
    return t.to(torch::kCUDA);
}
static void wr(FILE* f, const torch::Tensor& t) {
    torch::Tensor h = t.detach().to(torch::kCPU, torch::kFloat32).contiguous();
    fwrite(h.data_ptr<float>(), 4, (size_t)h.numel(), f);
}

int main(int argc, char** argv) {
    if (argc != 3) return 1;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[6];
    float fov[2];
    if (fread(hdr, 4, 6, f) != 6 || fread(fov, 4, 2, f) != 2) return 3;
    const int P = hdr[0], W = hdr[1], H = hdr[2], n_iter = hdr[3], lf_h = hdr[4], lf_w = hdr[5];
    auto gaussians = std::make_shared<GaussianModel>(3);
    gaussians->xyz_ = rd(f, {P, 3}).requires_grad_();
    gaussians->features_dc_ = rd(f, {P, 1, 3}).requires_grad_();
    gaussians->features_rest_ = rd(f, {P, 15, 3}).requires_grad_();
    gaussians->language_features_ = rd(f, {P, 64}).requires_grad_();
    gaussians->opacity_ = rd(f, {P, 1}).requires_grad_();
    gaussians->scaling_ = rd(f, {P, 3}).requires_grad_();
    gaussians->rotation_ = rd(f, {P, 4}).requires_grad_();
    gaussians->exist_since_iter_ = torch::zeros({P}, torch::TensorOptions().dtype(torch::kInt32).device(torch::kCUDA));
    gaussians->max_radii2D_ = torch::zeros({P}, torch::TensorOptions().device(torch::kCUDA));
    gaussians->spatial_lr_scale_ = 1.0f;
    auto kf = std::make_shared<GaussianKeyframe>();
    kf->FoVx_ = fov[0];
    kf->FoVy_ = fov[1];
    kf->image_width_ = W;
    kf->image_height_ = H;
    kf->world_view_transform_ = rd(f, {4, 4});
    kf->full_proj_transform_ = rd(f, {4, 4});
    kf->camera_center_ = rd(f, {3});
    torch::Tensor gt_image = rd(f, {3, H, W});
    kf->language_features_ = rd(f, {64, lf_h, lf_w});
    torch::Tensor gt_depth = rd(f, {1, H, W});
    fclose(f);

    GaussianOptimizationParams opt;
    opt.position_lr_init_ = 3.2e-4f;
    opt.position_lr_final_ = 3.2e-6f;
    opt.position_lr_max_steps_ = n_iter;
    gaussians->trainingSetup(opt);
    GaussianPipelineParams pipe;
    torch::Tensor background = torch::zeros({3}, torch::TensorOptions().device(torch::kCUDA));
    torch::Tensor mask;  // no undistortion mask
    mask = torch::empty({0}, torch::TensorOptions().device(torch::kCUDA));
    std::vector<float> losses;
    float lr = 0.0f;
    for (int it = 0; it < n_iter; ++it) {
        gaussians->setShDegree(3);
        lr = gaussians->updateLearningRate(it);
        torch::Tensor loss = mappingIterationBackward(gaussians, kf, pipe, background, gt_image, gt_depth, mask, 0.2f, true);
        losses.push_back(loss.item<float>());
        mappingIterationStep(gaussians);
    }
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 4;
    fwrite(losses.data(), 4, losses.size(), o);
    wr(o, gaussians->xyz_);
    wr(o, gaussians->opacity_);
    wr(o, gaussians->denom_);
    wr(o, gaussians->max_radii2D_);
    fwrite(&lr, 4, 1, o);
    fclose(o);
    printf("mapper_driver ok: %d iterations, last loss %.6f, xyz lr %.3e\n", n_iter, losses.back(), lr);
    return 0;
}
