"""TEST INFRASTRUCTURE -- torch restatement of the reference's adaptive density control, statement by statement.

Follows /root/reference/src/gaussian_model.cpp: addDensificationStats :834-847, densifyAndPrune :806-824,
densifyAndClone :775-804, densifyAndSplit :729-773, densificationPostfix :653-727, prunePoints :597-651,
resetOpacity :567-575 / replaceTensorToOptimizer :577-595, general_utils.h:25-53 (inverse_sigmoid, build_rotation),
and the max_radii2D update of src/gaussian_mapper.cpp:739-742.  Runs on CPU or CUDA tensors.  Only tests/ may
import it (the package never does).

PINNED by the reference's own class (tests/test_reference_model.py, CPU): the UNMODIFIED src/gaussian_model.cpp is compiled
into oracle/_ref/ref_model.so (oracle/build_ref.py build_model(); type-only stand-ins for the absent Eigen / OpenCV / Sophus
headers, see oracle/ref_model_wrap.cpp) and runs on CPU tensors; every method of Model below is bit-identical to the
reference's -- parameters, Adam moments, step counts, exist_since_iter and the statistics vectors -- including the split's
normal draws (same seeded CPU generator).  Its numeric helpers -- inverse_sigmoid, build_rotation -- are also held
bit-identical to the unmodified header on its own (include/general_utils.h in oracle/_ref/ref_utils.so;
tests/test_reference_utils.py).
"""
import torch


# The two helpers below are pinned against the UNMODIFIED reference header (oracle/_ref/ref_utils.so,
# tests/test_reference_utils.py): bit-identical on the same device.
def inverse_sigmoid(x):
    """general_utils::inverse_sigmoid (include/general_utils.h:25-27)."""
    return torch.log(x / (1 - x))


def build_rotation(r):
    """general_utils::build_rotation (include/general_utils.h:29-60): quaternion (r, x, y, z), normalised here, -> [n, 3, 3]."""
    q = r / torch.sqrt(r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1] + r[:, 2] * r[:, 2] + r[:, 3] * r[:, 3]).unsqueeze(1)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=1).view(-1, 3, 3)

PARAMS = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")


class Model:
    """The seven parameter tensors + Adam moments + the four statistics vectors of GaussianModel."""

    def __init__(self, params, percent_dense=0.01):
        self.p = {k: params[k].detach().clone() for k in PARAMS}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}   # exp_avg
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}   # exp_avg_sq
        P, dev = self.p["xyz"].shape[0], self.p["xyz"].device
        self.exist_since_iter = torch.zeros(P, dtype=torch.int32, device=dev)
        self.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        self.denom = torch.zeros(P, 1, device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)
        self.percent_dense = percent_dense

    # -- gaussian_mapper.cpp:739-742 + gaussian_model.cpp:834-847
    def add_stats(self, radii, means2D_grad):
        f = radii > 0
        self.max_radii2D[f] = torch.max(self.max_radii2D[f], radii[f].to(self.max_radii2D.dtype))
        self.xyz_gradient_accum[f] += torch.norm(means2D_grad[f][:, :2], dim=-1, keepdim=True)
        self.denom[f] = self.denom[f] + 1

    def _scale(self):
        return torch.exp(self.p["scaling"])

    # -- :597-651
    def prune_points(self, mask):
        keep = ~mask
        for k in PARAMS:
            self.p[k] = self.p[k][keep]
            self.m[k] = self.m[k][keep].clone()
            self.v[k] = self.v[k][keep].clone()
        self.exist_since_iter = self.exist_since_iter[keep]
        self.xyz_gradient_accum = self.xyz_gradient_accum[keep]
        self.denom = self.denom[keep]
        self.max_radii2D = self.max_radii2D[keep]

    # -- :653-727
    def postfix(self, new, new_exist):
        for k in PARAMS:
            self.m[k] = torch.cat([self.m[k].clone(), torch.zeros_like(new[k])], dim=0)
            self.v[k] = torch.cat([self.v[k].clone(), torch.zeros_like(new[k])], dim=0)
            self.p[k] = torch.cat([self.p[k], new[k]], dim=0)
        self.exist_since_iter = torch.cat([self.exist_since_iter, new_exist], dim=0)
        P, dev = self.p["xyz"].shape[0], self.p["xyz"].device
        self.xyz_gradient_accum = torch.zeros(P, 1, device=dev)
        self.denom = torch.zeros(P, 1, device=dev)
        self.max_radii2D = torch.zeros(P, device=dev)

    # -- :775-804
    def densify_and_clone(self, grads, thr, extent):
        sel = torch.norm(grads, dim=-1) >= thr
        sel = torch.logical_and(sel, self._scale().max(dim=1).values <= self.percent_dense * extent)
        self.postfix({k: self.p[k][sel] for k in PARAMS}, self.exist_since_iter[sel])

    # -- :729-773; `normal01(n)` returns n x 3 standard-normal draws (at::normal(0, stds) = z * stds)
    def densify_and_split(self, grads, thr, extent, normal01, N=2):
        n_init = self.p["xyz"].shape[0]
        padded = torch.zeros(n_init, device=grads.device)
        padded[:grads.shape[0]] = grads.squeeze()
        sel = padded >= thr
        sel = torch.logical_and(sel, self._scale().max(dim=1).values > self.percent_dense * extent)
        stds = self._scale()[sel].repeat(N, 1)
        samples = normal01(stds.shape[0]) * stds
        R = build_rotation(self.p["rotation"][sel]).repeat(N, 1, 1)
        new = {k: self.p[k][sel].repeat(*([N] + [1] * (self.p[k].dim() - 1))) for k in PARAMS}
        new["xyz"] = torch.bmm(R, samples.unsqueeze(-1)).squeeze(-1) + self.p["xyz"][sel].repeat(N, 1)
        new["scaling"] = torch.log(self._scale()[sel].repeat(N, 1) / (0.8 * N))
        self.postfix(new, self.exist_since_iter[sel].repeat(N))
        n_new = N * int(sel.sum())
        self.prune_points(torch.cat([sel, torch.zeros(n_new, dtype=torch.bool, device=sel.device)]))
        return int(sel.sum())

    # -- :806-824
    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, normal01):
        grads = self.xyz_gradient_accum / self.denom
        grads[grads.isnan()] = 0.0
        self.densify_and_clone(grads, max_grad, extent)
        self.densify_and_split(grads, max_grad, extent, normal01)
        prune = (torch.sigmoid(self.p["opacity"]) < min_opacity).squeeze(-1)
        if max_screen_size:
            big_vs = self.max_radii2D > max_screen_size
            big_ws = self._scale().max(dim=1).values > 0.1 * extent
            prune = torch.logical_or(torch.logical_or(prune, big_vs), big_ws)
        self.prune_points(prune)

    # -- :297-384 increasePcd(tensor overload): Gaussians for the new points of a keyframe.  `dist2(points)` is distCUDA2
    #    (third_party/simple-knn; oracle/ingest_ref.py restates it); RGB2SH = include/sh_utils.h:32,133-135
    def increase_pcd(self, new_points, new_colors, iteration, dist2, max_sh_degree=3):
        n = new_points.shape[0]
        if n == 0:
            return
        dev = new_points.device
        C0 = 0.282094806432724  # `const float C0 = 0.28209479177387814f` (sh_utils.h:32) as the float it is
        fused = (new_colors - 0.5) / C0
        t = max_sh_degree + 1
        features = torch.zeros(n, 3, t * t, dtype=torch.float32, device=dev)
        features[:, 0:3, 0] = fused
        features[:, 3:, 1:] = 0.0
        lang = torch.zeros(n, self.p["lang_feat"].shape[1], dtype=torch.float32, device=dev)
        d2 = torch.clamp_min(dist2(new_points.clone()), 0.0000001)
        scales = torch.log(torch.sqrt(d2)).unsqueeze(1).repeat(1, 3)
        rots = torch.zeros(n, 4, device=dev)
        rots[:, 0] = 1
        x = 0.1 * torch.ones(n, 1, dtype=torch.float32, device=dev)
        opac = inverse_sigmoid(x)
        new_exist = torch.full((n,), iteration, dtype=torch.int32, device=dev)
        new = dict(xyz=new_points, features_dc=features[:, :, 0:1].transpose(1, 2).contiguous(),
                   features_rest=features[:, :, 1:].transpose(1, 2).contiguous(), lang_feat=lang.contiguous(),
                   opacity=opac, scaling=scales, rotation=rots)
        self.postfix(new, new_exist)

    # -- :567-595 (the clamp is against ones: a no-op; the moments ARE zeroed)
    def reset_opacity(self):
        op = torch.sigmoid(self.p["opacity"])
        x = torch.min(op, torch.ones_like(op * 0.01))
        self.p["opacity"] = inverse_sigmoid(x)
        self.m["opacity"] = torch.zeros_like(self.p["opacity"])
        self.v["opacity"] = torch.zeros_like(self.p["opacity"])
