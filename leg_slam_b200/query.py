"""Semantic query: cosine similarity of Gaussian language features against text embeddings
(reference eval/find_objects_gaussians.py:160-175), on liblgs.so."""
import torch

from . import _lib
from ._lib import check, ptr


def cosine_query(feats, text, simt=False):
    """feats [P,64], text [Q,64] or [64] -> similarities [P,Q] (or [P]); rows are normalised like
    F.normalize(eps=1e-12) inside the kernel.  Default: tcgen05 tensor-core kernel (3xTF32);
    simt=True runs the fp32 SIMT cross-check kernel."""
    L = _lib.lib()
    if not feats.is_cuda:
        raise _lib.LgsError("cosine_query has no CPU path")
    squeeze = text.dim() == 1
    text2 = (text[None] if squeeze else text).to(torch.float32).contiguous()
    feats = feats.to(torch.float32).contiguous()
    if feats.size(1) != 64 or text2.size(1) != 64:
        raise ValueError("language features are 64-D")
    P, Q = feats.size(0), text2.size(0)
    out = torch.empty((P, Q), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        fn = L.lgs_cosine_query_simt if simt else L.lgs_cosine_query
        check(fn(P, Q, ptr(feats), ptr(text2), ptr(out), torch.cuda.current_stream(feats.device).cuda_stream),
              "lgs_cosine_query")
    return out[:, 0] if squeeze else out


def relevance_scores(feats, text):
    """`1 - (s - min) / (max - min)` of the cosine similarity to ONE text embedding
    (find_objects_gaussians.py:170-175)."""
    L = _lib.lib()
    s = cosine_query(feats, text.reshape(-1)).contiguous()
    scratch = torch.empty(2, dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        check(L.lgs_minmax_invert(s.numel(), ptr(s), ptr(scratch),
                                  torch.cuda.current_stream(s.device).cuda_stream), "lgs_minmax_invert")
    return s


def cosine_image(rendered_lf, text):
    """Per-pixel query (reference eval/find_objects_gaussians.py:323): rendered_lf [64,H,W], text [Q,64] or [64] ->
    F.cosine_similarity(rendered_lf, text[:, None, None], dim=0) as [Q,H,W] (or [H,W])."""
    L = _lib.lib()
    if not rendered_lf.is_cuda:
        raise _lib.LgsError("cosine_image has no CPU path")
    squeeze = text.dim() == 1
    text2 = (text[None] if squeeze else text).to(torch.float32).contiguous()
    img = rendered_lf.to(torch.float32).contiguous()
    if img.dim() != 3 or img.size(0) != 64 or text2.size(1) != 64:
        raise ValueError("language features are 64-D: rendered_lf [64,H,W], text [Q,64]")
    H, W, Q = img.size(1), img.size(2), text2.size(0)
    out = torch.empty((Q, H, W), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        check(L.lgs_cosine_image(H * W, Q, ptr(img), ptr(text2), ptr(out), torch.cuda.current_stream(img.device).cuda_stream),
              "lgs_cosine_image")
    return out[0] if squeeze else out


def heat_colors(scores, column=0):
    """scores [P] or [P,Q] in [0,1] -> colors_precomp [P,3] (blue -> red ramp) for a heat-map render."""
    L = _lib.lib()
    s = scores.to(torch.float32).contiguous()
    P = s.size(0)
    stride = 1 if s.dim() == 1 else s.size(1)
    colors = torch.empty((P, 3), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        check(L.lgs_heat_colors(P, s.data_ptr() + 4 * int(column) if P else None, stride, ptr(colors),
                                torch.cuda.current_stream(s.device).cuda_stream), "lgs_heat_colors")
    return colors


def heatmap_render(means3D, opacities, scales, rotations, lang_feats, text, camera, background=None, scale_modifier=1.0):
    """BASELINE.json configs[4] end to end: cosine similarity of every Gaussian's language feature against ONE text embedding,
    the reference's min-max inversion (find_objects_gaussians.py:160-175), heat colours, and one forward through the
    colors_precomp path (no SH, no feature channels) -> (heat image [3,H,W], depth [1,H,W], radii [P], scores [P])."""
    from . import rasterize_points as rp
    dev = means3D.device
    scores = relevance_scores(lang_feats, text)
    colors = heat_colors(scores)
    bg = torch.zeros(3, dtype=torch.float32, device=dev) if background is None else background
    e = torch.empty(0, device=dev)
    _R, color, _lf, depth, radii, *_ = rp.rasterize_gaussians(bg, means3D, colors, e, opacities, scales, rotations, scale_modifier, e,
                                                           camera.viewmatrix, camera.projmatrix, camera.tanfovx, camera.tanfovy,
                                                           camera.height, camera.width, e, 0, camera.campos, False, False)
    return color, depth, radii, scores
