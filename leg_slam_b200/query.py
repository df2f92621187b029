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
