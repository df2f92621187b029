#!/usr/bin/env python
"""Generate tests/golden/model.npz from the UNMODIFIED reference GaussianModel class, on CPU (no GPU needed).

    python tests/golden/make_model_golden.py          # where /root/reference exists (oracle/build_ref.py build_model())

Runs src/gaussian_model.cpp as compiled into oracle/_ref/ref_model.so (CPU tensors; nothing of ours computes a stored value)
through: two Adam steps of trainingSetup's seven groups, three views of addDensificationStats, densifyAndPrune with and without
the screen-size threshold, resetOpacity, and the xyz learning-rate schedule; stores every input (including the standard-normal
draws the split consumed, re-drawn from the same seeded CPU generator) and every output.  The file pins oracle/densify_ref.py,
the C oracle's Adam and the mapper's schedule where the compiled reference is absent
(tests/test_reference_model.py::test_restatements_match_the_reference_model_golden).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import build_ref  # noqa: E402
import test_reference_model as T  # noqa: E402

P = 96
NAMES = ("xyz", "features_dc", "features_rest", "lang_feat", "opacity", "scaling", "rotation")


def state(ref, tag, out, only=NAMES, stats=True):
    for i, k in enumerate(NAMES):
        if k not in only:
            continue
        out[f"{tag}_p_{k}"] = ref.params()[i].detach().numpy().copy()
        st = ref.moments(i)
        out[f"{tag}_step_{k}"] = np.int64(st[0])
        out[f"{tag}_m_{k}"] = st[1].numpy().copy()
        out[f"{tag}_v_{k}"] = st[2].numpy().copy()
    if not stats:
        return
    out[f"{tag}_exist"] = ref.exist_since_iter.numpy().copy()
    out[f"{tag}_accum"] = ref.xyz_gradient_accum.numpy().copy()
    out[f"{tag}_denom"] = ref.denom.numpy().copy()
    out[f"{tag}_max_radii"] = ref.max_radii2D.numpy().copy()


def main(path=None):
    RM = build_ref.load_model()
    out = {}
    for case, max_screen_size in (("a", 0), ("b", 20)):
        t, exist, g = T.make_params(P, seed=500 + max_screen_size)
        ref = RM.GaussianModel(3)
        ref.set_state(t, exist, 2.5)
        ref.training_setup(**T.OPT)
        out[f"{case}_lrs"] = np.array(ref.lrs(), np.float64)
        for k, x in zip(NAMES, t):
            out[f"{case}_init_{k}"] = x.numpy().copy()
        out[f"{case}_init_exist"] = exist.numpy().copy()
        for s in range(2):
            grads = [torch.randn(*x.shape, generator=g) * 0.01 for x in t]
            for k, gr in zip(NAMES, grads):
                out[f"{case}_grad{s}_{k}"] = gr.numpy().copy()
            ref.set_grads(grads)
            ref.step()
        state(ref, f"{case}_stepped", out)
        for v in range(3):
            radii = torch.randint(-5, 40, (P,), generator=g, dtype=torch.int32).clamp_min(0)
            grad = torch.randn(P, 3, generator=g) * 3e-4
            out[f"{case}_view{v}_radii"], out[f"{case}_view{v}_grad"] = radii.numpy().copy(), grad.numpy().copy()
            f = radii > 0
            mr = ref.max_radii2D            # src/gaussian_mapper.cpp:739-742, the mapper's half of the statistics
            mr[f] = torch.max(mr[f], radii[f].to(mr.dtype))
            ref.max_radii2D = mr
            ref.add_densification_stats(grad, f)
        state(ref, f"{case}_stats", out, only=())
        seed = 900 + max_screen_size
        torch.manual_seed(seed)
        ref.densify_and_prune(2e-4, 0.05, 4.0, max_screen_size)
        state(ref, f"{case}_densified", out)
        # the split's draws: torch.manual_seed(seed); torch.empty(2 * selected, 3).normal_() -- the stream depends on the row
        # count, which the consumer knows once it has selected; the seed is what is stored
        out[f"{case}_normal_seed"] = np.int64(seed)
        out[f"{case}_args"] = np.array([2e-4, 0.05, 4.0, max_screen_size], np.float64)
        ref.reset_opacity()
        state(ref, f"{case}_reset", out, only=("opacity",), stats=False)
    # the xyz schedule of three configurations, every step listed
    t, exist, _ = T.make_params(4, seed=1)
    for i, (scale, max_steps) in enumerate(((2.5, 30000), (1.0, 500), (6.0, 30000))):
        ref = RM.GaussianModel(3)
        ref.set_state(t, exist, scale)
        ref.training_setup(**dict(T.OPT, position_lr_max_steps=max_steps))
        steps = np.array(list(range(0, 40)) + list(range(40, max_steps + 2000, 97)) + [max_steps - 1, max_steps, max_steps + 1])
        out[f"lr{i}_cfg"] = np.array([scale, max_steps], np.float64)
        out[f"lr{i}_steps"] = steps
        out[f"lr{i}_values"] = np.array([ref.update_learning_rate(int(s)) for s in steps], np.float64)
    path = path or os.path.join(HERE, "model.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
