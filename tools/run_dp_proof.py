#!/usr/bin/env python
"""One-shot multi-GPU proof of the data-parallel data plane:  gpurun --gpus N -- python tools/run_dp_proof.py [N]
Runs tests/dp_multirank_worker.py on N GPUs (default: all) and keeps its JSON verdict and log under gpurun_out/dp_proof/
(copy them to profiles/ to commit)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dp_multirank import run_worker  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    out_dir = os.path.join(ROOT, "gpurun_out", "dp_proof")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"dp_proof_{n}gpu.json")
    r = run_worker(n, out, extra=sys.argv[2:])
    with open(os.path.join(out_dir, f"dp_proof_{n}gpu.log"), "w") as f:
        f.write(r.stderr[-20000:])
        f.write(f"\nrc={r.returncode}\n")
    print(r.stderr[-3000:])
    print(r.stdout[-6000:])
    sys.exit(r.returncode)


if __name__ == "__main__":
    main()
