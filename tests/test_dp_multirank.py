"""The multi-rank data plane on real GPUs: spawns tests/dp_multirank_worker.py under torchrun, one process per GPU (all the
GPUs of the box, at most 8), and requires its verdict.  What the worker checks is stated in its docstring: the fused
exchange + Adam kernel (P2P, multimem and overlapped schedules) bit-identical across ranks and to the CPU oracle, and a
K-view step on G GPUs against the reference kernels' K-view accumulation on one GPU."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_worker(n_gpus, out_path, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpus}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_multirank_worker.py"), "--out", out_path, *extra]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)


def test_dp_exchange_and_kview_step_multirank(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs (run tools/run_dp_proof.py under `gpurun --gpus N`)")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_rasterizer.so")):
        pytest.skip("oracle/_ref/ref_rasterizer.so was not shipped")
    n = min(torch.cuda.device_count(), 8)
    out = str(tmp_path / "dp.json")
    r = run_worker(n, out)
    assert r.returncode == 0, r.stderr[-4000:]
    res = json.load(open(out))
    assert res["ok"] and res["world"] == n, res["failures"]
    ran = [k for k, v in res.items() if isinstance(v, dict) and "skipped" not in v]
    assert any(k.startswith("exact/p2p") for k in ran) and any(k.startswith("random/p2p-sparse") for k in ran)
    assert any(k.startswith("kview/") for k in ran)
