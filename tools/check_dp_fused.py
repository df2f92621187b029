#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_dp_fused.py : fused peer-memory DP step vs NCCL all-reduce + Adam."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leg_slam_b200 import mapper as M, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    P, W, H = (int(os.environ.get("DP_P", "200000")), 640, 480)
    sc = synthetic.make_scene(P, seed=5, device=dev)
    cams = synthetic.make_cameras(max(8, world), W, H, seed=5)
    g = torch.Generator().manual_seed(6)
    win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                      (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams[:world]]
    res = {}
    for mode in ("allreduce", "fused"):
        m = M.Mapper(sc, sh_degree=3, dp_mode=mode)
        if mode == "fused":
            assert m.dp is not None
            if rank == 0:
                print("multicast:", m.dp.uses_multicast, flush=True)
        for _ in range(3):
            loss = m.train_step(win)
        torch.cuda.synchronize()
        # timing
        dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20):
            m.train_step(win)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        flat = torch.cat([m.params[k].detach().reshape(-1) for k in M.PARAM_ORDER])
        chk = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(chk, torch.stack([flat.double().sum(), flat.double().abs().sum()]))
        res[mode] = (flat.clone(), float(ms), chk, float(loss))
        if rank == 0:
            same = all(torch.equal(chk[0], c) for c in chk)
            print(f"{mode}: {float(ms):.3f} ms/step, replicas identical: {same}, loss {float(loss):.5f}", flush=True)
        del m
    a, b = res["allreduce"][0], res["fused"][0]
    err = float((a - b).abs().max() / a.abs().max())
    frac = float(((a - b).abs() > 1e-4).float().mean())
    if rank == 0:
        print(f"fused vs allreduce params: max rel err {err:.3e}, fraction differing > 1e-4: {frac:.2e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
