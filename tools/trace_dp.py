#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/trace_dp.py : kernel timeline (name, stream, start, duration) of a few data-parallel
kernel-path steps per rank, from the torch profiler's CUDA activity records -> gpurun_out/trace_dp_rank{r}.txt."""
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc, cam, up = bench.make_workload(rank, world, dev)
    kp = bench.KernelPath(sc, cam, up, dev, world)
    for i in range(10):
        kp.step(i)
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(4):
            kp.step(i)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), f"trace_{rank}.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"trace_dp_rank{rank}.txt"), "w") as f:
        for e in ev:
            f.write(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} s{e['args'].get('stream', '?'):<4} {e['name'][:70]}\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
