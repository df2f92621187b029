#!/usr/bin/env python
"""Where does the data-parallel step spend its time?  torchrun --nproc-per-node N tools/prof_dp_step.py
Per rank: forward+backward time, then the fused exchange+Adam split into barrier / kernel / barrier (CUDA events)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

def run(world, rank, dev, P, multimem):
    bench.P_GAUSS = P
    os.environ["LGS_DP_MULTIMEM"] = "1" if multimem else "0"
    sc, cam, up = bench.make_workload(rank, world, dev)
    os.environ["LGS_DP_OVERLAP"] = "0"  # this tool times the single-kernel exchange
    kp = bench.KernelPath(sc, cam, up, dev, world)
    for _ in range(5):
        kp.step()
    torch.cuda.synchronize()

    def ev():
        e = torch.cuda.Event(True)
        e.record()
        return e
    acc = [0.0] * 4
    n = 20
    for _ in range(n):
        dist.barrier()
        torch.cuda.synchronize()
        e0 = ev()
        kp.forward()
        kp.backward()
        e1 = ev()
        kp.dp.hg.barrier(channel=0)
        e2 = ev()
        kp.dp.step_count += 1
        import ctypes
        from leg_slam_b200 import _lib
        L = _lib.lib()
        d = kp.dp
        lr = (ctypes.c_double * len(d.lrs))(*d.lrs)
        _lib.check(L.lgs_dp_adam_shard(len(d.lrs), d._seg, lr, d.world, d.rank, d._gp, d._pp, ctypes.c_void_p(d.g_mc or None),
                                       ctypes.c_void_p(d.p_mc or None), d.begin, d.end, d.exp_avg.data_ptr(), d.exp_avg_sq.data_ptr(),
                                       0.9, 0.999, 1e-15, d.step_count, 0, torch.cuda.current_stream(dev).cuda_stream), "dp")
        e3 = ev()
        kp.dp.hp.barrier(channel=1)
        e4 = ev()
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(((e0, e1), (e1, e2), (e2, e3), (e3, e4))):
            acc[i] += a.elapsed_time(b) / n
    print(f"P={P} rank {rank}: fwd+bwd {acc[0]:.3f} ms | barrier {acc[1]:.3f} | exchange+Adam kernel {acc[2]:.3f} | barrier {acc[3]:.3f} | "
          f"multicast={kp.dp.uses_multicast}", flush=True)
    kp.dp.close()
    del kp
    torch.cuda.empty_cache()
    dist.barrier()


if __name__ == "__main__":
    world, rank, local = bench.dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    for P in (500_000, 1_000_000):
        for mm in (True, False):
            run(world, rank, dev, P, mm)
    dist.destroy_process_group()
