#!/usr/bin/env python
"""Multi-rank proof of the data-parallel data plane (`lgs_dp_adam_shard`, leg_slam_b200/dp.py), one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/dp_multirank_worker.py [--out result.json]

Part 1 -- the exchange kernel alone, no rendering, no atomics.  Every rank fills its symmetric gradient buffer with
seeded values and runs `FusedDPAdam.step()` through (a) the P2P branch (peer loads / stores), (b) the NVSwitch multimem
branch (multimem.ld_reduce / multimem.st; needs a multicast pointer), (c) the SPARSE P2P branch (a rank's copy of a
gradient row is loaded only where that rank reported the Gaussian as rendered; culled rows are exactly zero), each of them
also with the language-feature segment exchanged on the side stream (the overlapped schedule).  Required after every step and in every mode:
  * parameters bit-identical on all ranks (replicas), Adam moments (gathered from the shards) too;
  * "exact" data set (gradients are small integers times 2^-12, so every partial sum is exact in fp32 whatever the
    order): parameters and both moments bit-equal to the CPU oracle's Adam on the sum, in EVERY mode -- hence bit-identical
    across modes as well;
  * "random" data set (generic floats): the P2P modes bit-equal to the oracle's Adam on the fp32 rank-order sum
    ((g0 + g1) + g2 ...); the multimem modes (the switch adds in its own order) within 1e-6 relative of it.

Part 2 -- SURVEY.md section 8e "Parity definition": a K-view mapping iteration on G GPUs (views sharded over the ranks,
fused exchange) against K-view gradient accumulation on ONE GPU with the unmodified reference kernels (oracle/_ref) +
torch.optim.Adam: summed gradients <= 1e-3 relative per tensor, updated parameters equal up to Adam's sign-like first step
(the criterion of tests/test_gpu_parity.py::test_mapper_step_matches_reference_rasterizer_and_torch_adam), replicas
bit-identical.  K = G (one view per rank: the language-feature exchange starts at the backward hook) and K = 2 G.

Test infrastructure (it imports oracle/); exits non-zero when any requirement fails, rank 0 prints one JSON object."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

# Adam groups in the mapper's order and the reference's Replica learning rates (leg_slam_b200/mapper.py)
ROW = (3, 3, 45, 64, 1, 3, 4)
LRS = (3.2e-4, 2.5e-3, 1.25e-4, 1.5e-3, 0.05, 5e-3, 1e-3)
LATE = 3  # the language-feature segment


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def exchange_case(kind, P, world, steps):
    """Seeded initial parameters and per-(step, rank) gradients on the CPU; identical on every rank."""
    g = torch.Generator().manual_seed(1234 if kind == "exact" else 4321)
    n = P * sum(ROW)
    p0 = torch.randn(n, generator=g)
    grads = []
    for _ in range(steps):
        per_rank = []
        for _r in range(world):
            if kind == "exact":
                t = torch.randint(-64, 65, (n,), generator=g).float() * (2.0 ** -12)
            else:
                t = torch.randn(n, generator=g) * (10.0 ** torch.randint(-5, 0, (n,), generator=g).float())
            t[torch.rand(n, generator=g) < 0.1] = 0.0  # scattered exact zeros
            # culled Gaussians: radii == 0 and an exactly zero row in every tensor of this rank's gradient (60 % of the rows)
            radii = (torch.rand(P, generator=g) < 0.4).int() * torch.randint(1, 50, (P,), generator=g).int()
            off = 0
            for rl in ROW:
                t[off:off + P * rl].view(P, rl)[radii == 0] = 0.0
                off += P * rl
            per_rank.append((t, radii))
        grads.append(per_rank)
    return p0, grads


def oracle_trajectory(p0, grads, P):
    """CPU oracle Adam (oracle/lgs_oracle.c, bit-exact against the kernel on one rank) on the fp32 rank-order sum."""
    import oracle as O
    p, m, v = p0.numpy().copy(), np.zeros(p0.numel(), np.float32), np.zeros(p0.numel(), np.float32)
    starts = np.cumsum([0] + [P * r for r in ROW])
    traj = []
    for step, per_rank in enumerate(grads, start=1):
        s = per_rank[0][0].clone()
        for t, _radii in per_rank[1:]:
            s = s + t  # fp32, rank order
        sn = s.numpy()
        for i in range(len(ROW)):
            a, z = int(starts[i]), int(starts[i + 1])
            O.adam(p[a:z], sn[a:z], m[a:z], v[a:z], LRS[i], step=step)
        traj.append(p.copy())
    return traj, m, v


def all_equal_across_ranks(t, world):
    """True when `t` (a CUDA tensor) is bit-identical on every rank."""
    ref = t.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([1 if torch.equal(ref.view(torch.int32), t.view(torch.int32)) else 0], device=t.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    return bool(same.item())


def run_exchange(dev, rank, world, P, steps, result, fails):
    from leg_slam_b200 import dp as dp_mod
    sizes = [P * r for r in ROW]
    n = sum(sizes)
    for kind in ("exact", "random"):
        p0, grads = exchange_case(kind, P, world, steps)
        traj, m_ref, v_ref = oracle_trajectory(p0, grads, P)
        for mc, sparse in ((False, False), (True, False), (False, True)):
            for overlap in (False, True):
                name = f"{kind}/{'multimem' if mc else ('p2p-sparse' if sparse else 'p2p')}{'+overlap' if overlap else ''}"
                os.environ["LGS_DP_MULTIMEM"] = "1" if mc else "0"
                os.environ["LGS_DP_OVERLAP"] = "1" if overlap else "0"
                pflat, gflat = dp_mod.symmetric_empty(n, dev), dp_mod.symmetric_empty(n, dev)
                pflat.copy_(p0.to(dev))
                gflat.zero_()
                torch.cuda.synchronize(dev)
                dist.barrier()
                opt = dp_mod.FusedDPAdam(pflat, gflat, sizes, list(LRS), late_segment=LATE if overlap else None,
                                         rows=(P, list(ROW)) if sparse else None)
                row = dict(multicast=bool(opt.uses_multicast), overlap=bool(opt.overlap), sparse=bool(opt.sparse))
                assert opt.sparse == sparse
                if mc and not opt.uses_multicast:
                    row["skipped"] = "no multicast pointer on this fabric"
                    result[name] = row
                    opt.close()
                    continue
                ok_rep, ok_oracle, worst = True, True, 0.0
                for step in range(steps):
                    gflat.copy_(grads[step][rank][0].to(dev))
                    if sparse:  # what the mapper reports after each of its views: the radii of that forward
                        opt.mark_rows(grads[step][rank][1].to(dev), first=True)
                    opt.step(late_ready_at_hook=False)
                    opt.flush()
                    torch.cuda.synchronize(dev)
                    dist.barrier()
                    ok_rep &= all_equal_across_ranks(pflat, world)
                    got = pflat.cpu().numpy()
                    exact = np.array_equal(got.view(np.uint32), traj[step].view(np.uint32))
                    rel = float(np.abs(got - traj[step]).max() / np.abs(traj[step]).max())
                    worst = max(worst, rel)
                    ok_oracle &= exact
                m, v = opt.gather_moments()
                m_ok = np.array_equal(m.cpu().numpy().view(np.uint32), m_ref.view(np.uint32))
                v_ok = np.array_equal(v.cpu().numpy().view(np.uint32), v_ref.view(np.uint32))
                row.update(replicas_bit_identical=ok_rep, params_bit_equal_oracle=ok_oracle, max_rel_vs_oracle=worst,
                           exp_avg_bit_equal_oracle=m_ok, exp_avg_sq_bit_equal_oracle=v_ok)
                need_bits = (kind == "exact") or not mc  # P2P (dense or sparse) sums in rank order; the switch in its own
                if not ok_rep:
                    fails.append(f"{name}: replicas differ")
                if need_bits and not (ok_oracle and m_ok and v_ok):
                    fails.append(f"{name}: not bit-equal to the oracle (params {ok_oracle}, m {m_ok}, v {v_ok}, rel {worst:.3e})")
                if not need_bits and worst > 1e-6:
                    fails.append(f"{name}: {worst:.3e} > 1e-6 from the oracle")
                result[name] = row
                if rank == 0:
                    log(name, row)
                opt.close()
                del opt, pflat, gflat
    os.environ.pop("LGS_DP_MULTIMEM", None)
    os.environ.pop("LGS_DP_OVERLAP", None)


def run_kview(dev, rank, world, result, fails):
    """K-view step on G GPUs (fused exchange, default branch selection) vs reference kernels with K-view accumulation."""
    import bench
    import cases
    from leg_slam_b200 import mapper as M, synthetic
    W, H, P = 160, 120, 20000
    sc = synthetic.make_scene(P, seed=77, mean_scale=0.05, device=dev)
    for views_per_rank in (1, 2):
        K = world * views_per_rank
        cams = synthetic.make_cameras(K, W, H, seed=77)
        g = torch.Generator().manual_seed(78)
        win = [M.Keyframe(c.to(dev), torch.rand(3, H, W, generator=g).to(dev), torch.randn(64, 37, 37, generator=g).to(dev),
                          (torch.rand(1, H, W, generator=g) * 3).to(dev)) for c in cams]
        ours = M.Mapper(sc, sh_degree=3, dp_mode="fused")
        assert ours.dp is not None
        name = f"kview/K={K}/G={world}"
        row = dict(multicast=bool(ours.dp.uses_multicast), overlap=bool(ours.dp.overlap), sparse=bool(ours.dp.sparse), views_per_rank=views_per_rank)
        # summed gradients: the local flat buffers, all-reduced on a copy (the fused step itself never materialises the sum)
        mine = M.shard_views(K, rank, world)
        with torch.no_grad():
            ours._train_views_fused(win, mine)
        gsum = ours.grads.flat.clone()
        dist.all_reduce(gsum, op=dist.ReduceOp.SUM)
        # the real step (re-renders; same parameters, so the same gradients up to atomic order)
        ours.train_step(win)
        ours.dp.flush()
        torch.cuda.synchronize(dev)
        dist.barrier()
        row["replicas_bit_identical"] = all(all_equal_across_ranks(ours.params[k].data, world) for k in M.PARAM_ORDER)
        if not row["replicas_bit_identical"]:
            fails.append(f"{name}: replicas differ")
        if rank == 0:
            class _Shim(bench.E2EPath):  # the reference autograd glue of bench.py around oracle/_ref
                def __init__(self):
                    self.mapper = None
            bench.SH_DEGREE = 3
            refm = M.Mapper(sc, sh_degree=3, use_cuda_graph=False, optimizer_factory=lambda gr: torch.optim.Adam(gr, lr=0.0, eps=1e-15))
            refm.world_size, refm.rank = 1, 0  # the reference is single-GPU: all K views accumulate on this rank
            shim = _Shim()
            shim.mapper = refm
            refm.render_fn = shim._ref_render_fn()
            refm.train_step(win)
            torch.cuda.synchronize(dev)
            gref = refm.grads.flat
            worst, off = 0.0, 0
            for k in M.PARAM_ORDER:
                nk = refm.params[k].numel()
                e = cases.rel_err(gsum[off:off + nk].cpu().numpy(), gref[off:off + nk].cpu().numpy())
                row[f"grad_rel_{k}"] = e
                worst = max(worst, e)
                off += (nk + 3) & ~3
            if worst > 1e-3:
                fails.append(f"{name}: summed gradient {worst:.3e} > 1e-3 from the reference's K-view accumulation")
            for k in M.PARAM_ORDER:
                r = refm.params[k].detach().cpu().numpy()
                d = np.abs(ours.params[k].detach().cpu().numpy() - r)
                frac = float((d > 0.05 * M.DEFAULT_LRS[k]).mean())
                row[f"param_frac_off_{k}"] = frac
                if frac > 2e-3:
                    fails.append(f"{name}: {k}: {frac:.2e} of the updated parameters differ by more than 5 % of the step")
            log(name, row)
            del refm
        result[name] = row
        ours.dp.close()
        del ours
        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--P", type=int, default=4096 + 36)  # Gaussians in the exchange-only part (multiple of 4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--skip-kview", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    result, fails = dict(world=world, gpu=torch.cuda.get_device_name(dev)), []
    run_exchange(dev, rank, world, args.P, args.steps, result, fails)
    if not args.skip_kview:
        run_kview(dev, rank, world, result, fails)
    # a failure on any rank fails the job
    nf = torch.tensor([len(fails)], device=dev)
    dist.all_reduce(nf, op=dist.ReduceOp.SUM)
    result["failures"] = fails
    result["ok"] = int(nf.item()) == 0
    if rank == 0:
        line = json.dumps(result)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(nf.item()) == 0 else 1)


if __name__ == "__main__":
    main()
