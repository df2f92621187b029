"""CPU, world_size 2 over gloo: the data-parallel mapping iteration (leg_slam_b200.mapper) --
view sharding, flat-gradient all-reduce (sum), identical Adam on every replica -- must equal a
single-process iteration over the same K views (gradient accumulation), which is the parity
definition of SURVEY.md section 8e.  The CPU oracle stands in for the CUDA rasterizer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases  # noqa: F401  (sys.path)
from leg_slam_b200 import mapper as M, synthetic

K_VIEWS, P, W, H = 3, 600, 48, 32


def _window():
    sc = synthetic.make_scene(P, seed=41, mean_scale=0.08)
    cams = synthetic.make_cameras(K_VIEWS, W, H, seed=41)
    g = torch.Generator().manual_seed(42)
    win = [M.Keyframe(c, torch.rand(3, H, W, generator=g), torch.randn(64, H, W, generator=g),
                      torch.rand(1, H, W, generator=g) * 3) for c in cams]
    return sc, win


def _make_mapper(sc):
    import oracle_autograd
    return M.Mapper(sc, optimizer_factory=lambda g: torch.optim.Adam(g, lr=0.0, eps=1e-15),
                    render_fn=oracle_autograd.make_render_fn(torch.zeros(3)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc, win = _window()
        mp_ = _make_mapper(sc)
        assert mp_.world_size == world and mp_.rank == rank
        losses = []
        mp_.set_position_lr_schedule(3.2e-4, 3.2e-6, 0.01, 2)  # the per-iteration settings run on every replica alike
        for it in range(2):
            mp_.update_learning_rate(it)
            losses.append(float(mp_.train_step(win)))
        assert mp_.last_num_views == len(M.shard_views(K_VIEWS, rank, world))
        out[rank] = ({k: v.detach().clone() for k, v in mp_.params.items()}, losses)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_views_partitions_the_window():
    for k in (1, 3, 8):
        for w in (1, 2, 4, 8):
            parts = [M.shard_views(k, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(k))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_flat_grads_are_views_of_one_buffer():
    sc = synthetic.make_scene(16, seed=1)
    params = {k: torch.nn.Parameter(sc[k]) for k in M.PARAM_ORDER}
    fg = M.FlatGrads(params)
    assert fg.flat.numel() == 16 * 123  # 123 floats per Gaussian = 492 B all-reduce payload
    fg.attach(params)
    (params["xyz"].sum() * 2 + params["rotation"].sum() * 3).backward()
    assert params["xyz"].grad.data_ptr() == fg.flat.data_ptr()
    assert float(fg.flat[:48].sum()) == 96.0 and float(fg.flat[-64:].sum()) == 192.0
    fg.zero_()
    assert not params["rotation"].grad.any()


@pytest.mark.timeout(600)
def test_two_rank_data_parallel_equals_single_process_accumulation():
    # single process: all K views accumulated, then one Adam step (x2 iterations)
    sc, win = _window()
    torch.set_num_threads(1)
    single = _make_mapper(sc)
    single.set_position_lr_schedule(3.2e-4, 3.2e-6, 0.01, 2)
    single_losses = []
    for it in range(2):
        lr = single.update_learning_rate(it)
        assert single.optimizer.param_groups[0]["lr"] == lr
        single_losses.append(float(single.train_step(win)))
    assert abs(lr - 3.2e-5) < 1e-9  # half way down the log-linear schedule at step 1 of 2

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    (p0, l0), (p1, l1) = out[0], out[1]
    # replicas stay identical (same reduced gradient, same Adam)
    for k in M.PARAM_ORDER:
        assert torch.equal(p0[k], p1[k]), k
    # and equal the single-process K-view accumulation up to float summation order
    for k in M.PARAM_ORDER:
        ref = single.params[k].detach()
        err = (p0[k] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err <= 1e-5, (k, err)
    # per-rank losses add up to the single-process loss
    for it in range(2):
        assert abs((l0[it] + l1[it]) - single_losses[it]) <= 1e-4 * abs(single_losses[it])
