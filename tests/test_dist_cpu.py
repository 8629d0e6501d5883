"""CPU, world_size 2 over gloo: the N>1 host path -- contiguous pose sharding with all hypotheses of a pose on one
rank, per-rank partial metric sums, ONE all-reduce of three fp64 values at the end."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import diffpose_nw_b200 as D
from oracle import diffpose_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pred = O.synthetic_poses(n, seed=5)[:, :, 2:]
        gt = O.synthetic_targets(O.synthetic_poses(n, seed=5), seed=6)
        lo, hi = D.shard_range(n, rank, world)
        p, g = O.root_centre(pred[lo:hi]), O.root_centre(gt[lo:hi])
        # per-rank partial sums exactly as dp_metrics produces them on the device: [sum mpjpe, sum p_mpjpe, count]
        per_pose_mpjpe = torch.norm(p - g, dim=-1).mean(dim=-1).double()
        pm = O.p_mpjpe_per_pose(p.numpy(), g.numpy())
        sums = torch.tensor([per_pose_mpjpe.sum().item(), float(pm.sum()), float(hi - lo)], dtype=torch.float64)
        out = D.reduce_metrics(sums)
        q.put((rank, lo, hi, out))
    finally:
        dist.destroy_process_group()


def test_two_rank_metric_reduction_equals_single_process():
    n, world = 101, 2          # odd on purpose: ranks get 51 and 50 poses
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 51, 51, 101)
    pred = O.synthetic_poses(n, seed=5)[:, :, 2:]
    gt = O.synthetic_targets(O.synthetic_poses(n, seed=5), seed=6)
    p, g = O.root_centre(pred), O.root_centre(gt)
    want_mpjpe = O.mpjpe(p, g).item() * 1000.0
    want_pmpjpe = float(O.p_mpjpe_per_pose(p.numpy(), g.numpy()).mean()) * 1000.0
    for _, _, _, (m, pm, cnt) in res:       # every rank holds the same reduced values
        assert cnt == n
        assert abs(m - want_mpjpe) < 1e-3 and abs(pm - want_pmpjpe) < 1e-6


def test_reduce_metrics_without_process_group():
    m, pm, cnt = D.reduce_metrics(torch.tensor([0.2, 0.1, 4.0], dtype=torch.float64))
    assert (round(m, 6), round(pm, 6), cnt) == (50.0, 25.0, 4.0)
    assert np.isnan(D.reduce_metrics(torch.zeros(3, dtype=torch.float64))[0])
