"""Multi-GPU host logic on the CPU: frame sharding and the final metric all-reduce over gloo,
world_size 2 (frames are independent; the only collective is the metric reduction)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from riders_b200 import dist as rdist


def test_shard_frames_partitions_exactly():
    for n in (0, 1, 7, 16, 64, 65):
        for world in (1, 2, 3, 4, 8):
            spans = [rdist.shard_frames(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert rdist.shard_frames(64, 3, 8) == (24, 32)          # config 3: batch 64 over 8 GPUs
    with pytest.raises(ValueError):
        rdist.shard_frames(4, 2, 2)


def test_shard_points_rebases_ragged_offsets():
    offs = [0, 5, 5, 12, 20]
    lo, hi, local = rdist.shard_points(offs, 1, 3)
    assert (lo, hi, local) == (5, 12, [0, 0, 7])


def test_single_process_reduction_is_a_plain_mean():
    means, n = rdist.reduce_metric_sums([6.0, 3.0], 3)
    assert n == 3 and np.allclose(means.numpy(), [2.0, 1.0])
    assert rdist.max_over_ranks(1.5) == 1.5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = rdist.init(backend="gloo")
    # 7 "frames" with a per-frame metric f*f: each rank sums its shard, one all-reduce gives the mean
    b, e = rdist.shard_frames(7, r, w)
    vals = np.arange(7, dtype=np.float64) ** 2
    means, n = rdist.reduce_metric_sums([vals[b:e].sum(), float(e - b)], e - b)
    slow = rdist.max_over_ranks(10.0 + r)
    total = rdist.sum_over_ranks(1.0)
    # gradient averaging (config 4): each rank holds grad = rank+1 -> mean 1.5
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(r + 1))
    calls = rdist.allreduce_gradients([p])
    # the per-step form used inside bench.py's timed loop: in-place sum of [metric sums ..., frame count]
    vec = rdist.all_reduce_sums_device(torch.tensor([vals[b:e].sum(), float(e - b)], dtype=torch.float64))
    assert abs(vec[0].item() / vec[1].item() - np.mean(vals)) < 1e-12
    # two data-parallel training steps (riders_b200.train.TrainStep logic on a CPU stand-in model): identical weights on both
    # ranks after the gradient all-reduce even though every rank sees different data
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 1)
    opt = torch.optim.Adam(lin.parameters(), lr=1e-2)
    for step in range(2):
        x = torch.full((3, 4), float(r + 1 + step))
        opt.zero_grad()
        lin(x).sum().backward()
        rdist.allreduce_gradients(list(lin.parameters()))
        opt.step()
    rdist.barrier()
    q.put((r, means.tolist(), n, slow, total, p.grad.tolist(), calls, lin.weight.detach().flatten().tolist()))
    rdist.shutdown()


def test_two_rank_gloo_metric_reduction_and_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.mean(np.arange(7, dtype=np.float64) ** 2)
    for r, means, n, slow, total, grad, calls, w in res:
        assert n == 7 and abs(means[0] - expect) < 1e-12 and abs(means[1] - 1.0) < 1e-12
        assert slow == 11.0 and total == 2.0
        assert grad == [1.5] * 5 and calls == 1
    assert res[0][7] == res[1][7]                                         # data-parallel replicas stay in lock step
