"""CPU, world_size 2 over gloo: the host logic of the N>1 path (shard bounds, counter all-reduce,
sample all-gather, autocorrelation all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mjhmc_b200 import parallel
from oracle import mjhmc_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        N, d, T = 37, 3, 16
        rs = np.random.RandomState(0)
        Xg = rs.randn(d, T, N)
        lo, hi = parallel.shard_bounds(N, rank, world_size)
        local = torch.as_tensor(np.ascontiguousarray(Xg[:, :, lo:hi]))
        # counters
        c = parallel.allreduce_counters([rank + 1, 10 * (rank + 1), 0, 5, hi - lo, 7 * (hi - lo)])
        assert c == dict(l_count=3, f_count=30, fl_count=0, r_count=10, E_count=N, dEdX_count=7 * N)
        # samples
        full = parallel.allgather_samples(local)
        assert full.shape == (d, T, N)
        np.testing.assert_array_equal(full.numpy(), Xg)
        # autocorrelation: per-shard partial sums (numpy stand-in for the kernel) + all-reduce
        xl = local.numpy()
        part = torch.as_tensor(np.array([np.sum(xl * np.roll(xl, -t, axis=1)) for t in range(T)]))
        ac = parallel.autocorrelation(None, partial=part)
        ref = orc.fft_autocor(np.ascontiguousarray(Xg.transpose(0, 2, 1)))
        np.testing.assert_allclose(ac, ref, atol=1e-12)
        assert abs(parallel.effective_sample_size(ac) - orc.ess_from_autocor(ref)) < 1e-9
        # first failing iteration of a launch: min over the ranks (INT64_MAX = none) -> all ranks back off together
        big = (1 << 63) - 1
        assert parallel.allreduce_min(big) == big
        assert parallel.allreduce_min(5 if rank == 1 else big) == 5
        assert parallel.allreduce_min(3 + rank) == 3
        assert parallel.allreduce_sum_int(hi - lo) == N
        # dwell-time resampling over the sharded cloud: the plan reproduces the reference's global search
        # (markov_jump_hmc.py:321-328) when every rank resolves the draws that land in its own segments
        n_it = 5
        dwell_g = rs.exponential(size=(n_it, N))
        dwell_l = dwell_g[:, lo:hi]
        u = np.sort(rs.random_sample(n_it * N))
        plan = parallel.resample_plan(torch.as_tensor(dwell_l.sum(axis=1)), n_it * N, uniforms=u)
        cumul = np.cumsum(dwell_g.reshape(-1))
        assert abs(plan["total"] - cumul[-1]) < 1e-9
        want = np.searchsorted(cumul, u * cumul[-1], side="right")           # global flat index it * N + i
        own = (want % N >= lo) & (want % N < hi)
        pos = np.searchsorted(plan["bounds"], plan["r"], side="right")
        mine = pos % 2 == 1
        np.testing.assert_array_equal(mine, own)
        dl = dwell_l.copy()
        dl[:, 0] += plan["gaps"]
        got = np.searchsorted(np.cumsum(dl.reshape(-1)), plan["r"][mine], side="right")   # local flat index
        np.testing.assert_array_equal((got // (hi - lo)) * N + lo + got % (hi - lo), want[own])
        # shared uniforms: drawn on rank 0, identical everywhere
        np.random.seed(100 + rank)
        plan2 = parallel.resample_plan(torch.as_tensor(dwell_l.sum(axis=1)), 11)
        np.random.seed(100)
        np.testing.assert_array_equal(plan2["r"], np.sort(np.random.random(11)) * plan2["total"])
        cols = np.nonzero(mine)[0]
        full_r = parallel.allgather_resampled(np.vstack((cols, cols * 2.0)), cols, n_it * N)
        np.testing.assert_array_equal(full_r[0], np.arange(n_it * N))
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_cloud():
    for n in (1, 7, 100, 1000003):
        for w in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    X = np.arange(20).reshape(2, 10)
    blk, lo = parallel.shard_columns(X, 1, 3)
    np.testing.assert_array_equal(blk, X[:, 3:6])
    assert lo == 3


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
