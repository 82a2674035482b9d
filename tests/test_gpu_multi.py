"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): particle sharding through the public API.
The whole cloud sampled on one GPU == the two shards sampled on two GPUs (bit-identical samples,
identical all-reduced counters, identical autocorrelation up to summation order)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run_cloud(X0, V0, lo, hi, n_iter, device):
    from mjhmc_b200.misc.distributions import RoughWell
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    from tests import helpers
    dist = helpers.pin_init(RoughWell(X0.shape[0], hi - lo, scale1=6, scale2=4), X0[:, lo:hi])
    s = MarkovJumpHMC(distribution=dist, V=V0[:, lo:hi], particle_offset=lo, epsilon=0.6, beta=0.5,
                      num_leapfrog_steps=5, seed=77, resample=False, device=device)
    S = s.sample_device(n_iter)
    return s, S


def _worker(rank, world_size, port, out):
    import torch.distributed as dist
    from mjhmc_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world_size, device_id=torch.device("cuda", rank))
    try:
        rs = np.random.RandomState(21)
        d, N, n_iter = 2, 1001, 16
        X0, V0 = rs.randn(d, N) * 4, rs.randn(d, N)
        lo, hi = parallel.shard_bounds(N, rank, world_size)
        s, S = _run_cloud(X0, V0, lo, hi, n_iter, "cuda:%d" % rank)
        counters = parallel.allreduce_counters(s)
        full = parallel.allgather_samples(S)
        ac = parallel.autocorrelation(S)
        if rank == 0:
            s1, S1 = _run_cloud(X0, V0, 0, N, n_iter, "cuda:0")
            assert torch.equal(full, S1), "sharded samples differ from the single-GPU run"
            # (no collective here: rank 1 is already waiting in the barrier below)
            assert list(counters.values()) == parallel.local_counters(s1)
            ac1 = parallel.autocorr_partial(S1).cpu().numpy()
            np.testing.assert_allclose(ac, ac1 / ac1[0], rtol=1e-12)
            out.put("ok")
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_gpu_shards_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(90)
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out.get(timeout=5) == "ok"
