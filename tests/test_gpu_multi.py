"""GPU, 2 ranks: particle sharding through the public API.
The whole cloud sampled on one GPU == the two shards sampled by two ranks (bit-identical samples,
identical all-reduced counters, identical autocorrelation up to summation order), including the two
batch-global couplings of the reference: the infinite-rate back-off and dwell-time resampling.
Over NCCL on two GPUs where the box has them (skipped otherwise) and, so that a single-GPU box covers the
same code, over gloo with both ranks on cuda:0."""
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
            assert torch.equal(full.cpu(), S1.cpu()), "sharded samples differ from the single-GPU run"
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
    import queue
    msgs = []
    try:
        msgs.append(out.get(timeout=10))
        while True:
            msgs.append(out.get(timeout=1))
    except queue.Empty:
        pass
    assert msgs == ["ok"], "\n".join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]


# ---------------------------------------------------------------------------------------------------------------
# batch-global couplings on a sharded cloud (sharded=True): back-off and resampling
# ---------------------------------------------------------------------------------------------------------------
def _backoff_cloud():
    """TestGaussian 2-d, eps = 1, L = 1: particle 40 (second shard) reaches exp(dH) = inf at iterations 3 and 4 of a
    6-iteration launch (found with the oracle; checked again below)."""
    rs = np.random.RandomState(22)
    d, N = 2, 64
    X0, V0 = rs.randn(d, N), rs.randn(d, N)
    X0[:, 40] = rs.randn(2) * 40
    V0[:, 40] = rs.randn(2) * 40
    return X0, V0


def _mj(X0, V0, lo, hi, device, resample, sharded, hp, seed=5):
    from mjhmc_b200.misc.distributions import TestGaussian
    from mjhmc_b200.samplers.markov_jump_hmc import MarkovJumpHMC
    from tests import helpers
    dist = helpers.pin_init(TestGaussian(X0.shape[0], hi - lo), X0[:, lo:hi])
    return MarkovJumpHMC(distribution=dist, V=V0[:, lo:hi], particle_offset=lo, seed=seed, resample=resample,
                         device=device, sharded=sharded, **hp)


def _coupling_worker(rank, world_size, port, out, backend):
    import torch.distributed as dist
    from mjhmc_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev)
    kw = dict(device_id=torch.device("cuda", dev)) if backend == "nccl" else {}
    dist.init_process_group(backend, rank=rank, world_size=world_size, **kw)
    try:
        device = "cuda:%d" % dev
        hp = dict(epsilon=1.0, beta=0.5, num_leapfrog_steps=1)
        X0, V0 = _backoff_cloud()
        N, n = X0.shape[1], 6
        lo, hi = parallel.shard_bounds(N, rank, world_size)
        # ---- back-off: the failing particle lives on rank 1; rank 0 must replay and retry at the same iterations
        s = _mj(X0, V0, lo, hi, device, False, True, hp)
        S = s.sample_device(n)
        full = parallel.allgather_samples(S)
        counters = parallel.allreduce_counters(s)
        attempts = s._attempt
        dwell = parallel.allgather_samples(torch.as_tensor(s.dwelling_times, device=device).reshape(1, 1, -1))
        # ---- resampling over the whole cloud
        s2 = _mj(X0, V0, lo, hi, device, True, True, dict(epsilon=0.7, beta=0.5, num_leapfrog_steps=2))
        np.random.seed(3)                 # the sorted uniforms of the resampler are host draws (rank 0 draws, all ranks share)
        R = s2.sample(5)
        Rfull = parallel.allgather_resampled(R, s2.resample_columns, 5 * N)
        if rank == 0:
            from oracle import mjhmc_oracle as orc
            s1 = _mj(X0, V0, 0, N, device, False, None, hp)
            S1 = s1.sample_device(n)
            assert s1._attempt == n + 2 == attempts, "two back-offs expected"
            assert torch.equal(full.cpu(), S1.cpu()), "sharded back-off differs from the single-GPU run"
            assert list(counters.values()) == parallel.local_counters(s1)
            assert s.epsilon == 1.0 and s.num_leapfrog_steps == 1
            np.testing.assert_array_equal(dwell.reshape(-1).cpu().numpy(), s1.dwelling_times)
            o = orc.OracleSampler("MarkovJumpHMC", orc.TestGaussianEnergy(1.0), X0, V=V0, draws=orc.PhiloxDraws(5),
                                  resample=False, **hp)
            Xo = o.sample(n)
            assert o.attempt == n + 2
            np.testing.assert_allclose(S1.cpu().numpy().reshape(2, -1), Xo, rtol=1e-10, atol=1e-12)
            assert list(counters.values()) == [o.counters()[k] for k in ("l", "f", "fl", "r", "E", "dEdX")]
            s3 = _mj(X0, V0, 0, N, device, True, None, dict(epsilon=0.7, beta=0.5, num_leapfrog_steps=2))
            np.random.seed(3)
            R1 = s3.sample(5)
            np.testing.assert_array_equal(Rfull, R1)
            out.put("ok")
    except Exception:   # noqa: BLE001 -- the parent prints the failing rank's traceback
        import traceback
        out.put("rank %d: %s" % (rank, traceback.format_exc()))
    finally:
        try:
            dist.barrier()
        except Exception:   # noqa: BLE001
            pass
        dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_sharded_backoff_and_resampling_equal_single_gpu(backend):
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_coupling_worker, args=(r, 2, port, out, backend)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    for p in procs:
        if p.is_alive():
            p.kill()
    import queue
    msgs = []
    try:
        msgs.append(out.get(timeout=10))
        while True:
            msgs.append(out.get(timeout=1))
    except queue.Empty:
        pass
    assert msgs == ["ok"], "\n".join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
