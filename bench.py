#!/usr/bin/env python
"""Benchmark of the particle-parallel sampler loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

metric  : particle-leapfrog-steps/s == gradient evaluations/s, read from the bit-exact
          dEdX counter (every leapfrog step does exactly one dEdX on its particles:
          hmc_state.py:86-91, distributions.py:73-75).
step    : one launch of the fused sampler kernel = ITERS sampling iterations of the whole
          particle cloud (sampler.sample_device(ITERS)); samples of every iteration are written.
value   : device-timed (CUDA events around each step, L2 flushed between steps), state resident in HBM.
e2e     : the same work through the public API with HOST buffers: the particle state is uploaded
          from pinned host memory and sample(ITERS) returns a host numpy array, every step.
N > 1   : one process per GPU (torchrun); particles are independent chains, so each rank owns a
          contiguous shard of N_PER_GPU particles (weak scaling, no data-path collective); NCCL
          is used for the barrier, the max-over-ranks time and the int64 counter all-reduce.
--impl reference : the CPU restatement of the reference (oracle/, numpy float64) on all host cores,
          on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

# searched hyper-parameters of the reference (mjhmc/search/*/params*.json)
WORKLOADS = {
    # configs[1] of BASELINE.json: the configuration the metric is quoted on
    "roughwell2d_mjhmc": dict(dist="RoughWell", ndims=2, n=1_000_000, sampler="MarkovJumpHMC",
                              epsilon=3.0, beta=0.012314380146563053, L=25, iters=64,
                              source="search/MJHMC_rw/params.json"),
    "roughwell2d_control": dict(dist="RoughWell", ndims=2, n=1_000_000, sampler="ControlHMC",
                                epsilon=0.6687788963317871, beta=0.5385961532592773, L=22, iters=64,
                                source="search/control_rw/params_new.json"),
    "funnel10d_cthmc": dict(dist="Funnel", ndims=10, n=4_000_000, sampler="ContinuousTimeHMC",
                            epsilon=0.1, beta=0.5, L=10, iters=16, source="search/MJHMC_funnel/config.json midpoints"),
    # configs[2] / configs[3]: dense-contraction energies on the fp64 tensor pipe (DMMA)
    "gauss100d_mjhmc": dict(dist="GaussianRot", ndims=100, n=1_000_000, sampler="MarkovJumpHMC",
                            epsilon=1.4581446647644043, beta=0.009999999776482582, L=25, iters=2,
                            source="search/MJHMC_log_gauss/params_2.json; J = Q^T diag(10**linspace(-6,0,100)) Q"),
    "gauss100d_mjhmc_f32": dict(dist="GaussianRot", ndims=100, n=1_000_000, sampler="MarkovJumpHMC",
                                epsilon=1.4581446647644043, beta=0.009999999776482582, L=25, iters=2, dtype="float32",
                                source="search/MJHMC_log_gauss/params_2.json; fp32 states, tcgen05 3xTF32"),
    "pot100d_mjhmc": dict(dist="ProductOfT", ndims=100, n=1_000_000, sampler="MarkovJumpHMC",
                          epsilon=0.4827975928783417, beta=0.10154356807470322, L=10, iters=2,
                          source="search/MJHMC_poe_100/params.json; dense W = randn/sqrt(100)"),
    "pot100d_mjhmc_f32": dict(dist="ProductOfT", ndims=100, n=1_000_000, sampler="MarkovJumpHMC", dtype="float32",
                              epsilon=0.4827975928783417, beta=0.10154356807470322, L=10, iters=2,
                              source="search/MJHMC_poe_100/params.json; dense W = randn/sqrt(100); fp32 states, tcgen05 3xTF32"),
    # HBM-bound points of the fused leapfrog (one iteration per launch, L = 1)
    "testgauss2d_control_L1": dict(dist="TestGaussian", ndims=2, n=16_000_000, sampler="ControlHMC",
                                   epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point"),
    "gauss10d_control_L1": dict(dist="DiagGaussian", ndims=10, n=8_000_000, sampler="ControlHMC",
                                epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, diagonal Gaussian log_cond=1"),
    "gauss16d_control_L1": dict(dist="DiagGaussian", ndims=16, n=4_000_000, sampler="ControlHMC",
                                epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, diagonal Gaussian log_cond=1"),
    "gauss16d_control_L1_f32": dict(dist="DiagGaussian", ndims=16, n=8_000_000, sampler="ControlHMC", dtype="float32",
                                    epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, fp32 states"),
    "roughwell2d_control_L1": dict(dist="RoughWell", ndims=2, n=16_000_000, sampler="ControlHMC",
                                   epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point"),
    # the same points through the streaming kernel (TMA ring, csrc/stream_separable.cuh)
    "testgauss2d_control_L1_stream": dict(dist="TestGaussian", ndims=2, n=16_000_000, sampler="ControlHMC", kernel="stream",
                                          epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, streaming kernel"),
    "roughwell2d_control_L1_stream": dict(dist="RoughWell", ndims=2, n=16_000_000, sampler="ControlHMC", kernel="stream",
                                          epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, streaming kernel"),
    "gauss10d_control_L1_stream": dict(dist="DiagGaussian", ndims=10, n=8_000_000, sampler="ControlHMC", kernel="stream",
                                       epsilon=0.5, beta=0.1, L=1, iters=1,
                                       source="HBM roofline point, diagonal Gaussian log_cond=1, streaming kernel"),
    "gauss16d_control_L1_stream": dict(dist="DiagGaussian", ndims=16, n=4_000_000, sampler="ControlHMC", kernel="stream",
                                       epsilon=0.5, beta=0.1, L=1, iters=1,
                                       source="HBM roofline point, diagonal Gaussian log_cond=1, streaming kernel"),
    "gauss16d_control_L1_f32_stream": dict(dist="DiagGaussian", ndims=16, n=8_000_000, sampler="ControlHMC", dtype="float32",
                                           kernel="stream", epsilon=0.5, beta=0.1, L=1, iters=1,
                                           source="HBM roofline point, fp32 states, streaming kernel"),
    "roughwell2d_mjhmc_stream": dict(dist="RoughWell", ndims=2, n=1_000_000, sampler="MarkovJumpHMC", kernel="stream",
                                     epsilon=3.0, beta=0.012314380146563053, L=25, iters=64,
                                     source="search/MJHMC_rw/params.json; streaming kernel"),
    "roughwell2d_control_stream": dict(dist="RoughWell", ndims=2, n=1_000_000, sampler="ControlHMC", kernel="stream",
                                       epsilon=0.6687788963317871, beta=0.5385961532592773, L=22, iters=64,
                                       source="search/control_rw/params_new.json; streaming kernel"),
    "roughwell10d_control_L1_stream": dict(dist="RoughWell", ndims=10, n=8_000_000, sampler="ControlHMC", kernel="stream",
                                           epsilon=0.5, beta=0.1, L=1, iters=1, source="HBM roofline point, streaming kernel"),
    # configs[2] as the reference builds it: Gaussian(ndims=100) is DIAGONAL (distributions.py:257-263)
    "gauss100d_diag_mjhmc": dict(dist="DiagGaussian", log_cond=6, ndims=100, n=1_000_000, sampler="MarkovJumpHMC",
                                 epsilon=1.4581446647644043, beta=0.009999999776482582, L=25, iters=1,
                                 source="search/MJHMC_log_gauss/params_2.json; reference default diagonal J; streaming kernel"),
    "gauss100d_diag_mjhmc_x8": dict(dist="DiagGaussian", log_cond=6, ndims=100, n=1_000_000, sampler="MarkovJumpHMC",
                                    epsilon=1.4581446647644043, beta=0.009999999776482582, L=25, iters=8,
                                    source="search/MJHMC_log_gauss/params_2.json; reference default diagonal J; streaming kernel"),
    "gauss100d_diag_control_L1": dict(dist="DiagGaussian", log_cond=6, ndims=100, n=1_000_000, sampler="ControlHMC",
                                      epsilon=0.0010000000474974513, beta=0.009999999776482582, L=1, iters=1,
                                      source="search/control_log_gauss/params.json; reference default diagonal J; streaming kernel"),
    # configs[4]: Funnel 10-d ContinuousTimeHMC with the autocorrelation / ESS statistics reduced across GPUs
    "funnel10d_cthmc_ess": dict(dist="Funnel", scale=1.0, ndims=10, n=4_000_000, sampler="ContinuousTimeHMC",
                                epsilon=0.2, beta=0.9, L=25, iters=16, ess=dict(T=1024, n_lags=1024, block=500_000),
                                source="inside the ranges of search/MJHMC_funnel/config.json, where the fft_autocor curve "
                                       "crosses zero inside the window (lag ~300); ESS from 1024 recorded steps"),
}
DEFAULT_WORKLOAD = "roughwell2d_mjhmc"
DTYPE = "float64"          # the reference's arithmetic
METRIC = "particle_leapfrog_steps_per_s"
UNIT = "particle-leapfrog-steps/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_bf16_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            j = json.load(f)
        if "bf16_tflops_sustained" in j:
            return float(j["bf16_tflops_sustained"]), "measured dense bf16 %.0f TFLOP/s (MEASURED_PEAKS.json, sustained)" % j["bf16_tflops_sustained"]
    return 2250.0, "nominal dense bf16 2250 TFLOP/s (B200_PROFILING.md fallback)"


def algorithmic_bytes_per_launch(w, S=None):
    S = S or (4 if w.get("dtype") == "float32" else 8)
    """DESIGN.md 'Algorithmic traffic': per particle, one launch of `iters` iterations reads X,V,
    writes X,V, writes one sample column per iteration (+ dwell time of the last iteration for the
    jump samplers; + FLF cache scalar and flag read and written for MarkovJumpHMC)."""
    d, it = w["ndims"], w["iters"]
    per = 4 * d * S + it * d * S
    if w["sampler"] in ("ContinuousTimeHMC", "MarkovJumpHMC"):
        per += 8
    if w["sampler"] == "MarkovJumpHMC":
        per += 2 * S + 2
    return per * w["n"]


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on all host cores
# ----------------------------------------------------------------------------------------------
def _oracle_energy(w):
    from oracle import mjhmc_oracle as orc
    if w["dist"] == "RoughWell":
        return orc.RoughWellEnergy(100, 4)
    if w["dist"] == "TestGaussian":
        return orc.TestGaussianEnergy(1.0)
    if w["dist"] == "DiagGaussian":
        return orc.GaussianEnergy.log_conditioned(w["ndims"], w.get("log_cond", 1))
    if w["dist"] == "Funnel":
        return orc.FunnelEnergy(w.get("scale", 3.0))
    if w["dist"] == "GaussianRot":
        return orc.GaussianEnergy(_rotated_J(w["ndims"]))
    if w["dist"] == "ProductOfT":
        W, nu = _pot_params(w["ndims"])
        return orc.ProductOfTEnergy(W, nu)
    raise KeyError(w["dist"])


def _rotated_J(d, log_conditioning=6, seed=0):
    cond = 10 ** np.linspace(-log_conditioning, 0, d)
    Q, _ = np.linalg.qr(np.random.RandomState(seed).randn(d, d))
    return Q.T.dot(np.diag(cond)).dot(Q)


def _pot_params(d, seed=2015):
    rs = np.random.RandomState(seed)
    W = (rs.randn(d, d) / np.sqrt(d)).astype(np.float32)
    nu = (rs.rand(d) * 2 + 2.1).astype(np.float32)
    return W, nu


def _init_cloud(w, n, seed):
    rs = np.random.RandomState(seed)
    d = w["ndims"]
    if w["dist"] == "RoughWell":
        X = 100 * rs.randn(d, n)                               # distributions.py:308
    elif w["dist"] == "Funnel":
        x0 = rs.normal(scale=w.get("scale", 3.0), size=(1, n))
        X = np.vstack((x0, rs.normal(scale=np.exp(x0 / 2.), size=(d - 1, n))))
    elif w["dist"] == "GaussianRot":
        wv, Q = np.linalg.eigh(_rotated_J(d))
        X = Q.dot((1. / np.sqrt(wv)).reshape((-1, 1)) * rs.randn(d, n))
    elif w["dist"] == "DiagGaussian" and w.get("log_cond", 1) != 1:
        X = rs.randn(d, n) / np.sqrt(10 ** np.linspace(-w["log_cond"], 0, d)).reshape(-1, 1)   # distributions.py:277
    else:
        X = rs.randn(d, n)
    return X, rs.randn(d, n)


def _oracle_worker(args):
    w, n, steps, warmup, seed = args
    from oracle import mjhmc_oracle as orc
    X, V = _init_cloud(w, n, seed)
    s = orc.OracleSampler(w["sampler"], _oracle_energy(w), X, V=V, epsilon=w["epsilon"], beta=w["beta"],
                          num_leapfrog_steps=w["L"], draws=orc.FastNumpyDraws(seed), resample=False)
    for _ in range(warmup * w["iters"]):
        s.sampling_iteration()
    g0 = s.dEdX_count
    t0 = time.perf_counter()
    for _ in range(steps * w["iters"]):
        s.sampling_iteration()
    return s.dEdX_count - g0, time.perf_counter() - t0


def run_reference(w, steps, warmup, n_sample=None, quiet=False):
    """All host cores: the particle cloud is split over one process per core (chains are independent)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    # about 10-30 s of CPU work: scale the per-core particle count with the dimension
    n_sample = n_sample or max(1000, 200_000 // w["ndims"]) * cores
    per = max(1, n_sample // cores)
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_worker, [(w, per, steps, warmup, 100 + r) for r in range(cores)])
    wall = time.perf_counter() - t0
    grads = sum(r[0] for r in res)
    t = max(r[1] for r in res)
    value = grads / t
    return dict(value=value, unit=UNIT, cores=cores, kind="port",
                sample="%d particles (%d per core x %d cores), %d steps x %d iterations, numpy float64 oracle port "
                       "with vectorised draws; wall incl. process start %.1fs" % (per * cores, per, cores, steps, w["iters"], wall),
                ms_per_step=1e3 * t / max(steps, 1), n_particles=per * cores)


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    """Samples SM clock and throttle reasons DURING the timed region (NVML polled from a thread every
    10 ms; falls back to `nvidia-smi -lms` when pynvml is unavailable)."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.proc = None

    def _poll(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for nm, bit in names.items():
                if r & bit:
                    self.reasons.add(nm)
            self._stop.wait(0.01)

    def start(self):
        try:
            import pynvml  # noqa: F401
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:   # noqa: BLE001
            self._thread = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.sm:
            out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.max_mhz, samples=len(self.sm),
                       sm_mhz_min=float(np.min(self.sm)))
        out["reasons"] = sorted(self.reasons)
        return out


def _oracle_ess_worker(args):
    w, n, seed = args
    from oracle import mjhmc_oracle as orc
    X, V = _init_cloud(w, n, seed)
    s = orc.OracleSampler(w["sampler"], _oracle_energy(w), X, V=V, epsilon=w["epsilon"], beta=w["beta"],
                          num_leapfrog_steps=w["L"], draws=orc.FastNumpyDraws(seed), resample=False)
    t0 = time.perf_counter()
    S = s.sample(w["ess"]["T"], preserve_order=True)                   # (d, n, T)
    t_sample = time.perf_counter() - t0
    f = np.fft.fft(S, axis=-1)
    ac = np.real(np.sum(np.fft.ifft(f * np.conj(f), axis=-1), axis=(0, 1)))   # un-normalised, summed over dims and particles
    return ac, t_sample, time.perf_counter() - t0


def run_reference_ess(w, n_per_core=2000):
    """ESS/s of the numpy port on all host cores: T recorded steps per chain, fft_autocor over all chains."""
    import multiprocessing as mp
    from oracle import mjhmc_oracle as orc
    cores = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_oracle_ess_worker, [(w, n_per_core, 300 + r) for r in range(cores)])
    ac = sum(r[0] for r in res)
    ac = ac / ac[0]
    ess = _ess_from_curve(ac[:w["ess"]["n_lags"]], len(ac))
    t = max(r[2] for r in res)
    n = n_per_core * cores
    return dict(ess_per_chain=ess, chains=n, seconds=t, ess_per_s=ess * n / t, cores=cores, kind="port",
                sample="%d chains (%d per core x %d cores), %d recorded steps each" % (n, n_per_core, cores, w["ess"]["T"]))


def _ess_from_curve(ac, T):
    """ESS = T / (1 + 2 sum_{tau>=1}^{first rho<0} rho_tau) on the first len(ac) lags of the fft_autocor curve."""
    s = 0.0
    for tau in range(1, len(ac)):
        if ac[tau] < 0:
            break
        s += ac[tau]
    return T / (1.0 + 2.0 * s)


def _device_cloud(w, n, seed, dev, tdtype):
    """The synthetic particle cloud of _init_cloud drawn on the device (the secondary workloads of the default
    run: nothing of theirs is compared with the CPU arm, so the host need not hold a copy)."""
    import torch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    d = w["ndims"]
    rn = lambda *shape: torch.randn(*shape, generator=g, device=dev, dtype=torch.float64)
    if w["dist"] == "RoughWell":
        X = 100 * rn(d, n)
    elif w["dist"] == "Funnel":
        x0 = w.get("scale", 3.0) * rn(1, n)
        X = torch.cat((x0, torch.exp(x0 / 2.) * rn(d - 1, n)), dim=0)
    elif w["dist"] == "GaussianRot":
        wv, Q = np.linalg.eigh(_rotated_J(d))
        X = torch.as_tensor(Q, device=dev) @ (torch.as_tensor((1. / np.sqrt(wv)).reshape((-1, 1)), device=dev) * rn(d, n))
    elif w["dist"] == "DiagGaussian" and w.get("log_cond", 1) != 1:
        X = rn(d, n) / torch.as_tensor(np.sqrt(10 ** np.linspace(-w["log_cond"], 0, d)).reshape(-1, 1), device=dev)
    else:
        X = rn(d, n)
    return X.to(tdtype), rn(d, n).to(tdtype)


def make_sampler(w, rank, dtype=None, seed=2024, n=None, device_init=None):
    dtype = dtype or w.get("dtype", DTYPE)
    from mjhmc_b200.misc import distributions as D
    from mjhmc_b200.samplers import markov_jump_hmc as S
    n, d = n or w["n"], w["ndims"]
    if w["dist"] == "RoughWell":
        dist = D.RoughWell(ndims=d, nbatch=8)
    elif w["dist"] == "TestGaussian":
        dist = D.TestGaussian(ndims=d, nbatch=8)
    elif w["dist"] == "DiagGaussian":
        dist = D.Gaussian(ndims=d, nbatch=8, log_conditioning=w.get("log_cond", 1))
    elif w["dist"] == "Funnel":
        dist = D.Funnel(scale=w.get("scale", 3.0), ndims=d, nbatch=8)
    elif w["dist"] == "GaussianRot":
        dist = D.Gaussian(ndims=d, nbatch=8, J=_rotated_J(d))
    elif w["dist"] == "ProductOfT":
        W, nu = _pot_params(d)
        dist = D.ProductOfT(ndims=d, nbasis=d, nbatch=8, W=W, lognu=np.log(nu.astype(np.float64)))
    else:
        raise KeyError(w["dist"])
    dist.nbatch = n
    if device_init is not None:
        import torch
        X0, V0 = _device_cloud(w, n, 1000 + rank, device_init, torch.float32 if dtype == "float32" else torch.float64)
    else:
        X0, V0 = _init_cloud(w, n, 1000 + rank)
    dist.gen_init_X = lambda: setattr(dist, "Xinit", X0)
    kw = dict(resample=False) if w["sampler"] in ("ContinuousTimeHMC", "MarkovJumpHMC") else {}
    if w.get("kernel"):
        kw["kernel"] = w["kernel"]
    s = getattr(S, w["sampler"])(distribution=dist, epsilon=w["epsilon"], beta=w["beta"], num_leapfrog_steps=w["L"],
                                 V=V0, dtype=dtype, seed=seed, particle_offset=rank * n, **kw)
    return s, dist, X0, V0


class Ctx(object):
    """Process-wide plumbing of one bench run: rank / world, device, barrier, the L2 flush buffer."""

    def __init__(self):
        import torch
        import torch.distributed as dist_pkg
        self.torch, self.dist = torch, dist_pkg
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        # one rank per GPU: give every rank its own slice of the host cores (the e2e path is a host-side memcpy
        # pipeline; eight ranks bouncing over the same cores cost bandwidth)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = len(cores) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", self.world)))
            if self.world > 1 and per >= 1:
                os.sched_setaffinity(0, cores[self.local_rank * per:(self.local_rank + 1) * per])
        except (AttributeError, OSError):
            pass
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist_pkg.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=self.dev)   # 256 MB > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op, dtype):
        t = self.torch.tensor(vals, dtype=dtype, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=op)
        return t.tolist()

    def max_f(self, *vals):
        return self.reduce(list(vals), self.dist.ReduceOp.MAX, self.torch.float64)

    def sum_i(self, *vals):
        return self.reduce(list(vals), self.dist.ReduceOp.SUM, self.torch.int64)


def measured_tensor_peaks(ctx):
    """cuBLAS GEMM rates of THIS GPU for fp64 (DMMA, the denominator of the fp64 dense kernels' roofline:
    MEASURED_PEAKS.json only holds a bf16 figure) and, for reference, tf32; best of 5 runs of a 6144^3 (fp64) /
    8192^3 (tf32) torch.matmul, CUDA events."""
    torch = ctx.torch
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, dt, n, tf32 in (("fp64_tflops", torch.float64, 6144, False), ("tf32_tflops", torch.float32, 8192, True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            a = torch.randn(n, n, device=ctx.dev, dtype=dt)
            b = torch.randn(n, n, device=ctx.dev, dtype=dt)
            torch.matmul(a, b)
            best = 1e30
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
            del a, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()
    return out


def ncu_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the workload's sampler kernel, from the committed
    ncu --set full capture (profiles/traffic.json: written by tools/ncu_summary.py from the .ncu-rep, with the file
    it came from).  None when no capture of this workload is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f).get(name)
    if not t:
        return None, None
    return t["dram_bytes"], t["source"]


def measure_device(ctx, name, w, steps, warmup, n=None, host_init=False, clocks=None):
    """Device-timed run of one workload: `steps` launches of the fused sampler kernel (each `iters` sampling
    iterations over the cloud), L2 flushed between steps, CUDA events, max over ranks."""
    torch = ctx.torch
    n = n or w["n"]
    sampler, dist, X0, V0 = make_sampler(w, ctx.rank, n=n, device_init=None if host_init else ctx.dev)
    eng = sampler._engine
    iters = w["iters"]
    for _ in range(warmup):
        sampler.sample_device(iters)
    ctx.barrier()
    if clocks is not None:
        clocks.start()
    g0, x0, launches0 = dist.dEdX_count, sampler.grad_evals_executed, eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    eng.kernel_events = []                    # CUDA events right around each sampler-kernel launch (roofline)
    t_wall0 = time.perf_counter()
    for k in range(steps):
        ctx.flush.fill_(float(k))             # evict the state from L2 between timed steps
        ev[k][0].record()
        out = sampler.sample_device(iters)
        ev[k][1].record()
        del out
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop() if clocks is not None else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    kernel_ms = [a.elapsed_time(b) for a, b in eng.kernel_events]
    eng.kernel_events = None
    if os.environ.get("MJHMC_BENCH_DEBUG"):            # developer aid: per-launch kernel times of every workload
        print("[bench debug] %s kernel ms: %s" % (name, " ".join("%.4f" % t for t in kernel_ms)), file=sys.stderr)
    grads = sampler.grad_evals_executed - x0          # leapfrog steps actually integrated on the device
    grads_ref = dist.dEdX_count - g0                  # the reference's dEdX_count accounting
    launches = eng.launches - launches0
    ms_max, launch_ms = ctx.max_f(ms, float(np.mean(kernel_ms)) if kernel_ms else ms / steps)
    grads_all, launches_all, grads_ref_all = [int(v) for v in ctx.sum_i(grads, launches, grads_ref)]
    return dict(name=name, w=w, n=n, sampler=sampler, dist=dist, X0=X0, V0=V0, steps=steps, warmup=warmup,
                ms_max=ms_max, launch_ms=launch_ms, grads_all=grads_all, launches_all=launches_all,
                grads_ref_all=grads_ref_all, value=grads_all / (ms_max * 1e-3), clk=clk, t_wall=t_wall)


def roofline_of(ctx, m, peaks):
    """The roofline object of one measured workload."""
    w, name, world = m["w"], m["name"], ctx.world
    peak, peak_src = measured_peak()
    ww = dict(w, n=m["n"])
    alg = algorithmic_bytes_per_launch(ww)
    launch_ms = m["launch_ms"]
    achieved = alg / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(name)
    if w["dist"] in ("GaussianRot", "ProductOfT"):
        # dense-contraction energies: algorithmic flops per leapfrog step = 2 d^2 (S x) resp. 4 d nb (W^T x, W G)
        fl = (2 if w["dist"] == "GaussianRot" else 4) * w["ndims"] ** 2
        tf = m["grads_all"] * fl / (m["ms_max"] * 1e-3) / 1e12 / world
        f32 = w.get("dtype") == "float32"
        if f32:
            # bf16x3 operands: one product = 6 bf16 MMAs (csrc/dense_tc.cu); the denominator is the measured dense bf16
            # rate of this pool's B200s (MEASURED_PEAKS.json, sustained figure: the kernel runs for milliseconds)
            bf16 = measured_bf16_peak()
            tpeak = bf16[0] / 6.0
            src = "%s / 6 bf16 MMAs per fp32-grade product (bf16x3 split)" % bf16[1]
        else:
            tpeak = peaks.get("fp64_tflops", 40.0)
            src = ("cuBLAS fp64 GEMM (DMMA) measured in this run" if "fp64_tflops" in peaks
                   else "nominal B200 fp64 tensor rate")
        return {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": ("dense_tc_kernel (tcgen05.mma kind::f16, bf16x3 operands: 6 MMAs per product, TMEM accumulators)" if f32
                           else "dense_sample_kernel (mma.sync m8n8k4 f64 = DMMA)"),
                "peak_source": src, "algorithmic_flops_per_leapfrog_step": fl, "launch_ms": launch_ms,
                "hbm_gbs": achieved}
    streaming = w.get("kernel") == "stream" or w["ndims"] > 16
    if w["L"] <= 4:
        note = "HBM-bound point"
    elif streaming:
        note = "L=%d leapfrog steps per sample through the streaming kernel (DESIGN.md 3.1b)" % w["L"]
    else:
        note = "L=%d leapfrog steps per sample: instruction-issue bound, not HBM bound (DESIGN.md 3.1)" % w["L"]
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": "stream_sample_kernel" if streaming else "fused_sample_kernel",
                "algorithmic_bytes_per_launch": alg, "launch_ms": launch_ms, "note": note}
    if w["dist"] == "RoughWell" and not streaming and w["L"] > 4:
        # the pipe that does bound this kernel: fp64 instructions of the leapfrog loop per particle-dimension-step as
        # the kernel's SASS has them (15: one FMA each for the merged kick and the drift, 13 for
        # x/s1^2 - c sin(2 pi x / s2) -- two FMAs + one add of range reduction, f^2, eight polynomial FMAs, q * f, and
        # the FMA that adds x/s1^2; round 1 counted 18 before the kicks were merged and the reduction went to FMAs)
        clk = m.get("clk")
        fp64_inst = 15.0 * w["ndims"] * m["grads_all"] / world / (m["ms_max"] * 1e-3)
        fp64_peak = 148 * 64 * ((clk or {}).get("sm_mhz") or 1965.0) * 1e6
        roofline["fp64_pipe"] = {"achieved_inst_per_s": fp64_inst, "peak_inst_per_s": fp64_peak,
                                 "frac": fp64_inst / fp64_peak,
                                 "inst_per_dim_step": 15,
                                 "note": "fp64 instructions of the leapfrog loop only (transition, energies, Philox not "
                                         "counted); peak = 64 fp64 lanes per SM at the sampled SM clock, the rate "
                                         "tools/probe/fp64_probe.cu measures on B200 (2.0 warp-DFMA per cycle per SM, "
                                         "profiles/r2_variants_funnel_stash.txt); ncu pipe-active of the same kernel: 70 %"}
    return roofline


def measure_ess(ctx, m, cpu=False):
    """ESS/s (BASELINE metric iii): T recorded steps of every chain, per-GPU autocorrelation sums through the FFT kernel
    (csrc/autocorr_fft.cu: what misc/autocor.py:37-49 does), one all-reduce of float64[n_lags] over NCCL.  The chains are
    independent, so the cloud is walked in particle blocks whose recorded samples (ndims x T x block) fit in HBM."""
    torch = ctx.torch
    from mjhmc_b200 import parallel
    w = m["w"]
    T, n_lags, block = w["ess"]["T"], w["ess"]["n_lags"], w["ess"]["block"]
    for key in ("sampler", "dist", "X0", "V0"):
        m.pop(key, None)
    torch.cuda.empty_cache()
    n_total = m["n"]
    part = torch.zeros(n_lags, dtype=torch.float64, device=ctx.dev)
    t_s = t_a = 0.0
    done = 0
    blk = 0
    while done < n_total:
        nb = min(block, n_total - done)
        sampler, _, _, _ = make_sampler(dict(w, n=nb), ctx.rank * 64 + blk, n=nb, device_init=ctx.dev)
        sampler.sample_device(4)                                            # warm the block's launch path
        ctx.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        S = sampler.sample_device(T)
        e1.record()
        part += parallel.autocorr_partial(S, n_lags=n_lags, circular=True)
        e2.record()
        torch.cuda.synchronize()
        t_s += e0.elapsed_time(e1)
        t_a += e1.elapsed_time(e2)
        del S, sampler
        torch.cuda.empty_cache()
        done += nb
        blk += 1
    ctx.barrier()
    e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e3.record()
    if ctx.world > 1:
        ctx.dist.all_reduce(part)                                           # the statistics "gathered over NVLink"
    e4.record()
    ctx.barrier()
    t_s, t_a, t_r = ctx.max_f(t_s, t_a, e3.elapsed_time(e4))
    ac = part.double().cpu().numpy()
    ac = ac / ac[0]
    ess_chain = _ess_from_curve(ac, T)
    crossed = bool(np.any(ac[1:] < 0))
    ess = {"definition": "T / (1 + 2 sum_{tau>=1}^{first rho<0} rho_tau) on the circular fft_autocor curve (autocor.py:37-49); "
                         "the reference defines no ESS (its figure of merit is a fitted decay, search/objective.py:121-185)",
           "T": T, "n_lags": n_lags, "chains": n_total * ctx.world, "blocks_per_gpu": blk, "ess_per_chain": ess_chain,
           "sampling_ms": t_s, "autocorr_ms": t_a, "allreduce_ms": t_r, "autocorr_share_of_sampling": t_a / t_s,
           "ess_per_s": ess_chain * n_total * ctx.world / ((t_s + t_a + t_r) * 1e-3), "rho_1": float(ac[1]),
           "rho_last": float(ac[-1]), "rho_crosses_zero_inside_window": crossed,
           "first_negative_lag": int(np.argmax(ac < 0)) if crossed else None,
           "autocorr_kernel": "autocorr_fft_kernel (batched radix-4 forward FFTs in shared memory + one inverse transform)"}
    if cpu:
        ess["cpu"] = run_reference_ess(w)
    return ess


def measure_pcie(ctx, nbytes=256 << 20):
    """Plain device -> pinned host copy rate of this box, all ranks at once (the roof the e2e number sits on)."""
    torch = ctx.torch
    src = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    ctx.barrier()
    t = ctx.max_f(time.perf_counter() - t0)[0]
    return 4 * nbytes * ctx.world / t / 1e9


def measure_e2e(ctx, m, steps):
    """The same work through the public API with HOST buffers: state uploaded from pinned host memory, sample()
    returns a host numpy array, every step."""
    torch = ctx.torch
    from mjhmc_b200.samplers.hmc_state import HMCState
    w, sampler = m["w"], m["sampler"]
    Xh = torch.as_tensor(m["X0"]).pin_memory()                  # host inputs live in pinned memory
    Vh = torch.as_tensor(m["V0"]).pin_memory()
    e2e_steps = max(1, min(steps, 5))
    e2e_iters = w["iters"]

    def e2e_step():
        sampler.state = HMCState.from_buffers(sampler, Xh, Vh)   # H2D from pinned memory at the next launch
        return sampler.sample(e2e_iters)                         # D2H of (ndims, iters * n)

    for _ in range(2):                                           # warm the pinned staging buffers (not timed)
        res = e2e_step()
    ctx.barrier()
    g1 = sampler.grad_evals_executed
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    ctx.barrier()
    e2e_t = ctx.max_f(time.perf_counter() - t0)[0]
    ge = ctx.sum_i(sampler.grad_evals_executed - g1)[0]
    S = 4 if w.get("dtype") == "float32" else 8
    h2d = 2 * w["ndims"] * m["n"] * S                           # X and V; the empty FLF cache is cleared on the device
    d2h = res.nbytes
    pcie = measure_pcie(ctx)
    copy_s = e2e_steps * (h2d + d2h) * ctx.world / (pcie * 1e9)
    return {"value": int(ge) / e2e_t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": e2e_steps, "iterations_per_step": e2e_iters,
            "pcie_gbs": pcie, "pcie_note": "plain device->pinned-host copies of 256 MB, all %d ranks at once, same run" % ctx.world,
            "copy_time_share": copy_s / e2e_t,
            "roof_value": int(ge) / copy_s if copy_s > 0 else None}


# the workloads whose device-timed value + roofline ride along in the default line (BASELINE configs 2-5 and the two
# north_star targets: >= 0.8 of the HBM roofline on fused Gaussian / RoughWell leapfrog, >= 50 % tensor pipe on PoT)
SECONDARY = ["roughwell2d_control", "gauss10d_control_L1_stream", "gauss100d_diag_control_L1",
             "testgauss2d_control_L1_stream", "roughwell2d_control_L1_stream", "roughwell10d_control_L1_stream",
             "gauss100d_diag_mjhmc", "gauss100d_mjhmc", "gauss100d_mjhmc_f32", "pot100d_mjhmc", "pot100d_mjhmc_f32",
             "funnel10d_cthmc", "funnel10d_cthmc_ess"]


def summarise(ctx, m, peaks):
    w = m["w"]
    return {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms_max"] / m["steps"], "steps": m["steps"],
            "warmup": m["warmup"], "dtype": "f32" if w.get("dtype") == "float32" else "f64",
            "particles_per_gpu": m["n"], "iterations_per_step": w["iters"], "sampler": w["sampler"],
            "hyper_parameters": {"epsilon": w["epsilon"], "beta": w["beta"], "L": w["L"], "source": w["source"]},
            "gpu_launches": 2 * m["launches_all"], "roofline": roofline_of(ctx, m, peaks),
            "particle_iterations_per_s": m["n"] * ctx.world * w["iters"] * m["steps"] / (m["ms_max"] * 1e-3)}


def run_b200(args, w):
    ctx = Ctx()
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    strong = args.scaling == "strong"
    n_main = w["n"] // world if strong else w["n"]

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_reference(w, steps=2, warmup=1)

    peaks = measured_tensor_peaks(ctx)
    clocks = ClockSampler(ctx.local_rank) if rank == 0 else None
    m = measure_device(ctx, args.workload, w, args.steps, args.warmup, n=n_main, host_init=True, clocks=clocks)
    e2e = measure_e2e(ctx, m, args.steps)
    main_summary = summarise(ctx, m, peaks)
    ess = measure_ess(ctx, m, cpu=(rank == 0 and world == 1 and not args.no_cpu_baseline)) if w.get("ess") else None
    clk = m["clk"]

    # ---- strong scaling of the same workload (the cloud of ONE GPU's worth split over the ranks), N > 1 only
    strong_line = None
    if world > 1 and not strong and not args.no_secondary:
        for key in ("sampler", "dist", "X0", "V0"):
            m.pop(key, None)
        torch.cuda.empty_cache()
        ms_ = measure_device(ctx, args.workload, w, args.steps, args.warmup, n=w["n"] // world)
        strong_line = {"value": ms_["value"], "unit": UNIT, "particles_total": (w["n"] // world) * world,
                       "particles_per_gpu": w["n"] // world, "ms_per_step": ms_["ms_max"] / ms_["steps"],
                       "note": "strong scaling: the single-GPU cloud split over the %d ranks" % world}
        del ms_

    # ---- the other BASELINE configs and the north_star target points, device-timed (few steps each)
    workloads = {}
    if args.workload == DEFAULT_WORKLOAD and not strong and not args.no_secondary:
        for key in ("sampler", "dist", "X0", "V0"):
            m.pop(key, None)
        for name in SECONDARY:
            torch.cuda.empty_cache()
            w2 = WORKLOADS[name]
            try:
                # One-iteration launches of the discrete samplers: the batch-wide R coin (markov_jump_hmc.py:138) fires in
                # p_r = 5.3 % of the launches at beta = 0.1, and such a launch refreshes every momentum (Box-Muller for the
                # whole cloud: twice the time).  A window of 5 launches holds 0 or 1 of them (0 % or 20 %); 200 launches
                # (0.13 s) hold the long-run share.
                m2 = measure_device(ctx, name, w2, steps=200 if w2["iters"] == 1 else 5, warmup=3)
                workloads[name] = summarise(ctx, m2, peaks)
                if w2["iters"] == 1 and w2["sampler"] == "ControlHMC":
                    workloads[name]["steps_note"] = ("200 one-iteration launches: the batch-wide momentum refresh of "
                                                     "ControlHMC fires in about p_r = 5 % of them and doubles the launch")
                if w2.get("ess"):
                    workloads[name]["ess"] = measure_ess(ctx, m2)
            except Exception as exc:   # noqa: BLE001 -- a secondary workload must not lose the headline line
                workloads[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            m2 = None

    if rank == 0:
        iters = w["iters"]
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_max"] / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32" if w.get("dtype") == "float32" else "f64", "data": "synthetic",
            "config": {"workload": "%s: %s %d-d, %d particles per GPU, %s eps=%g beta=%g L=%d (%s); %d sampling "
                                   "iterations per step in one fused launch" % (
                                       args.workload, w["dist"], w["ndims"], n_main, w["sampler"], w["epsilon"], w["beta"],
                                       w["L"], w["source"], iters),
                       "particles_per_gpu": n_main, "iterations_per_step": iters,
                       "l2": "flushed with a 256 MB fill between timed steps", "rng": "philox4x32-10",
                       "parallelism": "particle shards, dp%d" % world},
            "roofline": main_summary["roofline"],
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": 2 * m["launches_all"],
            "gpu_launches_note": "per step one sampler kernel and the one-warp counter fold (mjhmc_counters_read); "
                                 "torch fills (L2 flush, counter reset copy) not counted",
            "ess": ess,
            "clocks": clk,
            "measured_tensor_peaks": peaks,
            "wall_s_timed_region": m["t_wall"],
            "grad_evals_executed": m["grads_all"],
            "dEdX_count_delta": m["grads_ref_all"],
            "dEdX_count_rate": m["grads_ref_all"] / (m["ms_max"] * 1e-3),
            "particle_iterations_per_s": n_main * world * iters * args.steps / (m["ms_max"] * 1e-3),
        }
        if strong_line:
            line["strong_scaling"] = strong_line
        if workloads:
            line["workloads"] = workloads
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads / strong-scaling leg")
    ap.add_argument("--ess-block", type=int, default=None, help="particles per block of the ESS leg (developer knob)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the workload's particle count is the TOTAL, split over the ranks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    w = WORKLOADS[args.workload]
    if args.ess_block and w.get("ess"):
        w = dict(w, ess=dict(w["ess"], block=args.ess_block))

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        steps = max(1, min(args.steps, 3))
        warm = min(args.warmup, 1)
        r = run_reference(w, steps=steps, warmup=warm)
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s %d-d, %s eps=%g beta=%g L=%d; bounded sample of %d particles; %d sampling "
                                   "iterations per step" % (args.workload, w["dist"], w["ndims"], w["sampler"],
                                                            w["epsilon"], w["beta"], w["L"], r["n_particles"], w["iters"])},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return
    run_b200(args, w)


if __name__ == "__main__":
    main()
