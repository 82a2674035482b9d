"""Particle sharding over the GPUs of one box (one process per GPU, torch.distributed).

Particles are independent chains, so the data path has NO collective: rank r owns the
contiguous block [r*N/G, (r+1)*N/G) of the (ndims, N) arrays and runs the fused sampler kernel on
it with ``particle_offset`` = its first global index, which keys the Philox stream by the GLOBAL
particle index -- results do not depend on G.  NCCL (gloo in the CPU tests) is used only for
  * the int64 operator / evaluation counters          -> all_reduce(SUM)
  * autocorrelation partial sums (a mean over particles is a sum)  -> all_reduce(SUM) of float64[n_lags]
  * samples, when the caller wants the full array     -> all_gather
The batch-wide R coin of the discrete samplers needs no exchange: it is drawn from a
particle-independent Philox counter, so every rank computes the same coin.

Samplers built with ``sharded=True`` (or ``sharded=<process group>``) also take the reference's two other
batch-global couplings over the whole cloud, so a sharded run equals the single-GPU run:
  * MarkovJumpHMC's infinite-rate back-off (markov_jump_hmc.py:376-389)   -> all_reduce(MIN) of the first failing
    iteration after every launch: all ranks replay, count and retry at the same iteration
  * dwell-time resampling (markov_jump_hmc.py:321-328)                    -> all_gather of the per-(iteration, rank)
    dwell sums + one broadcast of the sorted uniforms (resample_plan), then the local resampling kernel
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _device, _lib

COUNTER_NAMES = ("l_count", "f_count", "fl_count", "r_count", "E_count", "dEdX_count")


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def resolve_group(sharded):
    """``sharded=True`` -> the default (WORLD) group; a process group is taken as is."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("sharded=... needs an initialised torch.distributed process group")
    return dist.group.WORLD if sharded is True else sharded


def allreduce_min(value, group=None):
    """min over the ranks of one int64 (the first failing iteration of a launch, INT64_MAX = none)."""
    rank, ws = world(group)
    if ws == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def allreduce_sum_int(value, group=None):
    rank, ws = world(group)
    if ws == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def resample_plan(seg_sums, m_out, group=None, uniforms=None):
    """Global offsets for dwell-time resampling of a sharded cloud (markov_jump_hmc.py:321-328).

    seg_sums: (n_iter,) float64, dwell sum of the LOCAL particles per iteration.  The reference's cumulative sum
    runs over the flat order (iteration-major, particle-minor over the whole cloud), in which the local particles
    of iteration `it` form one contiguous segment.  Returns a dict with
      r       (m_out,) sorted uniforms * total dwell time -- np.sort(np.random.random(m_out)) drawn on rank 0 and
              broadcast (or `uniforms`, already sorted, for tests), identical on every rank
      bounds  (2 n_iter,) [A_0, B_0, A_1, B_1, ...]: this rank's segment `it` covers cumulative times [A_it, B_it)
      gaps    (n_iter,) A_it - B_{it-1}: dwell mass owned by other ranks between two consecutive local segments
      total   total dwell time of the cloud
    """
    rank, ws = world(group)
    seg = seg_sums.detach().double()
    n_iter = seg.numel()
    if ws > 1:
        dev = _comm_device(group)
        mine = seg.to(dev).contiguous()
        parts = [torch.empty_like(mine) for _ in range(ws)]
        dist.all_gather(parts, mine, group=group)
        table = torch.stack(parts, dim=1).cpu().numpy()          # (n_iter, ws)
    else:
        table = seg.cpu().numpy().reshape(n_iter, 1)
    incl = np.cumsum(table.reshape(-1)).reshape(n_iter, ws)      # inclusive prefix over (iteration, rank) segments
    total = float(incl[-1, -1]) if incl.size else 0.0
    excl = np.concatenate(([0.0], incl.reshape(-1)[:-1])).reshape(n_iter, ws)
    A, B = excl[:, rank], incl[:, rank]
    bounds = np.stack((A, B), axis=1).reshape(-1)
    gaps = A - np.concatenate(([0.0], B[:-1]))
    if uniforms is None:
        if ws > 1:
            dev = _comm_device(group)
            u = torch.empty(m_out, dtype=torch.float64, device=dev)
            if rank == 0:
                u.copy_(torch.as_tensor(np.sort(np.random.random(m_out))))
            dist.broadcast(u, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uniforms = u.cpu().numpy()
        else:
            uniforms = np.sort(np.random.random(m_out))
    return dict(r=np.asarray(uniforms) * total, bounds=bounds, gaps=gaps, total=total)


def allgather_resampled(local_cols, columns, m_out, group=None):
    """Assembles the reference's resampled array (ndims, m_out) from the per-rank parts a sharded
    ``sample()`` returns: local_cols (ndims, k) numpy and their positions ``sampler.resample_columns``."""
    rank, ws = world(group)
    d = local_cols.shape[0]
    out = np.zeros((d, m_out))
    if ws == 1:
        out[:, columns] = local_cols
        return out
    payload = [None] * ws
    dist.all_gather_object(payload, (np.asarray(columns), np.asarray(local_cols)), group=group)
    for cols, vals in payload:
        out[:, cols] = vals
    return out


def shard_bounds(n_global, rank, world_size):
    """Contiguous particle block of `rank`: [lo, hi)."""
    lo = (n_global * rank) // world_size
    hi = (n_global * (rank + 1)) // world_size
    return lo, hi


def shard_columns(X, rank, world_size):
    """The (ndims, n_local) block of a global (ndims, N) array owned by `rank` (a copy)."""
    lo, hi = shard_bounds(X.shape[1], rank, world_size)
    return np.ascontiguousarray(X[:, lo:hi]), lo


def _comm_device(group=None):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def local_counters(sampler):
    d = sampler.distribution
    return [int(sampler.l_count), int(sampler.f_count), int(sampler.fl_count), int(sampler.r_count),
            int(d.E_count), int(d.dEdX_count)]


def allreduce_counters(sampler_or_list, group=None):
    """Global (whole particle cloud) counters as a dict; every rank gets the same ints."""
    vals = sampler_or_list if isinstance(sampler_or_list, (list, tuple)) else local_counters(sampler_or_list)
    rank, ws = world(group)
    if ws > 1:
        t = torch.tensor(vals, dtype=torch.int64, device=_comm_device(group))
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        vals = [int(v) for v in t.tolist()]
    return dict(zip(COUNTER_NAMES, vals))


def allgather_samples(S, group=None):
    """Local samples (ndims, n_iter, n_local) -> global (ndims, n_iter, N), particle blocks in rank order."""
    rank, ws = world(group)
    if ws == 1:
        return S
    dev = _comm_device(group)
    S = S.to(dev).contiguous()
    sizes = torch.zeros(ws, dtype=torch.int64, device=dev)
    sizes[rank] = S.shape[2]
    dist.all_reduce(sizes, group=group)
    sizes = [int(n) for n in sizes.tolist()]
    n_max = max(sizes)
    if S.shape[2] < n_max:                                   # all_gather wants equal shapes: pad the short shards
        S = torch.cat([S, S.new_zeros((S.shape[0], S.shape[1], n_max - S.shape[2]))], dim=2)
    parts = [torch.empty_like(S) for _ in range(ws)]
    dist.all_gather(parts, S.contiguous(), group=group)
    return torch.cat([p[:, :, :n] for p, n in zip(parts, sizes)], dim=2)


def autocorr_partial(S, n_lags=None, circular=True, method="auto"):
    """Un-normalised autocorrelation sums of the LOCAL particles on the GPU (K7):
    ac[tau] = sum_{k,i,t} x[k,t,i] x[k,(t+tau) mod T,i]  (circular; linear: only t + tau < T);
    S is the device tensor (ndims, T, n_local).  method: "fft" (the reference's route, autocor.py:37-49: T a power of
    two up to 4096), "direct" (products, any T and the linear window) or "auto" (fft where it applies)."""
    lib = _lib.load()
    d, T, n = S.shape
    n_lags = T if n_lags is None else int(n_lags)
    ac = torch.zeros(n_lags, dtype=torch.float64, device=S.device)
    code, stream = _device.dtype_code(S.dtype), _device.stream_ptr(S.device)
    fft_bytes = int(lib.mjhmc_autocorr_fft_scratch_bytes(T)) if (circular and n_lags <= T) else -1
    if method == "fft" and fft_bytes < 0:
        raise ValueError("method='fft' needs the circular sums of a power-of-two T in [16, 4096]")
    if method != "direct" and fft_bytes >= 0:
        scratch = torch.empty(fft_bytes, dtype=torch.uint8, device=S.device)
        _lib.check(lib.mjhmc_autocorr_fft(code, d, _device.ptr(S), S.stride(0), S.stride(1), n, T, n_lags, _device.ptr(ac),
                                          _device.ptr(scratch), stream), "autocorr_fft")
        return ac
    _lib.check(lib.mjhmc_autocorr(code, d, _device.ptr(S), S.stride(0), S.stride(1), n, T, n_lags, 1 if circular else 0,
                                  _device.ptr(ac), stream), "autocorr")
    return ac


def autocorrelation(S, n_lags=None, group=None, partial=None):
    """fft_autocor of the reference (misc/autocor.py:37-49) over the WHOLE sharded cloud:
    per-GPU partial sums, one all_reduce of float64[n_lags], normalised by lag 0."""
    ac = autocorr_partial(S, n_lags) if partial is None else partial
    rank, ws = world(group)
    if ws > 1:
        ac = ac.to(_comm_device(group))
        dist.all_reduce(ac, op=dist.ReduceOp.SUM, group=group)
    ac = ac.double().cpu().numpy()
    return ac / ac[0]


def effective_sample_size(ac):
    """ESS = T / (1 + 2 sum_{tau >= 1}^{first rho < 0} rho_tau) on the fft_autocor curve.
    The reference defines no ESS (its figure of merit is a fitted decay, search/objective.py:121-185);
    this is the definition the build reports."""
    T = len(ac)
    s = 0.0
    for tau in range(1, T):
        if ac[tau] < 0:
            break
        s += ac[tau]
    return T / (1.0 + 2.0 * s)
