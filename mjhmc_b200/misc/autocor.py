"""Sample generation and autocorrelation (reference: mjhmc/misc/autocor.py).

Same entry points as the reference -- ``calculate_autocorrelation``, ``generate_samples``,
``fft_autocor``, ``autocorrelation``, ``slow_autocorrelation`` -- but

  * ``generate_samples`` draws all steps in fused launches instead of one ``sample(1)`` call plus a
    host-visible counter read per step (autocor.py:245-248); the per-step evaluation trace
    (``E_count / n_batch``, ``dEdX_count / n_batch`` after every step) is reconstructed exactly from
    the per-iteration operator choices the kernel records,
  * the autocorrelation sums run on the GPU (csrc/analysis.cu: autocorr_kernel) and, when the
    particles are sharded, are all-reduced over NCCL (mjhmc_b200/parallel.py).
"""
from time import time

import numpy as np
import torch

from .. import _lib, parallel


def calculate_autocorrelation(sampler, distribution, num_steps=None, num_grad_steps=None,
                              sample_steps=1, half_window=False, use_cached_var=False, **kwargs):
    """just a helper function (autocor.py:11-35)"""
    print("Now generating samples...")
    start_time = time()
    samples, e_evals, grad_evals = generate_samples(sampler, distribution.reset(), num_steps, num_grad_steps,
                                                    return_device=True, **kwargs)
    print("Took {} seconds".format(time() - start_time))

    cached_var = None
    if use_cached_var:
        print("Using cached variance")
        try:
            _, emc_var_estimate, true_var_estimate, _ = distribution.load_cache()
            cached_var = emc_var_estimate if sampler.__name__ == "MarkovJumpHMC" else true_var_estimate
        except FileNotFoundError:
            # the reference writes the cache inside every Distribution() constructor (distributions.py:96-149); here
            # that burn-in is opt-in, so the file may not exist.  The default (fft) estimate below
            # normalises by lag 0 and never reads the cached variance (autocor.py:107-111).
            print("Warning: no fair-initialisation cache for this distribution "
                  "(distribution.cached_init_X() generates it); continuing without a cached variance")

    print("Calculating autocorrelation...")
    return autocorrelation(samples, e_evals, grad_evals, half_window, cached_var=cached_var)


def _as_device_tnk(samples):
    """(n_dims, n_batch, n_samples) numpy  or  device tensor (n_dims, n_samples, n_batch) -> the latter."""
    if isinstance(samples, torch.Tensor):
        return samples
    from .. import _device
    dev = _device.require_cuda()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(samples).transpose(0, 2, 1)), device=dev)


def fft_autocor(samples):
    """Autocorrelation by the cross-correlation theorem in the reference (autocor.py:37-49: circular, no
    mean subtraction, normalised by lag 0); here the same sums as direct products on the GPU.

    samples: numpy [n_dims, n_batch, n_samples] (reference layout) or the device tensor
    (n_dims, n_samples, n_batch) that ``sample_device`` returns.  Returns autocor [n_samples]."""
    return parallel.autocorrelation(_as_device_tnk(samples))


def autocorrelation(samples, e_evals, grad_evals, half_window=True, normalize=True, cached_var=None,
                    brute_force=False, use_tf=False):
    """autocor.py:52-117.  brute_force=True is the linear-window estimate of the Theano / TensorFlow ops
    (autocor.py:121-174) computed by the device kernel in linear mode."""
    S = _as_device_tnk(samples)
    n_dims, n_samples, n_batch = S.shape
    if brute_force:
        max_t = int(n_samples / 2) - 1 if half_window else n_samples - 1
        sums = parallel.autocorr_partial(S, n_lags=max_t, circular=False)
        rank, ws = parallel.world()
        if ws > 1:
            import torch.distributed as dist
            dist.all_reduce(sums)
        sums = sums.double().cpu().numpy()
        n_batch_global = n_batch
        if ws > 1:
            t = torch.tensor([n_batch], dtype=torch.int64, device=sums.device if isinstance(sums, torch.Tensor) else S.device)
            import torch.distributed as dist
            dist.all_reduce(t)
            n_batch_global = int(t.item())
        counts = n_dims * n_batch_global * (n_samples - np.arange(max_t))
        c = sums / counts                                 # c[t] = mean(x[:, :, :-t] * x[:, :, t:])
        var = c[0] if cached_var is None else cached_var  # variance given assumption of *zero mean*
        ac_squeeze = c[1:]
        if normalize:
            autocor = np.vstack((1., (ac_squeeze / var).reshape(-1, 1)))
        else:
            autocor = np.vstack((var, ac_squeeze.reshape(-1, 1)))
        if half_window:
            e_evals = e_evals[:int(n_samples / 2) - 1]
            grad_evals = grad_evals[:int(n_samples / 2) - 1]
        else:
            e_evals = e_evals[:-1]
            grad_evals = grad_evals[:-1]
    else:
        autocor = fft_autocor(S)
        print("Warning: not using cached emc variance!!")
        assert autocor.shape == e_evals.shape
        assert e_evals.shape == grad_evals.shape
    return autocor, e_evals, grad_evals


def slow_autocorrelation(samples, e_evals, grad_evals, half_window=False):
    """autocor.py:177-211: c[t] = mean(x[:, :, :-t] x[:, :, t:]) for t < T-1 (or T/2-1), c[0] = mean(x^2)."""
    S = _as_device_tnk(samples)
    n_dims, T, n_batch = S.shape
    n_lags = (T // 2) - 1 if half_window else T - 1
    sums = parallel.autocorr_partial(S, n_lags=n_lags, circular=False).double().cpu().numpy()
    c = sums / (n_dims * n_batch * (T - np.arange(n_lags)))
    return c / c[0], e_evals, grad_evals


def generate_samples(sampler, distribution, num_steps=None, num_grad_steps=None, return_device=False,
                     chunk=256, **kwargs):
    """ Generate samples *without* using a dataframe (autocor.py:213-261)

    Args:
       sampler: sampler class
       distribution: distribution object
       num_steps: number of desired steps - optional
       num_grad_steps: number of desired grad steps - optional
       return_device: (B200) return the samples as the device tensor (n_dims, n_samples, n_batch)

    Returns:
       (samples - [n_dims, n_batch, n_samples]
        e_evals - [n_samples]
        grad_evals - [n_samples])
    """
    # ridiculous assert to make sure only one of them is ever None
    assert (((num_steps is None) and (num_grad_steps is not None)) or
            (num_steps is not None) and (num_grad_steps is None))
    smp = sampler(distribution=distribution, **kwargs)
    # fudge factor because grad per sampler step is only approximate
    num_steps = num_steps or int(num_grad_steps // smp.grad_per_sample_step) + 100
    n_dims, n_batch = distribution.ndims, distribution.nbatch

    # reset counters (the reference also re-draws Xinit here; the sampler keeps its own state)
    distribution.reset()

    if getattr(smp, "resample", False) or not smp._engine.fused:
        # the reference's literal loop: sample(1) per step (Q25: with resample=True that is two iterations
        # plus a particle-scrambling resample per step -- drivers therefore pass resample=False)
        samples = np.zeros((n_dims, n_batch, num_steps))
        grad_evals, e_evals = np.zeros(num_steps), np.zeros(num_steps)
        for t_idx in range(num_steps):
            samples[:, :, t_idx] = smp.sample(1)
            grad_evals[t_idx] = distribution.dEdX_count / float(n_batch)
            e_evals[t_idx] = distribution.E_count / float(n_batch)
            if (num_grad_steps is not None) and grad_evals[t_idx] >= num_grad_steps:
                return _finish(samples[:, :, :t_idx + 1], e_evals[:t_idx + 1], grad_evals[:t_idx + 1], return_device)
        if num_grad_steps is not None:
            assert grad_evals[-1] >= num_grad_steps
            sel = grad_evals <= num_grad_steps
            return _finish(samples[:, :, sel], e_evals[sel], grad_evals[sel], return_device)
        return _finish(samples, e_evals, grad_evals, return_device)

    # fused path: chunks of iterations in one launch each, per-step counters from the recorded choices
    parts, e_tr, g_tr = [], [], []
    done, stop = 0, None
    while done < num_steps and stop is None:
        m = min(chunk, num_steps - done)
        snap = smp._snapshot()
        try:
            S, e_c, g_c = _advance_with_trace(smp, distribution, m)
        except _BackoffInChunk:
            # a MarkovJumpHMC infinite-rate back-off (markov_jump_hmc.py:376-389) fired inside the chunk: its extra
            # attempt breaks the one-iteration-per-step bookkeeping of the trace, so this chunk is redone with the
            # reference's literal loop (one step + one counter read at a time, autocor.py:245-248)
            smp._restore(snap)
            S, e_c, g_c = _advance_stepwise(smp, distribution, m)
        if num_grad_steps is not None:
            hit = np.nonzero(g_c >= num_grad_steps)[0]
            if len(hit) and hit[0] < m - 1:
                # the budget was reached inside the chunk: replay exactly hit[0]+1 iterations so the sampler
                # and the counters stop where the reference's loop stops (streams are counter based)
                smp._restore(snap)
                S, e_c, g_c = _advance_with_trace(smp, distribution, int(hit[0]) + 1)
                stop = True
            elif len(hit):
                stop = True
        parts.append(S); e_tr.append(e_c); g_tr.append(g_c)
        done += S.shape[1]
    S = torch.cat(parts, dim=1) if len(parts) > 1 else parts[0]
    e_evals, grad_evals = np.concatenate(e_tr), np.concatenate(g_tr)
    if num_grad_steps is not None and stop is None:
        assert grad_evals[-1] >= num_grad_steps
        sel = grad_evals <= num_grad_steps
        S = S[:, torch.as_tensor(np.nonzero(sel)[0], device=S.device), :]
        e_evals, grad_evals = e_evals[sel], grad_evals[sel]
    if return_device:
        return S, e_evals, grad_evals
    return smp._d2h(S.permute(0, 2, 1).contiguous()).astype(np.float64, copy=False), e_evals, grad_evals


def _finish(samples, e_evals, grad_evals, return_device):
    if return_device:
        return _as_device_tnk(samples), e_evals, grad_evals
    return samples, e_evals, grad_evals


class _BackoffInChunk(Exception):
    pass


def _advance_stepwise(smp, distribution, m):
    """m iterations, one launch and one counter read per step (the reference's loop, autocor.py:245-248)."""
    n = float(distribution.nbatch)
    parts, e_c, g_c = [], np.zeros(m), np.zeros(m)
    for t in range(m):
        parts.append(smp._advance(1)[0])
        g_c[t] = distribution.dEdX_count / n
        e_c[t] = distribution.E_count / n
    return torch.cat(parts, dim=1), e_c, g_c


def _advance_with_trace(smp, distribution, m):
    """m iterations in one launch; returns the samples and E_count/n, dEdX_count/n after every iteration.

    Discrete and continuous-time samplers evaluate N energies and L*N gradients per iteration.  MarkovJumpHMC
    additionally evaluates the FLF state of every particle whose cache is inactive at the start of the
    iteration (hmc_state.py:114-116), i.e. whose previous move was F or R: that count is read off the
    choices recorded by the kernel (markov_jump_hmc.py:409-410)."""
    n = float(distribution.nbatch)
    E0, G0 = distribution.E_count, distribution.dEdX_count
    a0 = smp._attempt
    L = smp.num_leapfrog_steps
    mj = smp._sampler_code == _lib.SAMPLER_MARKOV_JUMP
    uncached0 = None
    if mj:
        uncached0 = int((smp._engine.ca[smp._engine.cur] & 1).eq(0).sum().item())
    S, _, choice = smp._advance(m, want_choice=mj)
    if smp._attempt - a0 != m:
        raise _BackoffInChunk()
    per_iter = np.full(m, distribution.nbatch, dtype=np.int64)
    if mj:
        moved_off_cache = (choice != 0).sum(dim=1).cpu().numpy().astype(np.int64)      # F or R moves per iteration
        per_iter += np.concatenate(([uncached0], moved_off_cache[:-1]))
    e_c = (E0 + np.cumsum(per_iter)) / n
    g_c = (G0 + np.cumsum(per_iter * L)) / n
    assert distribution.E_count == E0 + per_iter.sum() and distribution.dEdX_count == G0 + (per_iter * L).sum()
    return S, e_c, g_c
