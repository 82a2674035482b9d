"""Search objective (reference: mjhmc/search/objective.py).

The score of a hyper-parameter candidate is Re(r) of the complex exponential exp(r t) that fits the
autocorrelation curve best: ``exp(a t) cos(b t)`` fitted to (normalised gradient evaluations, autocorrelation),
``obj_func`` returns ``a`` (objective.py:34-47, 121-185).  The autocorrelation comes from
``mjhmc_b200.misc.autocor.calculate_autocorrelation`` (fused sampling launches + device autocorrelation sums);
the two-parameter fit itself is host arithmetic on a few hundred points:

  * ``estimate_params``, ``curve_fn``, ``fit`` follow objective.py:190-223 and :120-135 line for line (pinned by
    tests/golden/objective_reference.npz, outputs of the reference's own function source);
  * ``tf_fit`` of the reference (:137-183) builds a TensorFlow graph to run 1e4 Adam steps on the squared error;
    here the same optimiser (TF1 AdamOptimizer update rule, learning rate 0.01, fp64) runs in numpy with the
    analytic gradient.  TensorFlow is absent from this image, so this function is *parity unpinned*.

Plotting (plot_fit, plot_search_ac) and the trace pickles are out of scope; ``save_trace`` is kept behind
``SAVE_TRACE = False``.
"""
import os
import pickle
import time

import numpy as np
from scipy.optimize import curve_fit

from ..misc.autocor import calculate_autocorrelation

# gradient-evaluation budgets per distribution class name (objective.py:20-27)
grad_evals = {
    'Gaussian': int(5E4),
    'RoughWell': int(5E5),
    'MultimodalGaussian': int(2E5),
    'ProductOfT': int(1E5),
    'Funnel': int(1E5),
    'SparseImageCode': int(5E5),
    'TestGaussian': int(5E4),          # not in the reference's table: used by the tests
}

debug = False
SAVE_TRACE = False
TRACE_PATH = os.path.expanduser('~/data/mjhmc/autocor_traces')


def obj_func(sampler, distr, job_id, **kwargs):
    """Scores the performance of sampler (a class) on distribution given parameters (objective.py:34-47).
    Returns the fitted exponential coefficient (more negative = faster decay of the autocorrelation)."""
    cos_coef, normed_n_grad_evals, exp_coef, autocor, kwargs = obj_func_helper(sampler, distr, True, kwargs)
    return exp_coef


def obj_func_helper(sampler, distr, unpack, kwargs, overrides=None):
    """objective.py:49-85.  ``overrides`` (B200 extension) is applied after the reference's defaults, which
    themselves take precedence over ``kwargs`` exactly as in the reference (``kwargs.update(default_args)``)."""
    num_target_grad_evals = grad_evals[type(distr).__name__]
    default_args = {
        "num_grad_steps": num_target_grad_evals,
        "sample_steps": 1,
        "num_steps": None,
        "half_window": True,
        "use_cached_var": True,
    }
    if unpack:
        kwargs = unpack_params(kwargs)
    if sampler.__name__ == 'MarkovJumpHMC':
        default_args["resample"] = False
    kwargs.update(default_args)
    if overrides:
        kwargs.update(overrides)
        num_target_grad_evals = kwargs.get("num_grad_steps") or num_target_grad_evals

    print("Calculating autocorrelation for {} grad evals".format(num_target_grad_evals))
    autocor, _, n_grad_evals = calculate_autocorrelation(sampler, distr, **kwargs)

    # necessary to keep curve_fit from borking: THIS IS VERY IMPORTANT (objective.py:75-76)
    normed_n_grad_evals = n_grad_evals / (0.5 * num_target_grad_evals)
    print("Fitting curve")
    exp_coef, cos_coef = tf_fit(normed_n_grad_evals.copy(), autocor.copy())

    if SAVE_TRACE:
        formatted_time = time.strftime("%Y%m%d-%H%M%S")
        trace_name = '{}_{}'.format(type(distr).__name__, formatted_time)
        save_trace(normed_n_grad_evals, autocor, exp_coef, cos_coef, trace_name)
    return cos_coef, normed_n_grad_evals, exp_coef, autocor, kwargs


def save_trace(t_data, y_data, tf_ec, tf_cc, trace_name):
    """Save the trace for later inspection (objective.py:87-97)."""
    os.makedirs(TRACE_PATH, exist_ok=True)
    with open('{}/{}.pkl'.format(TRACE_PATH, trace_name), 'wb') as pkl_file:
        pickle.dump({'grad_evals': t_data, 'autocor': y_data, 'tf_exp_coeff': tf_ec, 'tf_cos_coeff': tf_cc}, pkl_file)


def min_idx(ac_df, target):
    """First 'num grad' at which the autocorrelation drops below target (objective.py:99-107)."""
    ac_df.index = ac_df['num grad']
    ac_trunc = ac_df.loc[:, 'autocorrelation'] < target
    small_ac = ac_trunc[ac_trunc]
    if len(small_ac) != 0:
        return small_ac.index[0]
    return None


def unpack_params(params):
    """Spearmint passes params as 1x1 arrays (objective.py:109-118); scalars pass through."""
    unpacked_params = {}
    for key, item in params.items():
        try:
            unpacked_params[key] = item[0]
        except (TypeError, IndexError):
            unpacked_params[key] = item
    return unpacked_params


def fit(t_data, y_data):
    """Fit a complex exponential to y_data with scipy's curve_fit (objective.py:120-135)."""
    # very fast way to check for nan
    if not np.isnan(np.sum(y_data)):
        p_0 = None
        opt_params = curve_fit(curve_fn, t_data, y_data, p0=p_0, maxfev=1000)[0]
        return opt_params
    return 1E3, 0


def tf_fit(t_data, y_data, n_steps=int(1e4), learning_rate=0.01):
    """Fit exp(a t) cos(b t) by Adam on the summed squared error, started from estimate_params; returns the
    parameters of the smallest loss seen (objective.py:137-183, TensorFlow graph replaced by numpy)."""
    t = np.asarray(t_data, dtype=np.float64).squeeze()
    y = np.asarray(y_data, dtype=np.float64).squeeze()
    a, b = (float(v) for v in estimate_params(t, y))
    beta1, beta2, eps_hat = 0.9, 0.999, 1e-8          # tf.train.AdamOptimizer defaults
    m = np.zeros(2)
    v = np.zeros(2)
    best = (np.inf, a, b)
    for step in range(1, int(n_steps) + 1):
        e = np.exp(a * t)
        c = e * np.cos(b * t)
        r = y - c
        loss = float(np.sum(r * r))
        if loss < best[0]:                              # np.argmin(losses): first minimum
            best = (loss, a, b)
        if not np.isfinite(loss):
            break
        g = np.array([np.sum(-2.0 * r * t * c), np.sum(2.0 * r * e * t * np.sin(b * t))])
        m = beta1 * m + (1.0 - beta1) * g
        v = beta2 * v + (1.0 - beta2) * g * g
        lr_t = learning_rate * np.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
        a, b = np.array([a, b]) - lr_t * m / (np.sqrt(v) + eps_hat)
        a, b = float(a), float(b)
    return best[1], best[2]


def estimate_params(t_data, y_data):
    """Initial guess of (exp_coef, cos_coef) from the slope at 0 and at the first zero crossing (objective.py:190-218)."""
    dydt_0 = (y_data[0] - y_data[1]) / float(t_data[0] - t_data[1])
    exp_coef = dydt_0

    zero_idx = np.where(y_data == 0)[0]
    zero_crossings = np.where(np.diff(np.sign(y_data)) != 0)[0]
    if len(zero_idx) != 0 and len(zero_crossings) != 0:
        first_zero_idx = min(zero_crossings[0], zero_idx[0])
    elif len(zero_crossings) != 0:
        first_zero_idx = zero_crossings[0]
    elif len(zero_idx) != 0:
        raise ValueError("zeros should be found if and only if a zero crossing is found")
    else:
        return exp_coef, 0

    dydt_zero_cross = (y_data[first_zero_idx] - y_data[first_zero_idx + 1]) / \
        float(t_data[first_zero_idx] - t_data[first_zero_idx + 1])
    cos_coef = - (dydt_zero_cross / np.exp(exp_coef * (first_zero_idx + 0.5)))
    return exp_coef, cos_coef


def curve_fn(n, a, b):
    return np.exp(a * n) * np.cos(b * n)
