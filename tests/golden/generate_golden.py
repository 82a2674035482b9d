"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/generate_golden.py

It imports the reference under Python 3 with the two shims of SURVEY.md 8(c)
(``builtins.xrange = range``; distribution subclasses whose ``init_X`` calls
``gen_init_X`` so the fair-initialisation burn-in cache is bypassed -- the same
bypass MultimodalGaussian uses, distributions.py:338-342) and writes

  seeded_runs.json     known answers of ``sample(10)`` after ``np.random.seed(1)``
                       (counters, sums) for the five sampler classes, plus the
                       infinite-rate back-off known answer
  inject_<case>.npz    full trajectories with *injected* draws: np.random.randn /
                       rand / random / exponential are replaced by readers of
                       pre-drawn arrays Z[a,d,N], U[a,3,N], U0[a] indexed by
                       (attempt, slot, particle); the reference's own draw_from /
                       min_idx / HMCState code runs unmodified on top of them.

No reference source is copied; only its outputs are stored.
"""
import builtins
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _import_reference():
    builtins.xrange = range
    link_dir = "/tmp/mjhmc_ref_link"
    os.makedirs(link_dir, exist_ok=True)
    link = os.path.join(link_dir, "MJHMC")          # utils.package_path wants 'MJHMC' in sys.path
    if not os.path.islink(link):
        os.symlink(REF, link)
    sys.dont_write_bytecode = True
    sys.path.insert(0, link)
    from mjhmc.samplers import markov_jump_hmc as mj
    from mjhmc.misc import distributions as dist
    return mj, dist


mj, dist = _import_reference()


def _bypass(cls):
    sub = type(cls.__name__, (cls,), {"init_X": lambda self: self.gen_init_X()})
    return sub


RoughWell = _bypass(dist.RoughWell)
Gaussian = _bypass(dist.Gaussian)
TestGaussian = _bypass(dist.TestGaussian)

SAMPLERS = {c.__name__: c for c in (mj.HMCBase, mj.HMC, mj.ControlHMC, mj.ContinuousTimeHMC, mj.MarkovJumpHMC)}


# ----------------------------------------------------------------------------
# 1. seeded known answers
# ----------------------------------------------------------------------------
def seeded_runs():
    out = []
    cases = [
        ("HMCBase", {}, "RoughWell", dict(ndims=2, nbatch=100)),
        ("HMC", {}, "RoughWell", dict(ndims=2, nbatch=100)),
        ("ControlHMC", {}, "RoughWell", dict(ndims=2, nbatch=100)),
        ("ContinuousTimeHMC", dict(resample=False), "RoughWell", dict(ndims=2, nbatch=100)),
        ("MarkovJumpHMC", dict(resample=False), "RoughWell", dict(ndims=2, nbatch=100)),
        ("MarkovJumpHMC", dict(resample=True), "RoughWell", dict(ndims=2, nbatch=100)),
        ("ContinuousTimeHMC", dict(resample=True), "RoughWell", dict(ndims=2, nbatch=100)),
        ("ControlHMC", {}, "Gaussian", dict(ndims=10, nbatch=50, log_conditioning=2)),
        ("MarkovJumpHMC", dict(resample=False), "Gaussian", dict(ndims=10, nbatch=50, log_conditioning=2)),
        ("HMCBase", {}, "TestGaussian", dict(ndims=3, nbatch=40)),
        ("MarkovJumpHMC", dict(resample=False), "TestGaussian", dict(ndims=3, nbatch=40)),
    ]
    dists = dict(RoughWell=RoughWell, Gaussian=Gaussian, TestGaussian=TestGaussian)
    for sname, skw, dname, dkw in cases:
        np.random.seed(1)
        d = dists[dname](**dkw)
        s = SAMPLERS[sname](distribution=d, epsilon=1.0, beta=0.1, num_leapfrog_steps=5, **skw)
        X = s.sample(10)
        out.append(dict(
            sampler=sname, sampler_kwargs=skw, distribution=dname, distribution_kwargs=dkw,
            hp=dict(epsilon=1.0, beta=0.1, num_leapfrog_steps=5), seed=1, n_samples=10,
            counters=dict(l=int(s.l_count), f=int(s.f_count), fl=int(s.fl_count), r=int(s.r_count),
                          E=int(d.E_count), dEdX=int(d.dEdX_count)),
            shape=list(X.shape), sum_X=float(X.sum()), sum_X2=float((X ** 2).sum()),
            sum_H=float(s.state.H().sum()), sum_final_X=float(s.state.X.sum()),
            sum_final_V=float(s.state.V.sum())))
    return out


def backoff_known_answer():
    np.random.seed(3)
    d = TestGaussian(ndims=1, nbatch=4)
    s = mj.MarkovJumpHMC(distribution=d, epsilon=1.0, beta=0.5, num_leapfrog_steps=1, resample=False)
    s.state.X[:] = np.array([[100., .1, .2, .3]])
    s.state.V[:] = 0.
    s.state.update_EX(); s.state.update_EV(); s.state.update_dEdX()
    e0, g0 = d.E_count, d.dEdX_count
    s.sampling_iteration()
    res = dict(dE=int(d.E_count - e0), ddEdX=int(d.dEdX_count - g0), l=int(s.l_count), f=int(s.f_count),
               r=int(s.r_count), X=s.state.X.tolist(), epsilon=float(s.epsilon), L=int(s.num_leapfrog_steps),
               cache_active=[bool(b) for b in s.state.cache_active])
    # same state under ContinuousTimeHMC raises
    np.random.seed(3)
    d2 = TestGaussian(ndims=1, nbatch=4)
    c = mj.ContinuousTimeHMC(distribution=d2, epsilon=1.0, beta=0.5, num_leapfrog_steps=1, resample=False)
    c.state.X[:] = np.array([[100., .1, .2, .3]])
    c.state.V[:] = 0.
    c.state.update_EX(); c.state.update_EV(); c.state.update_dEdX()
    try:
        c.sampling_iteration()
        res["ct_raises"] = False
    except ValueError:
        res["ct_raises"] = True
    return res


# ----------------------------------------------------------------------------
# 2. injected-draw trajectories
# ----------------------------------------------------------------------------
class Injector(object):
    """Replaces the four np.random entry points the hot path uses by array readers."""

    def __init__(self, kind, Z, U, U0):
        self.kind, self.Z, self.U, self.U0 = kind, Z, U, U0
        self.attempt = 0
        self.rand_calls = 0
        self.slot = 0
        self.exp_iter = None
        self.log = []

    # np.random.randn(d, N)
    def randn(self, *shape):
        continuous_ct = self.kind == "ContinuousTimeHMC"
        a = self.attempt - 1 if (continuous_ct or self.kind in ("HMCBase", "HMC", "ControlHMC")) else self.attempt
        z = self.Z[a]
        assert z.shape == tuple(shape)
        return z.copy()

    # np.random.rand(N): accept, then flip
    def rand(self, n):
        u = self.U[self.attempt, self.rand_calls].copy()
        self.rand_calls += 1
        return u

    # np.random.random(): the batch-wide coin closes a discrete attempt
    def random(self, *args):
        u = float(self.U0[self.attempt])
        self.attempt += 1
        self.rand_calls = 0
        return u

    def exponential(self, scale=1.0):
        return scale * (-np.log(1.0 - next(self.exp_iter)))

    def wrap_draw_from(self, real):
        def draw_from(rates):
            u = self.U[self.attempt, self.slot]
            self.exp_iter = iter(u[np.asarray(rates) != 0])
            try:
                out = real(rates)
            except ValueError:
                self.attempt += 1
                self.slot = 0
                raise
            self.slot += 1
            if self.slot == 3:
                self.slot = 0
                self.attempt += 1
            return out
        return draw_from


def run_injected(sname, dist_obj, hp, n_iter, seed, V0=None, force_state=None, extra_attempts=4):
    d, N = dist_obj.ndims, dist_obj.nbatch
    rs = np.random.RandomState(seed)
    A = n_iter + extra_attempts
    Z = rs.randn(A, d, N)
    U = rs.rand(A, 3, N)
    U0 = rs.rand(A)
    X0 = dist_obj.Xinit.copy()
    V0 = rs.randn(d, N) if V0 is None else V0

    inj = Injector(sname, Z, U, U0)
    saved = (np.random.randn, np.random.rand, np.random.random, np.random.exponential, mj.draw_from)
    # constructor draws (Xinit re-generation, V) are real np.random; we overwrite the state afterwards
    kwargs = dict(distribution=dist_obj, **hp)
    if sname in ("ContinuousTimeHMC", "MarkovJumpHMC"):
        kwargs["resample"] = False
    s = SAMPLERS[sname](**kwargs)
    s.state.X[:] = X0
    s.state.V[:] = V0
    if force_state is not None:
        force_state(s.state)
    s.state.update_EX(); s.state.update_EV(); s.state.update_dEdX()
    dist_obj.E_count = N
    dist_obj.dEdX_count = N
    X0, V0 = s.state.X.copy(), s.state.V.copy()
    rec = dict(X=[], V=[], EX=[], EV=[], dwell=[], counters=[], cache=[], attempts=[])
    try:
        np.random.randn, np.random.rand, np.random.random = inj.randn, inj.rand, inj.random
        np.random.exponential = inj.exponential
        mj.draw_from = inj.wrap_draw_from(saved[4])
        import contextlib, io
        for _ in range(n_iter):
            with contextlib.redirect_stdout(io.StringIO()):
                s.sampling_iteration()
            rec["X"].append(s.state.X.copy()); rec["V"].append(s.state.V.copy())
            rec["EX"].append(s.state.EX[0].copy()); rec["EV"].append(s.state.EV[0].copy())
            rec["dwell"].append(np.asarray(getattr(s, "dwelling_times", np.zeros(N))).copy())
            rec["cache"].append(s.state.cache_active.copy())
            rec["counters"].append([s.l_count, s.f_count, s.fl_count, s.r_count,
                                    dist_obj.E_count, dist_obj.dEdX_count])
            rec["attempts"].append(inj.attempt)
    finally:
        (np.random.randn, np.random.rand, np.random.random, np.random.exponential, mj.draw_from) = saved
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(X0=X0, V0=V0, Z=Z, U=U, U0=U0,
               epsilon=np.float64(hp["epsilon"]), beta_arg=np.float64(hp["beta"]),
               L=np.int64(hp["num_leapfrog_steps"]),
               final_epsilon=np.float64(s.epsilon), final_L=np.int64(s.num_leapfrog_steps))
    return out


def injected_cases():
    cases = {}
    hp_tame = dict(epsilon=0.5, beta=0.3, num_leapfrog_steps=4)
    np.random.seed(11)
    for sname in SAMPLERS:
        cases["%s_roughwell2" % sname] = (sname, RoughWell(ndims=2, nbatch=96, scale1=3, scale2=4), hp_tame, 12, 100)
        cases["%s_testgauss3" % sname] = (sname, TestGaussian(ndims=3, nbatch=64, sigma=1.5),
                                          dict(epsilon=0.9, beta=0.2, num_leapfrog_steps=3), 12, 101)
        cases["%s_gauss5" % sname] = (sname, Gaussian(ndims=5, nbatch=80, log_conditioning=2),
                                      dict(epsilon=0.7, beta=0.4, num_leapfrog_steps=5), 12, 102)
    out = {}
    for name, (sname, dobj, hp, n_iter, seed) in cases.items():
        out[name] = run_injected(sname, dobj, hp, n_iter, seed)
        out[name]["dist"] = np.array(type(dobj).__name__)
        if isinstance(dobj, RoughWell):
            out[name]["dist_params"] = np.array([dobj.scale1, dobj.scale2], dtype=np.float64)
        elif isinstance(dobj, TestGaussian):
            out[name]["dist_params"] = np.array([dobj.sigma], dtype=np.float64)
        else:
            out[name]["dist_params"] = np.asarray(dobj.J, dtype=np.float64)

    # back-off under injection: one particle with a 937.5 energy drop
    def force(state):
        state.X[0, 0] = 100.
        state.V[:, 0] = 0.
    np.random.seed(12)
    dobj = TestGaussian(ndims=1, nbatch=8, sigma=1.)
    out["MarkovJumpHMC_backoff"] = run_injected("MarkovJumpHMC", dobj, dict(epsilon=1.0, beta=0.5, num_leapfrog_steps=1),
                                                4, 103, force_state=force)
    out["MarkovJumpHMC_backoff"]["dist"] = np.array("TestGaussian")
    out["MarkovJumpHMC_backoff"]["dist_params"] = np.array([1.0])
    return out


def main():
    seeded = dict(runs=seeded_runs(), backoff=backoff_known_answer())
    with open(os.path.join(HERE, "seeded_runs.json"), "w") as f:
        json.dump(seeded, f, indent=1, sort_keys=True)
    for name, arrs in injected_cases().items():
        np.savez_compressed(os.path.join(HERE, "inject_%s.npz" % name), **arrs)
        print(name, arrs["counters"][-1], "attempts", arrs["attempts"][-1])
    print("wrote", len(seeded["runs"]), "seeded runs")


if __name__ == "__main__":
    main()
