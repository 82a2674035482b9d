"""Host view of the particle state (reference: mjhmc/samplers/hmc_state.py).

On the device only X and V (plus one cached energy and one flag per particle for the FLF
cache) are stored; EX, EV and dEdX are functions of (X, V).  ``HMCState`` is the numpy view
callers of the reference read and assign (``sampler.state.X``, ``.V``, ``.EX``, ``.EV``,
``.dEdX``, ``.H()``, ``.copy()``); the derived arrays are evaluated lazily on the GPU without
touching the evaluation counters.
"""
import numpy as np


class HMCState(object):
    """Holds all the state variables for sampling particles (numpy, float64)."""

    def __init__(self, X, parent, V=None, cache_active=None, H_cache=None):
        self.parent = parent
        self.X = np.array(X, dtype=np.float64)
        self.nbatch = self.X.shape[1]
        self.V = np.random.randn(*self.X.shape) if V is None else np.array(V, dtype=np.float64)   # hmc_state.py:26
        self.active_idx = np.arange(self.nbatch)
        self.cache_active = (np.zeros(self.nbatch, dtype=bool) if cache_active is None
                             else np.array(cache_active, dtype=bool))
        self.H_cache = np.zeros(self.nbatch) if H_cache is None else np.array(H_cache, dtype=np.float64)

    @classmethod
    def from_buffers(cls, parent, X, V, cache_active=None, H_cache=None):
        """B200 extension: wrap existing host buffers (numpy arrays or pinned torch tensors) without copying;
        assigning the result to ``sampler.state`` uploads them at the next launch."""
        st = cls.__new__(cls)
        st.parent, st.X, st.V = parent, X, V
        st.nbatch = X.shape[1]
        if cache_active is not None:
            st.cache_active = cache_active
        if H_cache is not None:
            st.H_cache = H_cache
        # an empty FLF cache is cleared on the device instead of being uploaded; active_idx and the empty cache arrays
        # are made when somebody reads them (__getattr__): a megabyte-sized np.arange per assignment is a millisecond
        # of the end-to-end step
        st._empty_cache = cache_active is None and H_cache is None
        return st

    def __getattr__(self, name):
        # only reached when normal lookup fails: the lazily built members of a from_buffers() state
        nb = self.__dict__.get("nbatch")
        if nb is not None and name in ("active_idx", "cache_active", "H_cache"):
            val = np.arange(nb) if name == "active_idx" else (np.zeros(nb, dtype=bool) if name == "cache_active" else np.zeros(nb))
            self.__dict__[name] = val
            return val
        raise AttributeError(name)

    # derived arrays (hmc_state.py:28-39, 46-53), evaluated on the device, not counted
    @property
    def EX(self):
        return np.asarray(self.parent._uncounted_E(self.X), dtype=np.float64).reshape((1, -1))

    @property
    def EV(self):
        return (np.sum(self.V ** 2, axis=0) / 2.).reshape((1, -1))

    @property
    def dEdX(self):
        return np.asarray(self.parent._uncounted_dEdX(self.X), dtype=np.float64)

    def H(self):
        """returns the full energy of the state (hmc_state.py:80-84)"""
        return self.EX + self.EV

    def copy(self):
        return HMCState(self.X.copy(), self.parent, V=self.V.copy(), cache_active=self.cache_active.copy(),
                        H_cache=self.H_cache.copy())

    def get_state(self):
        return np.concatenate((self.X, self.V))

    # -- the operators of hmc_state.py:86-129 on the host view.  Arithmetic stays on the GPU: the leapfrog pieces are
    # the unfused kernels (mjhmc_kick_drift / mjhmc_kick), the gradient and energy the distribution's device
    # evaluation, counted like the reference counts them (parent.dEdX / parent.E).
    def _device_L(self, sign):
        import torch
        from .. import _device, _lib
        par = self.parent
        eng = par._engine
        lib = _lib.load()
        with eng.ctx():
            X = _device.to_device(self.X, eng.dtype, eng.device)
            V = _device.to_device(sign * self.V, eng.dtype, eng.device)
            G = eng._callback(X, True, count=False)          # dEdX of the current state: evaluated when it was made
            m, eps = X.shape[1], float(par.epsilon)
            for _ in range(int(par.num_leapfrog_steps)):
                _lib.check(lib.mjhmc_kick_drift(eng.code, eng.d, _device.ptr(X), _device.ptr(V), _device.ptr(G), m, m, eps,
                                                eng._stream()), "kick_drift")
                G = eng._callback(X, True)                    # counted: hmc_state.py:90 -> parent.dEdX
                _lib.check(lib.mjhmc_kick(eng.code, eng.d, _device.ptr(V), _device.ptr(G), m, m, eps, eng._stream()), "kick")
            eng._callback(X, False)                           # update_EX (hmc_state.py:99): one counted energy evaluation
            self.X = X.double().cpu().numpy()
            self.V = sign * V.double().cpu().numpy()
        return self

    def leapfrog(self):
        """One leapfrog step (hmc_state.py:86-91)."""
        par, keep = self.parent, self.parent.num_leapfrog_steps
        par.num_leapfrog_steps = 1
        try:
            return self._device_L(1.0)
        finally:
            par.num_leapfrog_steps = keep

    def L(self):
        """Integration operator: num_leapfrog_steps leapfrog steps, then the energies (hmc_state.py:93-100)."""
        return self._device_L(1.0)

    def F(self):
        """Explicit flip operator (hmc_state.py:102-107)."""
        self.V = -self.V
        return self

    def FLF(self):
        """F L F (hmc_state.py:109-119); the host view integrates every column (its cache flags describe the
        sampler's device state, not this copy)."""
        return self._device_L(-1.0)

    def R(self):
        """Momentum corruption V = V sqrt(1 - beta) + randn sqrt(beta) with the parent's beta (hmc_state.py:121-129)."""
        beta = float(self.parent.beta)
        self.V = self.V * np.sqrt(1 - beta) + np.random.randn(*self.V.shape) * np.sqrt(beta)
        return self

    def update(self, idx, state):
        """Replace the columns idx of this state by those of `state` (hmc_state.py:63-72)."""
        if len(idx) == 0:
            return
        self.X[:, idx] = state.X[:, idx]
        self.V[:, idx] = state.V[:, idx]
        self.cache_active[idx] = state.cache_active[idx]
        self.H_cache[idx] = state.H_cache[idx]

    def reset_flf_cache(self):
        self.cache_active = np.zeros_like(self.cache_active)
