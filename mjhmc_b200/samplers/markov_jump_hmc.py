"""The sampler classes of the reference (mjhmc/samplers/markov_jump_hmc.py) on the GPU.

Same names, constructor arguments, attributes and error behaviour as the reference:
``HMCBase``, ``HMC``, ``ControlHMC``, ``ContinuousTimeHMC``, ``MarkovJumpHMC`` with
``sample(n_samples, preserve_order)``, ``sampling_iteration()``, ``burn_in()``, the operator
counters ``l_count / f_count / fl_count / r_count``, ``dwelling_times``, ``state`` ...

Underneath, one ``sample(n)`` call is one launch of the fused sampler kernel
(csrc/fused_elementwise.cuh, csrc/dense.cu) that keeps every particle on chip across the n
iterations; energies that only exist as Python callables run the unfused path
(kernels for the leapfrog pieces and the transition, the callable in between).

B200-specific keyword arguments (all optional, accepted by every class):
    dtype            'float64' (default, the reference's arithmetic) or 'float32'
    seed             Philox key; default: drawn from np.random so np.random.seed() pins a run
    device           torch device string, default current CUDA device
    injected_draws   dict(Z=(A,d,N), U=(A,3,N), U0=(A,)) -> INJECT mode (trajectory parity tests)
    particle_offset  global index of this shard's first particle (multi-GPU sharding)
    sharded          True or a torch.distributed process group: this sampler holds one shard of a particle cloud
                     spread over the ranks of the group; the batch-global couplings of the reference (infinite-rate
                     back-off markov_jump_hmc.py:376-389, dwell-time resampling :321-328) are then taken over the
                     WHOLE cloud (mjhmc_b200/parallel.py), so results do not depend on the number of ranks
    V                initial momentum (default np.random.randn like hmc_state.py:26)
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _device, _lib, parallel
from ..misc.distributions import Distribution
from ..misc.utils import overrides
from .hmc_state import HMCState

#pylint: disable=too-many-instance-attributes
#pylint: disable=too-many-arguments

INFINITE_RATE_MSG = ("Infinite rate. This occurs when calculating transition rates "
                     "between states that have a very large energy difference, such that "
                     "the transition probability is less than the numerical precision. "
                     "Try decreasing the leapfrog stepsize/number of steps or dividing "
                     " the energy by a large constant.")

_B200_KWARGS = ("dtype", "seed", "device", "injected_draws", "particle_offset", "V", "kernel", "sharded", "literal_race", "register_state")


class _CallableEnergy(Distribution):
    """Wraps the (Xinit, E, dEdX) constructor form (markov_jump_hmc.py:56-65)."""

    def __init__(self, Xinit, E, dEdX):
        self._E, self._dEdX, self._Xinit = E, dEdX, Xinit
        super(_CallableEnergy, self).__init__(ndims=Xinit.shape[0], nbatch=Xinit.shape[1])

    def E_val(self, X):
        return np.asarray(self._E(X)).reshape((1, -1))

    def dEdX_val(self, X):
        return np.asarray(self._dEdX(X))

    def gen_init_X(self):
        self.Xinit = self._Xinit

    def __hash__(self):
        return id(self)


def _bound_distribution(E, dEdX):
    """(dist.E, dist.dEdX) of a built-in distribution -> that distribution (fused path)."""
    a, b = getattr(E, "__self__", None), getattr(dEdX, "__self__", None)
    if a is not None and a is b and isinstance(a, Distribution) and \
            getattr(E, "__name__", "") == "E" and getattr(dEdX, "__name__", "") == "dEdX":
        return a
    return None


class _Engine(object):
    """Device state + launches.  Fused when the distribution has a kernel descriptor the
    library supports, unfused (callback) otherwise."""

    def __init__(self, sampler, distribution, X0, V0, opts):
        self.sampler = sampler
        self.dist = distribution
        self.device = _device.require_cuda(opts.get("device"))
        self.lib = _lib.load()
        self.dtype = _device.norm_dtype(opts.get("dtype") or "float64")
        self.tdtype = _device.torch_dtype(self.dtype)
        self.code = _device.dtype_code(self.dtype)
        self.d, self.n = X0.shape
        self.offset = int(opts.get("particle_offset") or 0)
        # B200 extension (testing aid): evaluate the three holding times of every attempt literally in fp64 instead of
        # screening the race in single precision first (include/mjhmc_b200.h MJHMC_RNG_FLAG_LITERAL_RACE)
        self.literal_race = bool(opts.get("literal_race"))
        # B200 extension (testing aid): the register-resident form of the fused kernel also for ndims >= 6
        self.register_state = bool(opts.get("register_state")) or os.environ.get("MJHMC_B200_REGISTER_STATE") == "1"
        inj = opts.get("injected_draws")
        self.inj = None
        if inj is not None:
            f64 = lambda a: None if a is None else _device.to_device(np.asarray(a, dtype=np.float64), "float64", self.device)
            self.inj = dict(Z=f64(inj.get("Z")), U=f64(inj.get("U")), U0=f64(inj.get("U0")))
            self.inj_ld = int(np.asarray(inj["U"]).shape[-1])
        seed = opts.get("seed")
        self.seed = int(np.random.randint(0, 2 ** 62)) if seed is None else int(seed)
        with torch.cuda.device(self.device):
            desc = distribution.kernel_descriptor(self.dtype, self.device)
            self.fused = False
            if desc is not None:
                self.desc, self._desc_keep = desc
                self.fused = bool(self.lib.mjhmc_fused_supported(C.byref(self.desc)))
            # kernel="stream": force the TMA-ring streaming kernel (separable energies); default: the library
            # picks (register kernel for ndims <= 16, streaming kernel above, dense kernels for J / W)
            self.kernel = opts.get("kernel") or "auto"
            if self.kernel not in ("auto", "stream"):
                raise ValueError("kernel must be 'auto' or 'stream'")
            if self.kernel == "stream":
                if desc is None or not self.lib.mjhmc_stream_supported(C.byref(self.desc)):
                    raise ValueError("kernel='stream' needs a separable built-in energy with ndims <= 128")
                self.fused = True
            mk = lambda a: _device.to_device(a, self.dtype, self.device)
            self.X = [mk(X0), torch.empty((self.d, self.n), dtype=self.tdtype, device=self.device)]
            self.V = [mk(V0), torch.empty((self.d, self.n), dtype=self.tdtype, device=self.device)]
            self.Hc = [torch.zeros(self.n, dtype=self.tdtype, device=self.device) for _ in range(2)]
            self.ca = [torch.zeros(self.n, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self.cur = 0
            tmpl = np.zeros((_lib.COUNTER_ROWS, _lib.N_COUNTERS), dtype=np.int64)
            tmpl[:_lib.COUNTER_STRIPES, _lib.CNT_FAIL] = _lib.INT64_MAX
            self.cnt_template = torch.as_tensor(tmpl, device=self.device)
            self.counters = self.cnt_template.clone()
            # double-buffered like the state: a launch that is not committed (failed attempt, counting launch)
            # must leave sampler.dwelling_times untouched, as the reference does when draw_from raises
            self.dwell = [torch.zeros(self.n, dtype=torch.float64, device=self.device) for _ in range(2)]
            if not self.fused:
                # the reference's full HMCState (hmc_state.py:28-39): callables cannot be re-evaluated on chip
                self.G = self._callback(self.X[0], True, count=False)
                self.EX = self._callback(self.X[0], False, count=False).reshape(-1)
                self.EV = self._kinetic(self.V[0])
        self.launches = 0
        self.kernel_events = None
        # (epsilon, L) of the committed launch that may have left "energy valid" flags without cache_active
        # (flag value 2, set by an F move: H_cache = H_L at THAT epsilon / L); None = no such flags exist
        self._cache_key = None
        self._launch_key = None

    # ---------------------------------------------------------------- helpers
    def ctx(self):
        return torch.cuda.device(self.device)

    def _stream(self):
        return _device.stream_ptr(self.device)

    def _hp(self):
        s = self.sampler
        hp = _lib.HP()
        hp.sampler = s._sampler_code
        hp.num_leapfrog_steps = int(s.num_leapfrog_steps)
        hp.epsilon = float(s.epsilon)
        hp.beta = float(s.beta)
        hp.p_flip = float(s.p_flip)
        hp.p_r = float(s.p_r)
        return hp

    def _rng(self, attempt0):
        r = _lib.RNG()
        r.seed = self.seed
        r.attempt0 = int(attempt0)
        r.particle0 = self.offset
        r.flags = ((_lib.RNG_FLAG_LITERAL_RACE if self.literal_race else 0)
                   | (_lib.RNG_FLAG_REGISTER_STATE if self.register_state else 0))
        if self.inj is not None:
            r.mode = _lib.RNG_INJECT
            r.Z = self.inj["Z"].data_ptr() if self.inj["Z"] is not None else None
            r.U = self.inj["U"].data_ptr()
            r.U0 = self.inj["U0"].data_ptr() if self.inj["U0"] is not None else None
            r.inj_ld = self.inj_ld
            n_att = self.inj["U"].shape[0]
            if attempt0 >= n_att:
                raise IndexError("injected draws exhausted (attempt %d of %d)" % (attempt0, n_att))
        else:
            r.mode = _lib.RNG_PHILOX
        return r

    def _state(self, which):
        st = _lib.State()
        st.X = self.X[which].data_ptr()
        st.V = self.V[which].data_ptr()
        st.H_cache = self.Hc[which].data_ptr()
        st.cache_active = self.ca[which].data_ptr()
        st.n = self.n
        st.ld = self.n
        return st

    def _outputs(self, samples, it0, dwell, choice, energy=None):
        o = _lib.Outputs()
        esz = None
        if samples is not None:
            esz = samples.element_size()
            o.samples = samples.data_ptr() + it0 * samples.stride(1) * esz
            o.stride_k = samples.stride(0)
            o.stride_it = samples.stride(1)
        if dwell is not None:
            o.dwell = dwell.data_ptr() + it0 * self.n * 8
        if choice is not None:
            o.choice = choice.data_ptr() + it0 * self.n
        if energy is not None:
            o.energy = energy.data_ptr() + it0 * self.n * 8
        o.dwell_last = self.dwell[self.cur ^ 1 if self.fused else self.cur].data_ptr()
        o.counters = self.counters.data_ptr()
        return o

    def _read_counters(self):
        out = (C.c_int64 * _lib.N_COUNTERS)()
        _lib.check(self.lib.mjhmc_counters_read(_device.ptr(self.counters), out, self._stream()), "counters_read")
        return list(out)

    @property
    def dwell_last(self):
        return self.dwell[self.cur]

    def reset_cache(self):
        self.ca[self.cur].zero_()
        self._cache_key = None

    def _drop_stale_flf_energies(self):
        """The kernels keep H_L as the FLF energy of a particle that took an F move (flag 2: FLF(F z) = F L z).
        That energy belongs to the (epsilon, L) it was integrated with: after the infinite-rate back-off
        (markov_jump_hmc.py:376-389) or when the caller edits sampler.epsilon / num_leapfrog_steps the reference
        re-integrates with the new values (cache_active is False there), so those flags are dropped.  Flags of
        L movers (3) stay: the reference keeps their cached pre-move state whatever the hyper-parameters."""
        s = self.sampler
        key = (float(s.epsilon), int(s.num_leapfrog_steps))
        if self._cache_key is not None and key != self._cache_key:
            ca = self.ca[self.cur]
            ca.masked_fill_(ca == 2, 0)
            self._cache_key = None
        self._launch_key = key

    # ---------------------------------------------------------------- fused launch
    def launch(self, attempt0, n_iter, samples=None, it0=0, dwell=None, choice=None, energy=None):
        """Runs n_iter iterations from the current state into the spare buffers (no commit).
        Returns the folded counters of this launch."""
        with torch.cuda.device(self.device):
            self.counters.copy_(self.cnt_template)
            if self.fused:
                if self.sampler._sampler_code == _lib.SAMPLER_MARKOV_JUMP:
                    self._drop_stale_flf_energies()
                hp, rng = self._hp(), self._rng(attempt0)
                if self.inj is not None and attempt0 + n_iter > self.inj["U"].shape[0]:
                    raise IndexError("injected draws exhausted")
                src, dst = self._state(self.cur), self._state(self.cur ^ 1)
                o = self._outputs(samples, it0, dwell, choice, energy)
                entry = self.lib.mjhmc_sample_stream if self.kernel == "stream" else self.lib.mjhmc_sample_fused
                if self.kernel_events is not None:      # bench.py: CUDA events right around the sampler kernel
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                _lib.check(entry(C.byref(self.desc), C.byref(hp), C.byref(rng), C.byref(src),
                                 C.byref(dst), int(n_iter), C.byref(o), self._stream()), "sample_fused")
                if self.kernel_events is not None:
                    ev[1].record()
                    self.kernel_events.append(ev)
                self.launches += 1
                return self._read_counters()
            return self._launch_unfused(attempt0, n_iter, samples, it0, dwell, choice, energy)

    def commit(self):
        if self.fused:
            self.cur ^= 1
            self._cache_key = self._launch_key

    # ---------------------------------------------------------------- unfused (callback) path
    def _callback(self, Xd, want_grad, count=True):
        """Evaluates the distribution's callable on a device array (d, m) -> device array."""
        dist = self.dist
        if getattr(dist, "accepts_device_arrays", False):
            arg = Xd
        else:
            arg = Xd.detach().cpu().numpy().astype(np.float64)
        if want_grad:
            out = dist.dEdX(arg) if count else dist.dEdX_val(arg)
        else:
            out = dist.E(arg) if count else dist.E_val(arg)
        if isinstance(out, torch.Tensor):
            return out.to(device=self.device, dtype=self.tdtype).contiguous()
        return _device.to_device(np.asarray(out, dtype=np.float64), self.dtype, self.device)

    def _kinetic(self, Vd):
        m = Vd.shape[1]
        EV = torch.empty(m, dtype=self.tdtype, device=self.device)
        _lib.check(self.lib.mjhmc_kinetic(self.code, self.d, _device.ptr(Vd), m, m, _device.ptr(EV), self._stream()),
                   "kinetic")
        return EV

    def _L(self, X, V, G):
        """hmc_state.py:93-100 on (d, m) device arrays, in place; returns (G, EX, EV)."""
        s = self.sampler
        m = X.shape[1]
        eps = float(s.epsilon)
        for _ in range(int(s.num_leapfrog_steps)):
            _lib.check(self.lib.mjhmc_kick_drift(self.code, self.d, _device.ptr(X), _device.ptr(V), _device.ptr(G),
                                                 m, m, eps, self._stream()), "kick_drift")
            G = self._callback(X, True)
            _lib.check(self.lib.mjhmc_kick(self.code, self.d, _device.ptr(V), _device.ptr(G), m, m, eps,
                                           self._stream()), "kick")
        EV = self._kinetic(V)
        EX = self._callback(X, False).reshape(-1)
        return G, EX, EV

    def _launch_unfused(self, attempt0, n_iter, samples, it0, dwell, choice, energy=None):
        s = self.sampler
        total = [0] * _lib.N_COUNTERS
        total[_lib.CNT_FAIL] = _lib.INT64_MAX
        mj = s._sampler_code == _lib.SAMPLER_MARKOV_JUMP
        X, V = self.X[0], self.V[0]
        for it in range(n_iter):
            H_flf = None
            if mj:
                H_flf = torch.zeros(self.n, dtype=self.tdtype, device=self.device)
                idx = torch.nonzero(self.ca[0] == 0).reshape(-1)
                if idx.numel():
                    Xs, Vs, Gs = X[:, idx].contiguous(), (-V[:, idx]).contiguous(), self.G[:, idx].contiguous()
                    _, EXs, EVs = self._L(Xs, Vs, Gs)
                    H_flf[idx] = EXs + EVs
            Xp, Vp = X.clone(), V.clone()
            Gp, EXp, EVp = self._L(Xp, Vp, self.G.clone())
            snap = None
            if s._sampler_code != _lib.SAMPLER_DISCRETE:
                snap = [t.clone() for t in (X, V, self.G, self.EX, self.EV, self.Hc[0], self.ca[0], self.dwell[0])]
            cur = _lib.FullState(X.data_ptr(), V.data_ptr(), self.G.data_ptr(), self.EX.data_ptr(), self.EV.data_ptr())
            prop = _lib.FullState(Xp.data_ptr(), Vp.data_ptr(), Gp.data_ptr(), EXp.data_ptr(), EVp.data_ptr())
            hp, rng = self._hp(), self._rng(attempt0 + it)
            o = self._outputs(samples, it0 + it, dwell, choice, energy)
            self.counters.copy_(self.cnt_template)
            _lib.check(self.lib.mjhmc_transition(self.code, self.d, C.byref(hp), C.byref(rng), self.n, self.n,
                                                 C.byref(cur), C.byref(prop), _device.ptr(H_flf),
                                                 _device.ptr(self.Hc[0]), _device.ptr(self.ca[0]), C.byref(o),
                                                 self._stream()), "transition")
            self.launches += 1
            cnt = self._read_counters()
            if cnt[_lib.CNT_FAIL] != _lib.INT64_MAX:
                for dst, src in zip((X, V, self.G, self.EX, self.EV, self.Hc[0], self.ca[0], self.dwell[0]), snap):
                    dst.copy_(src)
                total[_lib.CNT_FAIL] = it
                return total
            for c in (_lib.CNT_L, _lib.CNT_F, _lib.CNT_FL, _lib.CNT_R):
                total[c] += cnt[c]
        return total

    # ---------------------------------------------------------------- host views
    def download(self):
        c = self.cur
        return (self.X[c].double().cpu().numpy(), self.V[c].double().cpu().numpy(),
                (self.ca[c].cpu().numpy() & 1).astype(bool), self.Hc[c].double().cpu().numpy())

    def upload(self, st):
        c = self.cur
        for dst, src in ((self.X[c], st.X), (self.V[c], st.V)):
            if isinstance(src, torch.Tensor) and src.dtype == dst.dtype and src.is_pinned() and src.shape == dst.shape:
                dst.copy_(src, non_blocking=True)           # pinned host buffer: one async H2D, no staging copy
            else:
                dst.copy_(_device.to_device(src, self.dtype, self.device))
        if getattr(st, "_empty_cache", False):
            self.ca[c].zero_()
            self.Hc[c].zero_()
        else:
            self.ca[c].copy_(torch.as_tensor(np.asarray(st.cache_active, dtype=np.uint8) * 3, device=self.device))
            self.Hc[c].copy_(_device.to_device(st.H_cache, self.dtype, self.device))
        self._cache_key = None
        if not self.fused:
            self.G = self._callback(self.X[0], True, count=False)
            self.EX = self._callback(self.X[0], False, count=False).reshape(-1)
            self.EV = self._kinetic(self.V[0])


class HMCBase(object):
    """
    The base class for all HMC samplers in this file.
    Not a useful sampler in of itself but provides a useful structure
      and serves as a control
    """
    _sampler_code = _lib.SAMPLER_DISCRETE

    def __init__(self, Xinit=None, E=None, dEdX=None,
                 epsilon=1e-4, alpha=0.2, beta=None,
                 num_leapfrog_steps=5, distribution=None, **b200):
        unknown = set(b200) - set(_B200_KWARGS)
        if unknown:
            raise TypeError("__init__() got an unexpected keyword argument %r" % sorted(unknown)[0])
        self._opts = b200
        sharded = b200.get("sharded")
        self._group = None if not sharded else parallel.resolve_group(sharded)
        self._engine = None
        self._host_state = None
        self._attempt = 0
        # B200 extension: gradient evaluations the device actually executed (<= dEdX_count, which
        # keeps the reference's accounting even where an FLF energy comes from the cache)
        self.grad_evals_executed = 0
        # do not execute this block if I am an instance of MarkovJumpHMC (markov_jump_hmc.py:46)
        if not isinstance(self, MarkovJumpHMC):
            if isinstance(distribution, Distribution):
                distribution.mjhmc = False
                distribution.reset()
                self._bind(distribution)
            else:
                assert Xinit is not None
                assert E is not None
                assert dEdX is not None
                bound = _bound_distribution(E, dEdX)
                if bound is not None and bound.kernel_descriptor("float64", None if not torch.cuda.is_available() else "cuda") is not None:
                    bound.Xinit = np.array(Xinit)
                    bound.nbatch = Xinit.shape[1]
                    self._bind(bound)
                else:
                    self._bind(_CallableEnergy(np.array(Xinit), E, dEdX))

        self.num_leapfrog_steps = num_leapfrog_steps
        self.epsilon = epsilon
        self.beta = beta or alpha**(1./(self.epsilon*self.num_leapfrog_steps))

        self.original_epsilon = epsilon
        self.original_l = self.num_leapfrog_steps

        self.n_burn_in = 500

        # these settings for the base class only
        self.p_flip = 0.5
        self.p_r = 1

        # total operator counts. counted per particle
        self.l_count = 0
        self.f_count = 0
        # this one is necessary since we're not always flipping the momentum
        self.fl_count = 0
        self.r_count = 0

        # only approximate!! lower bound
        self.grad_per_sample_step = self.num_leapfrog_steps

    # ------------------------------------------------------------------ construction
    def _bind(self, distribution):
        """Builds the device state: HMCState(Xinit.copy(), self) of markov_jump_hmc.py:54,233."""
        self.distribution = distribution
        self.ndims = distribution.Xinit.shape[0]
        self.nbatch = distribution.Xinit.shape[1]
        self.energy_func = distribution.E
        self.grad_func = distribution.dEdX
        X0 = distribution.Xinit
        V0 = self._opts.get("V")
        if V0 is None:
            V0 = np.random.randn(self.ndims, self.nbatch)       # hmc_state.py:26
        self._engine = _Engine(self, distribution, X0, V0, self._opts)
        if self._engine.fused:
            # the state constructor evaluates E and dEdX once (hmc_state.py:28-39); here both are
            # recomputed on chip whenever needed, so only the counters move
            distribution.E_count += self.nbatch
            distribution.dEdX_count += self.nbatch
        else:
            # the callbacks did run (uncounted above); count them like the reference
            distribution.E_count += self.nbatch
            distribution.dEdX_count += self.nbatch
        self._host_state = None

    # to deprecate
    def E(self, X):
        """compute energy function at X"""
        return np.asarray(self.energy_func(X)).reshape((1, -1))

    # to deprecate
    def dEdX(self, X):
        """compute energy function gradient at X"""
        return self.grad_func(X)

    def _uncounted_E(self, X):
        return self.distribution.E_val(X)

    def _uncounted_dEdX(self, X):
        return self.distribution.dEdX_val(X)

    # ------------------------------------------------------------------ state view
    @property
    def state(self):
        if self._host_state is None:
            X, V, ca, Hc = self._engine.download()
            self._host_state = HMCState(X, self, V=V, cache_active=ca, H_cache=Hc)
        return self._host_state

    @state.setter
    def state(self, st):
        self._host_state = st

    def _sync_state_to_device(self):
        """A handed-out HMCState may have been edited (or replaced) by the caller."""
        if self._host_state is not None:
            self._engine.upload(self._host_state)
            self._host_state = None

    @property
    def dwelling_times(self):
        return self._engine.dwell_last.cpu().numpy()

    # ------------------------------------------------------------------ iteration driver
    def _accumulate(self, cnt, transitions=True):
        if transitions:
            self.l_count += cnt[_lib.CNT_L]
            self.f_count += cnt[_lib.CNT_F]
            self.fl_count += cnt[_lib.CNT_FL]
            self.r_count += cnt[_lib.CNT_R]
        if self._engine.fused:
            self.distribution.E_count += cnt[_lib.CNT_E]
            self.distribution.dEdX_count += cnt[_lib.CNT_DEDX]
            self.grad_evals_executed += cnt[_lib.CNT_EXEC]

    def _run(self, n, samples=None, it0=0, dwell=None, choice=None, energy=None):
        """n sampling iterations, incl. the infinite-rate protocol (markov_jump_hmc.py:364-389)."""
        eng = self._engine
        done = 0
        while done < n:
            m = n - done
            cnt = eng.launch(self._attempt, m, samples, it0 + done, dwell, choice, energy)
            fail = cnt[_lib.CNT_FAIL]
            if self._group is not None:
                # the back-off is batch-wide in the reference (markov_jump_hmc.py:376-389): every rank replays,
                # counts and retries at the first failing iteration of the WHOLE cloud
                if not eng.fused:
                    raise NotImplementedError("sharded runs need a fused energy (callables advance in place)")
                fail = parallel.allreduce_min(fail, self._group)
            if fail == _lib.INT64_MAX:
                eng.commit()
                self._accumulate(cnt)
                self._attempt += m
                done += m
                continue
            if eng.fused:
                if fail > 0:
                    # replay the iterations before the failing one (deterministic streams), keep them
                    cnt = eng.launch(self._attempt, fail, samples, it0 + done, dwell, choice, energy)
                    assert cnt[_lib.CNT_FAIL] == _lib.INT64_MAX
                    eng.commit()
                    self._accumulate(cnt)
                # the failed attempt: its energy / gradient evaluations stay counted, nothing else happens
                cnt = eng.launch(self._attempt + fail, 1)
                self._accumulate(cnt, transitions=False)
            else:
                # the unfused engine advanced in place up to the failing iteration and restored it
                self._accumulate(cnt)
            self._attempt += fail
            done += fail
            self._attempt += 1
            self._on_infinite_rate(samples, it0 + done, dwell, choice, energy)
            done += 1

    def _on_infinite_rate(self, samples, it, dwell, choice, energy=None):
        raise ValueError(INFINITE_RATE_MSG)

    def _advance(self, n, record=True, want_dwell=False, want_choice=False, want_energy=False):
        eng = self._engine
        self._sync_state_to_device()
        samples = dwell = choice = None
        with eng.ctx():
            if record:
                samples = torch.empty((self.ndims, n, self.nbatch), dtype=eng.tdtype, device=eng.device)
            if want_dwell:
                dwell = torch.empty((n, self.nbatch), dtype=torch.float64, device=eng.device)
            if want_choice:
                choice = torch.empty((n, self.nbatch), dtype=torch.uint8, device=eng.device)
            energy = torch.empty((n, self.nbatch), dtype=torch.float64, device=eng.device) if want_energy else None
            self._run(n, samples, 0, dwell, choice, energy)
        if want_energy:
            return samples, dwell, choice, energy
        return samples, dwell, choice

    def _snapshot(self):
        """Device state + counters, so a chunk of iterations can be replayed (generate_samples)."""
        eng, d = self._engine, self.distribution
        c = eng.cur
        tensors = [t.clone() for t in (eng.X[c], eng.V[c], eng.Hc[c], eng.ca[c], eng.dwell_last)]
        ints = (self._attempt, self.l_count, self.f_count, self.fl_count, self.r_count, d.E_count, d.dEdX_count,
                self.grad_evals_executed)
        return tensors, ints, eng._cache_key

    def _restore(self, snap):
        eng, d = self._engine, self.distribution
        c = eng.cur
        for dst, src in zip((eng.X[c], eng.V[c], eng.Hc[c], eng.ca[c], eng.dwell_last), snap[0]):
            dst.copy_(src)
        (self._attempt, self.l_count, self.f_count, self.fl_count, self.r_count, d.E_count, d.dEdX_count,
         self.grad_evals_executed) = snap[1]
        eng._cache_key = snap[2]
        self._host_state = None

    def sampling_iteration(self):
        """Perform a single sampling step"""
        self._advance(1, record=False)

    def sample_device(self, n_samples=1000):
        """B200 extension: like sample() but returns the device tensor (ndims, n_samples, nbatch)
        without the device->host copy."""
        return self._advance(n_samples)[0]

    def sample(self, n_samples=1000, preserve_order=False, num_steps=None):
        """
        Draws nsamples, returns them all

        Args:
           n_samples: number of samples to draw - int  (``num_steps`` is accepted as an alias, README.md:35)
           preserve_order: if True, time is given it's own axis.
              otherwise, it is rolled into the batch axis

        Returns:
           if preserve_order:
               samples - [n_dim, n_batch, n_samples]
           else:
               samples - [n_dim, n_batch * n_samples]
        """
        if num_steps is not None:
            n_samples = num_steps
        return self._sample_plain(n_samples, preserve_order)

    def _sample_plain(self, n_samples, preserve_order):
        """n iterations, every state recorded (markov_jump_hmc.py:166-173 / :331-338)."""
        if not preserve_order and self._pipeline_chunks(n_samples) > 1:
            return self._sample_pipelined(n_samples)
        S = self._advance(n_samples)[0]
        return self._to_host(S, preserve_order)

    # -- host transfer ------------------------------------------------------------------------------
    PIPELINE_MIN_BYTES = 64 << 20      # below this one launch + one copy is cheaper than a pipeline

    def _pipeline_chunks(self, n_samples):
        eng = self._engine
        if eng.device.type != "cuda" or n_samples < 8:
            return 1
        nbytes = self.ndims * n_samples * self.nbatch * torch.empty((), dtype=eng.tdtype).element_size()
        return min(8, n_samples // 4) if nbytes >= self.PIPELINE_MIN_BYTES else 1

    def _sample_pipelined(self, n_samples):
        """sample() for large outputs: the iterations run in a few launches and the device->host copy of
        chunk c (side stream, pinned destination) overlaps the kernel of chunk c+1."""
        eng = self._engine
        chunks = self._pipeline_chunks(n_samples)
        d, N = self.ndims, self.nbatch
        out = torch.empty((d, n_samples, N), dtype=eng.tdtype, pin_memory=True)
        if not hasattr(eng, "copy_stream"):
            eng.copy_stream = torch.cuda.Stream(device=eng.device)
        main = torch.cuda.current_stream(eng.device)
        # the copy engine is the bottleneck (a chunk's copy outlasts the next chunk's kernel), so the only kernel
        # time it ever waits for is the first chunk's: that chunk is a quarter of the others
        first = max(1, n_samples // (4 * chunks)) if chunks > 1 else n_samples
        rest = n_samples - first
        bounds = [0] + [first + rest * c // max(1, chunks - 1) for c in range(chunks)] if chunks > 1 else [0, n_samples]
        keep = []
        for c in range(chunks):
            c0, c1 = bounds[c], bounds[c + 1]
            S = self._advance(c1 - c0)[0]                  # returns after the launch completed (counter read)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(eng.copy_stream):
                eng.copy_stream.wait_event(ready)
                for k in range(d):                         # contiguous (c1-c0)*N runs: plain async memcpys
                    out[k, c0:c1, :].copy_(S[k], non_blocking=True)
            S.record_stream(eng.copy_stream)
            keep.append(S)
        eng.copy_stream.synchronize()
        return out.numpy().reshape(d, n_samples * N).astype(np.float64, copy=False)

    def _d2h(self, t):
        """Device tensor -> fresh host numpy array through a pinned staging tensor (torch's caching
        host allocator recycles the pinned blocks once the caller drops the array)."""
        if self._engine.device.type != "cuda":
            return t.numpy()
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t, non_blocking=True)
        torch.cuda.current_stream(self._engine.device).synchronize()
        return buf.numpy()

    def _to_host(self, S, preserve_order):
        d, n, N = S.shape
        if preserve_order:
            return self._d2h(S.permute(0, 2, 1).contiguous()).astype(np.float64, copy=False)
        return self._d2h(S.reshape(d, n * N)).astype(np.float64, copy=False)

    def burn_in(self):
        """Runs the sample for a number of burn in sampling iterations"""
        self._advance(self.n_burn_in, record=False)


class HMC(HMCBase):
    """Implements standard HMC
    """

    def __init__(self, *args, **kwargs):
        super(HMC, self).__init__(*args, **kwargs)
        self.p_flip = 1


class ControlHMC(HMCBase):
    """Standard HMC but randomize all of the momentum some of the time
    """

    def __init__(self, *args, **kwargs):
        super(ControlHMC, self).__init__(*args, **kwargs)
        self.p_flip = 1
        with np.errstate(divide='ignore'):
            self.p_r = - np.log(1 - self.beta) * 0.5
        # tells hmc state to randomize all of the momentum when R is called
        self.beta = 1


class ContinuousTimeHMC(HMCBase):
    """Base class for all markov jump HMC samplers
    """
    _sampler_code = _lib.SAMPLER_CONTINUOUS_TIME

    def __init__(self, *args, **kwargs):
        """ Initalizer method for continuous-time samplers

        :param resample: boolean flag whether to resample or not. ALWAYS set to true unless you
           have a specific reason not to. Produced samples will be biased if resample is false
        """
        self.resample = kwargs.pop('resample', True)
        distribution = kwargs.get('distribution')
        super(ContinuousTimeHMC, self).__init__(*args, **kwargs)
        # transformation from discrete beta to insure matching autocorrelation
        with np.errstate(divide='ignore'):
            self.p_r = - np.log(1 - self.beta) * 0.5
        # tells hmc state to randomize all of the momentum when R is called
        self.beta = 1

        if isinstance(distribution, Distribution):
            distribution.mjhmc = True
            if not distribution.generation_instance:
                distribution.reset()
            self._bind(distribution)
        else:
            raise NotImplementedError(
                ("Unfortunately, you must define your distribution by"
                 " subclassing mjhmc.misc.Distribution."
                 "This is due to subtle issues having to do with generating"
                 " a fair initialization for"
                 "the embedded Markov Chain. See the docs in mjhmc.misc.Distribution."
                ))

    @overrides(HMCBase)
    def sample(self, n_samples=1000, preserve_order=False, num_steps=None):
        """ Runs sampler and returns a list of n_samples (resampled to be fair)

        preserve_order has no effect if resample is enabled (markov_jump_hmc.py:293-338).
        """
        if num_steps is not None:
            n_samples = num_steps
        if not self.resample:
            return self._sample_plain(n_samples, preserve_order)
        eng = self._engine
        # 1 + n iterations; sample k is paired with the dwelling time recorded before iteration k+1
        S, dwell, _ = self._advance(n_samples + 1, want_dwell=True)
        n, N, d = n_samples, self.nbatch, self.ndims
        m = n * N
        if self._group is not None:
            return self._resample_sharded(S, dwell[:n], n)
        dwell_t = dwell[:n].reshape(-1)
        total_t = np.sum(dwell_t.cpu().numpy())
        r = np.sort(np.random.random(m)) * total_t
        self.resample_columns = None
        with eng.ctx():
            out = self._resample_device(S, dwell_t, torch.as_tensor(r, device=eng.device))
            return self._d2h(out).astype(np.float64, copy=False)

    def _resample_device(self, S, dwell_t, r_d):
        """out[:, j] = samples[:, first i with cumsum(dwell_t)[i] > r[j]] (markov_jump_hmc.py:321-328)."""
        eng = self._engine
        d, m, m_out = self.ndims, dwell_t.numel(), r_d.numel()
        out = torch.zeros((d, m_out), dtype=eng.tdtype, device=eng.device)
        if m == 0 or m_out == 0:
            return out
        nbytes = int(eng.lib.mjhmc_resample_scratch_bytes(m))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
        _lib.check(eng.lib.mjhmc_resample(eng.code, d, _device.ptr(dwell_t), m, _device.ptr(r_d), m_out,
                                          _device.ptr(S), S.stride(0), _device.ptr(out), m_out, None,
                                          _device.ptr(scratch), eng._stream()), "resample")
        return out

    def _resample_sharded(self, S, dwell, n):
        """Dwell-time resampling over the WHOLE sharded cloud.  The reference's flat order is
        iteration-major, particle-minor over all particles (np.concatenate(dwell_t_k), :321): one all_gather of
        the per-(iteration, rank) dwell sums gives every local segment its global offset, the sorted uniforms are
        drawn once (rank 0) and shared, and each rank resolves the draws that land in its own segments with the
        single-GPU kernel.  Returns the resampled columns whose source particle lives on this rank, in increasing
        column order; ``self.resample_columns`` holds their positions in the reference's (ndims, n * N) array
        (parallel.allgather_resampled assembles it)."""
        eng = self._engine
        n_local = self.nbatch
        with eng.ctx():
            seg = dwell.sum(dim=1) if n_local else torch.zeros(n, dtype=torch.float64, device=eng.device)
            plan = parallel.resample_plan(seg, n * parallel.allreduce_sum_int(n_local, self._group), self._group)
            r_d = torch.as_tensor(plan["r"], device=eng.device)
            # which draws fall into a segment of this rank: bounds = [A_0, B_0, A_1, B_1, ...] ascending
            bounds = torch.as_tensor(plan["bounds"], device=eng.device)
            pos = torch.searchsorted(bounds, r_d, right=True)
            cols = torch.nonzero(pos % 2 == 1).reshape(-1)
            self.resample_columns = cols.cpu().numpy()
            if n_local == 0 or cols.numel() == 0:
                return np.zeros((self.ndims, 0))
            # fold the dwell mass other ranks own between two local segments into the first local element of the
            # later one: the local prefix sums then ARE the global ones at every local element
            dwell_g = dwell.clone()
            dwell_g[:, 0] += torch.as_tensor(plan["gaps"], device=eng.device)
            out = self._resample_device(S, dwell_g.reshape(-1), r_d[cols].contiguous())
            return self._d2h(out).astype(np.float64, copy=False)


class MarkovJumpHMC(ContinuousTimeHMC):
    """This class implements Markov Jump HMC as described in http://arxiv.org/abs/1509.03808
    """
    _sampler_code = _lib.SAMPLER_MARKOV_JUMP

    @overrides(ContinuousTimeHMC)
    def _on_infinite_rate(self, samples, it, dwell, choice, energy=None):
        # infinite rate due to taking too large of a step (markov_jump_hmc.py:376-389):
        # take smaller steps, but go the same overall distance -- for the whole batch
        self.epsilon *= 0.5
        self.num_leapfrog_steps *= 2
        depth = np.log(self.original_epsilon / self.epsilon) / np.log(2)
        print("Ecountered infinite rate, doubling back. Depth: {}".format(depth))
        self._engine.reset_cache()
        self._run(1, samples, it, dwell, choice, energy)
        # restore the old guys
        self.epsilon *= 2
        self.num_leapfrog_steps = int(self.num_leapfrog_steps / 2)
