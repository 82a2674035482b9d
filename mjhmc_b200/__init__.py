"""mjhmc_b200 -- B200-native implementation of MJHMC's particle-parallel sampler loop.

Drop-in for the reference's ``mjhmc.samplers.markov_jump_hmc`` and ``mjhmc.misc.distributions``
(same class names, arguments and attributes); see INTEGRATION.md.  ``install_alias()`` registers
the package under the reference's module names so ``from mjhmc.samplers.markov_jump_hmc import
MarkovJumpHMC`` resolves to this implementation.
"""
import sys

__version__ = "0.1.0"


def install_alias(name="mjhmc"):
    from . import experiments, misc, samplers, search
    from .experiments import spectral
    from .search import objective
    from .misc import autocor, distributions, gen_mj_init, tf_distributions, utils
    from .samplers import hmc_state, markov_jump_hmc
    pkg = sys.modules[__name__]
    sys.modules[name] = pkg
    sys.modules[name + ".misc"] = misc
    sys.modules[name + ".misc.distributions"] = distributions
    sys.modules[name + ".misc.tf_distributions"] = tf_distributions   # Funnel (re-exported), SparseImageCode
    sys.modules[name + ".experiments"] = experiments
    sys.modules[name + ".experiments.spectral"] = spectral
    sys.modules[name + ".misc.utils"] = utils
    sys.modules[name + ".misc.autocor"] = autocor
    sys.modules[name + ".misc.gen_mj_init"] = gen_mj_init
    sys.modules[name + ".samplers"] = samplers
    sys.modules[name + ".search"] = search
    sys.modules[name + ".search.objective"] = objective
    sys.modules[name + ".samplers.markov_jump_hmc"] = markov_jump_hmc
    sys.modules[name + ".samplers.hmc_state"] = hmc_state
    return pkg
