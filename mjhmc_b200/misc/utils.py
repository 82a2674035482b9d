"""Small host helpers kept from the reference's misc/utils.py surface.

``draw_from`` and ``min_idx`` (misc/utils.py:15-49) are not host functions here: they are
fused into the transition stage of the CUDA kernels (csrc/common.cuh: decide_*).
"""


def overrides(interface_class):
    """Same contract as the reference decorator (misc/utils.py:5-12)."""
    def overrider(method):
        assert method.__name__ in dir(interface_class)
        return method
    return overrider


def package_path():
    """Directory that holds the package and its data folders (``initializations/``, ``distr_data/``); the reference
    searches sys.path for an entry containing 'MJHMC' (misc/utils.py:60-72)."""
    import os
    return os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
