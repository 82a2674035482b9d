"""Golden trajectories for MultimodalGaussian from the UNMODIFIED reference class
mjhmc.misc.distributions.MultimodalGaussian (distributions.py:314-346), with injected draws.

    python tests/golden/generate_multimodal_golden.py     (build container only: needs /root/reference)

Reuses the machinery of generate_golden.py (same shims, same Injector); writes inject_<Sampler>_multimodal3.npz.
separation = 1 keeps exp(4 s.x) inside the float32 range so the fp32 parity tests can use the same fixtures.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import generate_golden as gg          # noqa: E402  (imports the reference with the shims)


def main():
    hp = dict(epsilon=0.35, beta=0.3, num_leapfrog_steps=4)
    for k, sname in enumerate(("MarkovJumpHMC", "ContinuousTimeHMC", "ControlHMC")):
        np.random.seed(20 + k)
        dobj = gg.dist.MultimodalGaussian(ndims=3, nbatch=72, separation=1)
        arrs = gg.run_injected(sname, dobj, hp, 12, 200 + k)
        arrs["dist"] = np.array("MultimodalGaussian")
        arrs["dist_params"] = np.array([1.0])
        np.savez_compressed(os.path.join(HERE, "inject_%s_multimodal3.npz" % sname), **arrs)
        print(sname, arrs["counters"][-1], "attempts", arrs["attempts"][-1], "finite", np.isfinite(arrs["X"]).all())


if __name__ == "__main__":
    main()
