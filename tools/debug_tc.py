import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mjhmc_b200.misc import distributions as D
from mjhmc_b200.samplers import markov_jump_hmc as S
from oracle import mjhmc_oracle as orc
from tests import helpers
np.set_printoptions(precision=5, linewidth=200, suppress=True)
d, N = int(sys.argv[1]) if len(sys.argv) > 1 else 8, int(sys.argv[2]) if len(sys.argv) > 2 else 40
mode = sys.argv[3] if len(sys.argv) > 3 else "diag"
rs = np.random.RandomState(0)
if mode == "eye":
    J = np.eye(d)
elif mode == "diag":
    J = np.diag(np.arange(1, d + 1) / 4.0)
else:
    A = rs.randn(d, d); J = A.dot(A.T) / d + np.eye(d)
dist = D.Gaussian(ndims=d, nbatch=N, J=J + 1e-12 * (np.arange(d)[:, None] != np.arange(d)[None, :]))
dist._diagonal = False
X0, V0 = rs.randn(d, N), rs.randn(d, N)
helpers.pin_init(dist, X0)
hp = dict(epsilon=0.25, beta=0.3, num_leapfrog_steps=int(sys.argv[4]) if len(sys.argv) > 4 else 1)
s = S.ContinuousTimeHMC(distribution=dist, V=V0, seed=3, dtype="float32", resample=False, **hp)
print("fused", s._engine.fused)
o = orc.OracleSampler("ContinuousTimeHMC", orc.GaussianEnergy(J), X0, V=V0, draws=orc.PhiloxDraws(3), resample=False, **hp)
S_, dw, ch = s._advance(1, want_dwell=True, want_choice=True)
o.sampling_iteration()
X = S_[:, 0, :].double().cpu().numpy()
print("choice gpu", ch[0, :16].cpu().numpy(), " oracle", o.last_choice[:16])
print("dwell gpu", dw[0, :6].cpu().numpy(), " oracle", o.dwelling_times[:6])
print("X gpu   ", X[:4, :6]); print("X oracle", o.X[:4, :6]); print("X0      ", X0[:4, :6])
print("V gpu   ", s.state.V[:4, :6]); print("V oracle", o.V[:4, :6])
print("max err X", np.abs(X - o.X).max(), "V", np.abs(s.state.V - o.V).max())
