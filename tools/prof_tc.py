"""Inner-loop driver for work on mpnn_tc_kernel: correctness of the tcgen05 kernel against the CUDA-core kernel on the bench
workload's states (worst |dq| in units of the parity tolerance), then ms per [MPNN + env step] over a rollout segment.
ECO_PROF_GRAPHS=er|ba picks ER-200 (p = 0.15) or BA-200 (m = 4) graphs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

G = int(os.environ.get("ECO_PROF_G", "4096"))
B = int(os.environ.get("ECO_PROF_B", "4096"))
n = int(os.environ.get("ECO_PROF_N", "200"))
STEPS = int(os.environ.get("ECO_PROF_STEPS", "40"))
T = 2 * n
kind = os.environ.get("ECO_PROF_GRAPHS", "ba")
J = bench.ba_graphs(G, n, 4, seed=0) if kind == "ba" else bench.er_graphs(G, n, 0.15, seed=0)
gs = engine.GraphSet(J)
env = engine.BatchedSpinSystem(gs, B, T, 1.0 / n, mpnn_impl=_lib.MPNN_TCGEN05)
w = engine.MPNNWeights(bench.load_weights())
rng = np.random.default_rng(0)
env.reset(spins=(2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8))
env.rollout(w, n_steps=7)
worst = 0.0
for nm in (None, -1.0):
    q_tc, a_tc = env.q_values(w, impl=_lib.MPNN_TCGEN05, norm_max=nm)
    q_tc, a_tc = q_tc.clone(), a_tc.clone()
    q_si, _ = env.q_values(w, impl=_lib.MPNN_SIMT, norm_max=nm)
    tol = 1e-3 * q_si.abs() + 1e-4 * q_si.abs().max(1, keepdim=True).values
    worst = max(worst, float(((q_tc - q_si).abs() / tol).max()))
    assert torch.equal(a_tc.long(), q_tc.argmax(1)), "fused argmax differs from argmax of the written Q"
q2, _ = env.q_values(w, impl=_lib.MPNN_TCGEN05, norm_max=-1.0)
assert torch.equal(q2, q_tc), "run-to-run differences"
print("tcgen05 vs CUDA-core kernel: worst |dq| = %.3f of the tolerance (1e-3 |q| + 1e-4 max|q|)%s" %
      (worst, "" if worst <= 1.0 else "   <-- FAIL"))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    env.rollout(w, n_steps=STEPS)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / STEPS)
print("%s-%d B=%d G=%d: %.4f ms per [MPNN + env step] (best of 3 x %d steps)" % (kind, n, B, G, best, STEPS))
