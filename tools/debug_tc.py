"""Debug helper: tcgen05 vs SIMT Q-values on BA graphs; prints where they differ."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

n = int(os.environ.get("N", "200")); B = int(os.environ.get("B", "8")); T = 2 * n
J = bench.ba_graphs(B, n, 4, seed=0)
gs = engine.GraphSet(J)
w = engine.MPNNWeights(bench.load_weights())
rng = np.random.default_rng(0)
spins = (2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8)
qs = {}
for name, impl in (("simt", _lib.MPNN_SIMT), ("tc", _lib.MPNN_TCGEN05)):
    env = engine.BatchedSpinSystem(gs, B, T, 1.0 / n, mpnn_impl=impl)
    env.reset(spins=spins)
    for rep in range(3):
        q, a = env.q_values(w)
    torch.cuda.synchronize()
    qs[name] = q.cpu().numpy()[:, :n]
d = np.abs(qs["tc"] - qs["simt"])
print("max abs diff", d.max(), "max |q|", np.abs(qs["simt"]).max())
for b in range(min(B, 4)):
    blocks = [d[b, i:i + 16].max() for i in range(0, n, 16)]
    print("episode", b, " per-16-block max diff:", " ".join("%.1e" % x for x in blocks))
