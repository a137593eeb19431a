"""One MPNN forward on a large graph set (default ER-500, B=1024) a few times; run under
`ncu --metrics gpu__time_duration.sum` to get the per-kernel split of the large-graph path (mpnn_tcl.cu)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

n = int(os.environ.get("ECO_N", 500))
p = float(os.environ.get("ECO_P", 0.15))
B = int(os.environ.get("ECO_B", 1024))
G = int(os.environ.get("ECO_G", 64))
reps = int(os.environ.get("ECO_REPS", 3))
rng = np.random.default_rng(0)
Js = np.zeros((G, n, n), dtype=np.int8)
for g in range(G):
    up = np.triu(rng.random((n, n)) < p, 1)
    a = (up * np.where(rng.random((n, n)) < 0.5, -1, 1)).astype(np.int8)
    Js[g] = a + a.T
gs = engine.GraphSet(Js)
env = engine.BatchedSpinSystem(gs, B, 2 * n, 1.0 / n, mpnn_impl=_lib.MPNN_TCGEN05)
w = engine.MPNNWeights(bench.load_weights())
env.reset(spins=(2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8))
for _ in range(reps):
    env.q_values(w, impl=_lib.MPNN_TCGEN05)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    env.q_values(w, impl=_lib.MPNN_TCGEN05)
e1.record()
torch.cuda.synchronize()
print("N=%d B=%d forward %.3f ms" % (n, B, e0.elapsed_time(e1) / reps))
