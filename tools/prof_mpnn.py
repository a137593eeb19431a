"""Profiling driver: a few MPNN forward + env-step launches on the bench workload (BA-200, B=4096, G distinct graphs).
Used under ncu (see profiles/README.md); prints event timings when run plainly."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

G = int(os.environ.get("ECO_PROF_G", "4096"))
B = int(os.environ.get("ECO_PROF_B", "4096"))
IMPL = {"tc": _lib.MPNN_TCGEN05, "simt": _lib.MPNN_SIMT}[os.environ.get("ECO_PROF_IMPL", "tc")]
n, T = 200, 400
J = bench.ba_graphs(G, n, 4, seed=0)
gs = engine.GraphSet(J)
env = engine.BatchedSpinSystem(gs, B, T, 1.0 / n, mpnn_impl=IMPL)
w = engine.MPNNWeights(bench.load_weights())
rng = np.random.default_rng(0)
env.reset(spins=(2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8))
env.rollout(w, n_steps=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
env.rollout(w, n_steps=5)
e1.record()
torch.cuda.synchronize()
print("5 rollout steps: %.3f ms/step" % (e0.elapsed_time(e1) / 5))
