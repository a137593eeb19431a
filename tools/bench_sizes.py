"""Per-step time of [MPNN + argmax, env step] for the BASELINE.json configs' sizes, SIMT vs tcgen05 MPNN."""
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


def er_graphs(count, n, p, rng):
    out = np.zeros((count, n, n), dtype=np.int8)
    for g in range(count):
        up = np.triu(rng.random((n, n)) < p, 1)
        a = (up * np.where(rng.random((n, n)) < 0.5, -1, 1)).astype(np.int8)
        out[g] = a + a.T
    return out


w_all = bench.load_weights()
rng = np.random.default_rng(0)
cases = [("ER-20  (C1) ", 20, 0.15, 100, 5000), ("ER-40  (C5) ", 40, 0.15, 64, 4096), ("ER-60       ", 60, 0.15, 64, 4096),
         ("ER-100      ", 100, 0.15, 64, 4096), ("BA/ER-200   ", 200, 0.04, 256, 4096), ("ER-500 (C3) ", 500, 0.15, 64, 1024),
         ("N=2000 (C4) ", 2000, 0.01, 8, 128)]
for name, n, p, G, B in cases:
    gs = engine.GraphSet(er_graphs(G, n, p, rng))
    for impl_name, impl in (("simt", _lib.MPNN_SIMT), ("tcgen05", _lib.MPNN_TCGEN05)):
        env = engine.BatchedSpinSystem(gs, B, 2 * n, 1.0 / n, mpnn_impl=impl)
        w = engine.MPNNWeights(w_all)
        env.reset(spins=(2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8))
        steps = 10 if n <= 500 else 4
        env.rollout(w, n_steps=3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.rollout(w, n_steps=steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print("%s B=%5d %-8s %8.3f ms/step  %10.0f env-steps/s" % (name, B, impl_name, ms, B / ms * 1e3))
