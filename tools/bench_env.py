"""Env-step kernel alone at an HBM-bound size (B episodes, BA-200, G distinct graphs, random actions)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

B = int(os.environ.get("ECO_ENV_B", "262144"))
G = int(os.environ.get("ECO_ENV_G", "1024"))
n, T = 200, 400
L = _lib.lib()
gs = engine.GraphSet(bench.ba_graphs(G, n, 4, seed=0))
env = engine.BatchedSpinSystem(gs, B, T, None if os.environ.get("ECO_ENV_NOBASIN") else 1.0 / n)
rng = np.random.default_rng(0)
env.reset(spins=torch.from_numpy((2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8)).cuda())
gen = torch.Generator(device="cuda").manual_seed(7)
acts = [torch.randint(0, n, (B,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(44)]
for a in acts[:4]:
    env.step(a)
torch.cuda.synchronize()
L.eco_profile_enable(1)
for a in acts[4:]:
    env.step(a)
tot, cnt = C.c_double(), C.c_int64()
L.eco_profile_read(1, C.byref(tot), C.byref(cnt))
avg = tot.value / cnt.value / 1000.0
gbs = bench.bytes_env(n) * B / avg / 1e9
print("env_step B=%d: %.1f us/launch, %.1f M env-steps/s, %.0f GB/s algorithmic = %.3f of measured HBM peak" %
      (B, avg * 1e6, B / avg / 1e6, gbs, gbs / bench.peaks()["hbm_gbs"]))
