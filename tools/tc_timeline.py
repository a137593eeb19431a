"""Print the per-warp event timeline of mpnn_tc_kernel (CTA 0) for one steady-state episode.
Run with ECO_TC_TIMELINE=1 on a GPU box."""
import os
import sys

import numpy as np
import torch

os.environ["ECO_TC_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import eco_dqn_b200.engine as engine  # noqa: E402
from eco_dqn_b200 import _lib  # noqa: E402

G = B = int(os.environ.get("ECO_TL_B", str(148 * 6)))
n = int(os.environ.get("ECO_TL_N", "200"))     # ECO_TL_N=20: the packed mode (several small graphs per CTA pass)
T = 2 * n
J = bench.ba_graphs(G, n, 4, seed=0) if n >= 100 else bench.er_graphs(G, n, 0.15, seed=0)
gs = engine.GraphSet(J)
env = engine.BatchedSpinSystem(gs, B, T, 1.0 / n, mpnn_impl=_lib.MPNN_TCGEN05)
w = engine.MPNNWeights(bench.load_weights())
rng = np.random.default_rng(0)
env.reset(spins=(2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8))
for _ in range(3):
    env.q_values(w)
torch.cuda.synchronize()
buf = env._scratch[:20 * 1024 * 8].view(torch.int64).cpu().numpy().reshape(20, 1024)
ev = [[(int(x) >> 48, int(x) & 0xFFFFFFFFFFFF) for x in row if x != 0] for row in buf]
# episode boundaries: event 1 starts an episode; take the 3rd episode of the CTA
names = {1: "ep start", 2: "loads issued", 3: "xf sync", 4: "S/D done", 5: "A conv done", 6: "cta sync", 7: "edge MMA done",
         8: "h0 done", 14: "Wef issued", 15: "before wait", 10: "g epi", 11: "grp sync", 12: "Wef done", 13: "e epi", 20: "layer top", 21: "cta sync",
         22: "weights+ahead issued", 23: "A.H done", 30: "agg epi", 31: "grp sync", 32: "issued", 33: "Wm done",
         34: "m epi", 35: "grp sync", 36: "issued", 37: "Wu done", 38: "h epi", 40: "layers done", 41: "ep end", 42: "zero |A| start", 43: "zero |A| done",
         70: "staged inputs landed", 71: "cta sync",
         60: "readout start", 61: "readout: pooled", 62: "readout: W_p pooled", 63: "readout: Q, warp argmax", 64: "readout end", 65: "readout: argmax",
         50: "issuer: S/D signalled", 51: "issuer: A, |A| landed", 52: "issuer: edge MMAs issued",
         53: "issuer: layer signalled", 54: "issuer: A.H issued"}
base = None
for wi in (0, 8, 16, 19):
    first = 1 if wi < 16 else (50 if wi == 16 else 60)
    starts = [i for i, (e, _) in enumerate(ev[wi]) if e == first]
    if not starts:
        continue
    seg = ev[wi][starts[2]:starts[3]] if len(starts) > 3 else ev[wi][starts[-1]:]
    t0 = seg[0][1] if base is None else base
    base = t0
    print("---- warp %d (%s), episode 3 of CTA 0: %d cycles" % (wi, "group %d" % (wi // 8) if wi < 16 else ("issuer" if wi == 16 else "tail"), seg[-1][1] - seg[0][1]))
    prev = t0
    for e, t in seg:
        print("%8d  +%6d  %s" % (t - t0, t - prev, names.get(e, str(e))))
        prev = t
