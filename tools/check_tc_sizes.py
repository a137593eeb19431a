"""Quick check of the resident tcgen05 kernel against the CUDA-core kernel at several graph sizes (worst |dq| in units of the
parity tolerance); used while changing the MMA shapes (instruction N trimmed to 8 ceil(N/8))."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eco_dqn_b200.engine as engine
from eco_dqn_b200 import _lib
for n in (200, 196, 185, 130):
    B, G = 300, 10
    J = bench.er_graphs(G, n, 0.15, seed=0)
    gs = engine.GraphSet(J)
    env = engine.BatchedSpinSystem(gs, B, 2*n, 1.0/n, mpnn_impl=_lib.MPNN_TCGEN05)
    w = engine.MPNNWeights(bench.load_weights())
    rng = np.random.default_rng(0)
    env.reset(spins=(2*rng.integers(0,2,size=(B,n))-1).astype(np.int8))
    q_tc,_ = env.q_values(w, impl=_lib.MPNN_TCGEN05, norm_max=-1.0); q_tc=q_tc.clone()
    q_si,_ = env.q_values(w, impl=_lib.MPNN_SIMT, norm_max=-1.0)
    tol = 1e-3*q_si.abs() + 1e-4*q_si.abs().max(1,keepdim=True).values
    err = ((q_tc-q_si).abs()/tol)
    print("n=%d worst %.4f" % (n, float(err.max())))
