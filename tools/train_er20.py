"""Evidence that DQN.learn learns: ECO-DQN on ER-20 (p = 0.15, +-1) with the reference's hyper-parameters
(experiments/train_eco.py:114-161, 336-345: minibatch 64, update every 32 timesteps, target sync 1000, replay 5000, epsilon
1 -> 0.05 over 150 000 steps, lr 1e-4, gamma 0.95, test every 10 000 timesteps with TestMetric.BEST), acting with 16
lock-step environments on the device.  Test set: the 16 ER-20 validation graphs of tests/golden/graphsets.npz (from the
reference's _graphs/validation, optimal cuts known), one greedy-Q episode per graph from a random start per evaluation.
The reference's own curve (ER_20spin/eco/max_cut/network/training_curve.png) plateaus at a mean best cut of ~10.55 on ITS
50 test graphs; here the yardstick is the mean optimal cut of the 16 validation graphs.

    python tools/train_er20.py [timesteps=600000]  > profiles/r02_training_curve_er20.txt
"""
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eco_dqn_b200.envs.core as ising_env  # noqa: E402
from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,  # noqa: E402
                                     Stopping, RandomErdosRenyiGraphGenerator, SetGraphGenerator, EdgeType)
from eco_dqn_b200.networks.mpnn import MPNN  # noqa: E402
from eco_dqn_b200.agents.dqn.dqn import DQN  # noqa: E402
from eco_dqn_b200.agents.dqn.utils import TestMetric  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
n = 20
gs = np.load(os.path.join(ROOT, "tests", "golden", "graphsets.npz"))
test_graphs = [g.astype(np.float64) for g in gs["er20"]]
opt = gs["er20_opt"].astype(np.float64)
env_args = {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': 1. / n,
            'reversible_spins': True, 'stopping': Stopping.NORMAL}
train_env = ising_env.make("SpinSystem", RandomErdosRenyiGraphGenerator(n, 0.15, EdgeType.DISCRETE), 2 * n, **env_args)
test_env = ising_env.make("SpinSystem", SetGraphGenerator(test_graphs, ordered=True), 2 * n, **env_args)
tmp = tempfile.mkdtemp()
agent = DQN([train_env], lambda: MPNN(), init_weight_std=0.01, double_dqn=True, clip_Q_targets=False, gamma=0.95,
            replay_start_size=500, replay_buffer_size=5000, update_target_frequency=1000, update_learning_rate=False,
            initial_learning_rate=1e-4, peak_learning_rate=1e-4, peak_learning_rate_step=20000, final_learning_rate=1e-4,
            final_learning_rate_step=200000, update_frequency=32, minibatch_size=64, max_grad_norm=None, weight_decay=0,
            update_exploration=True, initial_exploration_rate=1, final_exploration_rate=0.05, final_exploration_step=150000,
            adam_epsilon=1e-8, logging=False, loss="mse", save_network_frequency=10 ** 9,
            network_save_path=os.path.join(tmp, "net"), evaluate=True, test_envs=[test_env], test_episodes=len(test_graphs),
            test_frequency=10000, test_save_path=os.path.join(tmp, "scores"), test_metric=TestMetric.BEST, seed=1, n_envs=16)
untrained = agent.evaluate_agent()[1]
torch.cuda.synchronize()
t0 = time.time()
with open(os.devnull, "w") as dn:
    old, sys.stdout = sys.stdout, dn
    try:
        agent.learn(timesteps=steps)
    finally:
        sys.stdout = old
torch.cuda.synchronize()
wall = time.time() - t0
with open(os.path.join(tmp, "solution.pkl"), "rb") as f:
    curve = pickle.load(f)
print("# ECO-DQN on ER-20, reference hyper-parameters, %d timesteps in %.1f s on one B200 (%.1f us per timestep incl. %d evaluations)"
      % (steps, wall, wall / steps * 1e6, len(curve)))
print("# mean optimal cut of the 16 test graphs: %.3f; untrained network (greedy-Q, random start): %.3f" % (opt.mean(), untrained))
print("# timestep  mean_best_cut  ratio_to_optimal")
for ts, sol in curve:
    print("%9d  %8.3f  %6.3f" % (ts, sol, sol / opt.mean()))
tail = np.array([s for _, s in curve[-10:]])
print("# mean of the last 10 evaluations: %.3f = %.3f of optimal" % (tail.mean(), tail.mean() / opt.mean()))
assert tail.mean() > 0.93 * opt.mean() and tail.mean() > untrained, "the agent did not learn"
