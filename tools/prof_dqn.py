"""Where the time of DQN.learn goes (ER-40, reference hyper-parameters of experiments/train_eco.py, one GPU)."""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eco_dqn_b200.envs.core as ising_env  # noqa: E402
from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,  # noqa: E402
                                     Stopping, RandomErdosRenyiGraphGenerator, EdgeType)
from eco_dqn_b200.networks.mpnn import MPNN  # noqa: E402
from eco_dqn_b200.agents.dqn.dqn import DQN  # noqa: E402
from eco_dqn_b200.agents.dqn.utils import TestMetric  # noqa: E402

n = 40
n_envs = int(os.environ.get("N_ENVS", "16"))
env_args = {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': 1. / n,
            'reversible_spins': True, 'stopping': Stopping.NORMAL}
env = ising_env.make("SpinSystem", RandomErdosRenyiGraphGenerator(n, 0.15, EdgeType.DISCRETE), 2 * n, **env_args)
tmp = tempfile.mkdtemp()
agent = DQN([env], lambda: MPNN(), init_weight_std=0.01, double_dqn=True, gamma=0.95, update_learning_rate=False,
            initial_learning_rate=1e-4, minibatch_size=64, update_frequency=32, update_target_frequency=1000,
            replay_start_size=500, replay_buffer_size=5000, final_exploration_step=3000, final_exploration_rate=0.05,
            test_frequency=10 ** 9, save_network_frequency=10 ** 9, logging=False, seed=5, test_metric=TestMetric.BEST,
            test_save_path=os.path.join(tmp, "s"), network_save_path=os.path.join(tmp, "n"), n_envs=n_envs)
acc = {"train_step": 0.0, "n": 0}
orig = agent.train_step


def timed(tr):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = orig(tr)
    torch.cuda.synchronize()
    acc["train_step"] += time.perf_counter() - t0
    acc["n"] += 1
    return out


agent.train_step = timed
steps = int(os.environ.get("STEPS", "6000"))
agent.learn(timesteps=1000)          # warm-up (fills the replay)
acc["train_step"], acc["n"] = 0.0, 0
torch.cuda.synchronize()
t0 = time.perf_counter()
agent.learn(timesteps=steps)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print("n_envs %d: %.1f ms per 1000 timesteps; %d train steps, %.2f ms each = %.0f%% of the time" %
      (n_envs, tot / steps * 1e6, acc["n"], acc["train_step"] / max(acc["n"], 1) * 1e3, 100 * acc["train_step"] / tot))
