import os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eco_dqn_b200.envs.core as ising_env
from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis, Stopping, RandomErdosRenyiGraphGenerator, EdgeType)
from eco_dqn_b200.networks.mpnn import MPNN
from eco_dqn_b200.agents.dqn.dqn import DQN
from eco_dqn_b200.agents.dqn.utils import TestMetric
n = 40
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
            test_save_path=os.path.join(tmp, "s"), network_save_path=os.path.join(tmp, "n"), n_envs=16)
agent.learn(timesteps=1500)
tr = agent.replay_buffer.sample(64)
def T(f, reps=200):
    for _ in range(20): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("train_step            %.3f ms" % T(lambda: agent.train_step(tr)))
g = tr["graph"]
print("norm_max item         %.3f ms" % T(lambda: float(agent._graphs.gstat[g.long(), 0].max().clamp(min=1).item())))
nm = float(agent._graphs.gstat[g.long(), 0].max().clamp(min=1).item())
print("q_kernel target (cached weights) %.3f ms" % T(lambda: agent._q_kernel(agent.target_network, tr["xn_next"], tr["xg_next"], g, nm)))
def online():
    agent.network._engine_cache = None
    agent._q_kernel(agent.network, tr["xn_next"], tr["xg_next"], g, nm, want_q=False)
print("q_kernel online (rebuild weights) %.3f ms" % T(online))
td = torch.zeros(64, 1, device=agent.device)
print("grad kernel + views   %.3f ms" % T(lambda: agent._grad_kernel(tr["xn"], tr["xg"], g, nm, tr["action"], td)))
print("optimizer.step        %.3f ms" % T(lambda: agent.optimizer.step()))
print("sample minibatch      %.3f ms" % T(lambda: agent.replay_buffer.sample(64)))
