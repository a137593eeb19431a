"""Data-parallel DQN check (run with torchrun, one process per GPU): every rank trains on its own environments and replay
shard; the gradients are averaged once per update -- dp_mode "peer": inside the Adam kernel, straight from the peers' memory
(eco_dp_adam, CUDA IPC over NVLink); dp_mode "nccl": one NCCL all-reduce of the flat gradient buffer, the mean folded into the
Adam kernel -- and the parameters must stay bit-identical across ranks.  Prints ms per 1000 timesteps and per update.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dqn_dp_check.py
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

import eco_dqn_b200.envs.core as ising_env  # noqa: E402
from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,  # noqa: E402
                                     Stopping, RandomErdosRenyiGraphGenerator, EdgeType)
from eco_dqn_b200.networks.mpnn import MPNN  # noqa: E402
from eco_dqn_b200.agents.dqn.dqn import DQN  # noqa: E402
from eco_dqn_b200.agents.dqn.utils import TestMetric  # noqa: E402

n = 40
env_args = {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': 1. / n,
            'reversible_spins': True, 'stopping': Stopping.NORMAL}
results = {}
for mode, graph in (("peer", True), ("nccl", True), ("nccl", False)):
    env = ising_env.make("SpinSystem", RandomErdosRenyiGraphGenerator(n, 0.15, EdgeType.DISCRETE), 2 * n, **env_args)
    tmp = tempfile.mkdtemp()
    agent = DQN([env], lambda: MPNN(), init_weight_std=0.01, double_dqn=True, gamma=0.95, update_learning_rate=False,
                initial_learning_rate=1e-4, minibatch_size=64, update_frequency=32, update_target_frequency=1000,
                replay_start_size=500, replay_buffer_size=5000, final_exploration_step=3000, final_exploration_rate=0.05,
                test_frequency=10 ** 9, save_network_frequency=10 ** 9, logging=False, seed=5, test_metric=TestMetric.BEST,
                test_save_path=os.path.join(tmp, "s%d" % rank), network_save_path=os.path.join(tmp, "n%d" % rank), n_envs=16,
                dp_mode=mode, cuda_graph=graph)
    acc = {"s": 0.0, "n": 0}
    orig = agent._train_step_device

    def timed(idx, orig=orig, acc=acc):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = orig(idx)
        torch.cuda.synchronize()
        acc["s"] += time.perf_counter() - t0
        acc["n"] += 1
        return out

    agent._train_step_device = timed
    agent.learn(timesteps=16 * 80)            # fills the replay, first updates, graph capture
    acc["s"], acc["n"] = 0.0, 0
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = agent.learn(timesteps=16 * 80 * 4)
    torch.cuda.synchronize()
    per_1000 = (time.perf_counter() - t0) / (16 * 80 * 4) * 1e6
    flat = torch.cat([p.detach().reshape(-1) for p in agent.network.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    err = int(agent.optimizer.err_dev.item())
    first_spins = agent.replay_buffer.xn[:2, 0, :8].cpu().numpy().tolist()
    print("rank %d/%d dp_mode=%s cuda_graph=%s: %d updates, last loss %.4g, params identical across ranks: %s, "
          "%.1f ms per 1000 timesteps, %.3f ms per update, err flag %d; replay differs per rank: %s"
          % (rank, world, mode, graph, len(losses), losses[-1][1], same, per_1000, acc["s"] / max(acc["n"], 1) * 1e3, err,
             first_spins[0]), flush=True)
    assert same and len(losses) > 0 and np.isfinite(losses[-1][1]) and err == 0
    results[(mode, graph)] = flat.clone()
    agent.optimizer.close()
    del agent
    dist.barrier()
# the two exchange paths average the same gradients: same training run up to the summation order of the mean
d = (results[("peer", True)] - results[("nccl", True)]).abs().max().item()
print("rank %d: max |param(peer) - param(nccl)| = %.3g, captured vs eager nccl identical: %s"
      % (rank, d, torch.equal(results[("nccl", True)], results[("nccl", False)])), flush=True)
# (destroy_process_group() can block while captured NCCL work is still referenced by a CUDA graph: leave without it)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
