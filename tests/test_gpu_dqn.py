"""GPU tests of the DQN trainer: one update step against the reference's train_step, and a short learn() run."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def env_args(n):
    from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,
                                         Stopping)
    return {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': 1. / n,
            'reversible_spins': True, 'stopping': Stopping.NORMAL}


def make_agent(tmp_path, graphs, T, **kw):
    import eco_dqn_b200.envs.core as ising_env
    from eco_dqn_b200.envs.utils import SetGraphGenerator
    from eco_dqn_b200.networks.mpnn import MPNN
    from eco_dqn_b200.agents.dqn.dqn import DQN
    from eco_dqn_b200.agents.dqn.utils import TestMetric
    n = graphs[0].shape[0]
    env = ising_env.make("SpinSystem", SetGraphGenerator([g.astype(np.float64) for g in graphs], ordered=True), T,
                         **env_args(n))
    args = dict(double_dqn=True, gamma=0.95, update_learning_rate=False, initial_learning_rate=1e-4, minibatch_size=16,
                replay_buffer_size=400, replay_start_size=64, logging=False, seed=3, test_metric=TestMetric.BEST,
                test_save_path=str(tmp_path / "scores"), network_save_path=str(tmp_path / "net"), n_envs=4)
    args.update(kw)
    return DQN([env], lambda: MPNN(), **args)


def test_train_step_matches_reference(tmp_path):
    z = np.load(os.path.join(GOLDEN, "dqn_er40.npz"))
    graphs = list(z["graphs"])
    agent = make_agent(tmp_path, graphs, 10)
    keys = [k[3:] for k in z.files if k.startswith("w::")]
    agent.network.load_state_dict({k: torch.tensor(z["w::" + k]) for k in keys})
    agent.target_network.load_state_dict({k: torch.tensor(z["wt::" + k]) for k in keys})
    slots = agent._write_ring([g.astype(np.float64) for g in graphs])
    NP = agent._env.NP
    dev = agent.device

    def feats(rows):
        xn = np.zeros((rows.shape[0], 3, NP), dtype=np.float32)
        xn[:, :, :40] = rows[:, :3]
        return torch.tensor(xn, device=dev), torch.tensor(np.ascontiguousarray(rows[:, 3:7, 0]), device=dev)

    xn, xg = feats(z["rows"])
    xn2, xg2 = feats(z["rows_next"])
    trans = dict(xn=xn, xg=xg, xn_next=xn2, xg_next=xg2, action=torch.tensor(z["actions"], device=dev),
                 reward=torch.tensor(z["rewards"], device=dev), done=torch.tensor(z["dones"], device=dev),
                 graph=torch.tensor(slots[z["graph_idx"]], device=dev))
    loss = agent.train_step(trans)
    assert abs(loss - float(z["loss"])) <= 1e-4 * abs(float(z["loss"])), (loss, float(z["loss"]))
    for k, p in agent.network.named_parameters():
        g, ref = p.grad.cpu().numpy(), z["g::" + k]
        assert np.allclose(g, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max() + 1e-9), k
        w1 = p.detach().cpu().numpy()
        assert np.allclose(w1, z["w1::" + k], rtol=0, atol=2.5e-5), k        # Adam step of 1e-4 per element at most
        assert not np.array_equal(w1, z["w::" + k])


def test_learn_runs_and_checkpoints(tmp_path):
    gs = np.load(os.path.join(GOLDEN, "graphsets.npz"))
    graphs = list(gs["er20"][:6])
    agent = make_agent(tmp_path, graphs, 40, n_envs=8, update_frequency=8, update_target_frequency=160,
                       replay_start_size=160, replay_buffer_size=1000, test_frequency=320, test_episodes=6,
                       save_network_frequency=480, init_weight_std=0.01, final_exploration_step=600,
                       final_exploration_rate=0.05)
    before = {k: v.clone() for k, v in agent.network.state_dict().items()}
    losses = agent.learn(timesteps=8 * 40 * 3)
    assert len(losses) == (960 - 160) // 8 and all(np.isfinite(l) for _, l in losses)
    assert any(not torch.equal(before[k], v) for k, v in agent.network.state_dict().items())
    assert len(agent.replay_buffer) == 960 and agent.epsilon == 0.05
    assert os.path.exists(str(tmp_path / "net_best.pth")) and os.path.exists(str(tmp_path / "net480.pth"))
    assert os.path.exists(str(tmp_path / "scores.pkl")) and os.path.exists(str(tmp_path / "losses.pkl"))
    sd = torch.load(str(tmp_path / "net_best.pth"), map_location="cpu")
    from oracle.mpnn import KEYS
    assert tuple(sd.keys()) == KEYS                    # reference checkpoint format
    score, sol = agent.evaluate_agent()
    assert np.isfinite(score) and 0 < sol <= gs["er20_opt"][:6].max()
    # replay content: features are the env's observations, graphs point into the ring
    rb = agent.replay_buffer
    assert float(rb.done.sum()) == 960 / 40 and int(rb.graph.max()) < agent._ring_size
    assert torch.isfinite(rb.reward).all() and set(rb.xn[:960, 0, :20].unique().tolist()) <= {-1.0, 1.0}
