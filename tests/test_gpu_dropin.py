"""GPU tests of the reference-facing Python surface: make / SpinSystem facade, solvers, test_network."""
import os
from copy import deepcopy

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def env_args_for(z):
    from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,
                                         Stopping)
    b = float(z["basin_reward"])
    return {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None,
            'basin_reward': None if b < 0 else b, 'reversible_spins': True, 'stopping': Stopping.NORMAL}


def network_for(z):
    from eco_dqn_b200.networks.mpnn import MPNN
    from oracle.mpnn import weights_from_npz
    net = MPNN()
    net.load_state_dict({k: torch.tensor(v) for k, v in weights_from_npz(z).items()})
    return net.cuda().eval()


@pytest.mark.parametrize("name", ["er20_g0", "er20_g1", "er20_g3_nobasin", "er40_g0", "ba40u_g0"])
def test_test_network_matches_reference_frames(name):
    """Same seed -> same random starts; result frames equal the reference's own test_network output."""
    from eco_dqn_b200.experiments.utils import test_network
    z = load(name)
    n_attempts = z["init_spins"].shape[0]
    np.random.seed(int(z["seed"]))
    res, raw, hist = test_network(network_for(z), env_args_for(z), [z["J"].astype(np.float64)], "cuda", 2,
                                  n_attempts=n_attempts, return_raw=True, return_history=True)
    assert list(res.columns) == ["cut", "sol", "mean cut", "greedy (+1 init) cut", "greedy (+1 init) sol",
                                 "greedy (rand init) cut", "greedy (rand init) sol", "greedy (rand init) mean cut", "time"]
    assert np.array_equal(np.array(raw["init spins"][0]), z["init_spins"])
    assert res["greedy (+1 init) cut"][0] == float(z["greedy_single_cut"])
    assert np.array_equal(res["greedy (+1 init) sol"][0], z["greedy_single_spins"])
    assert res["greedy (rand init) cut"][0] == float(z["res_greedy_rand_cut"])
    assert res["greedy (rand init) mean cut"][0] == float(z["res_greedy_rand_mean_cut"])
    assert np.array_equal(np.array(raw["greedy cuts"][0]), z["greedy_cuts"])
    acts = np.array([row[1:] for row in hist["actions"][0]])
    # a trajectory may leave the reference's only at a step where the reference's own top Q-values are a near-tie
    # (checked against the oracle's Q at the first differing step); every identical one must agree exactly
    from test_gpu_round2 import assert_divergence_only_at_near_ties
    from oracle.mpnn import weights_from_npz
    b = float(z["basin_reward"])
    same = assert_divergence_only_at_near_ties(z["J"], weights_from_npz(z), z["init_spins"], z["actions"], acts,
                                               None if b < 0 else b)
    assert same.mean() >= 0.5, "too few identical trajectories: %s" % same
    rews = np.array([row[1:] for row in hist["rewards"][0]], dtype=np.float64)
    assert np.array_equal(rews[same].view(np.uint64), z["rewards"][same].view(np.uint64))
    assert np.array_equal(np.array(hist["scores"][0])[same], z["scores"][same])
    assert np.array_equal(np.array(raw["cuts"][0])[same], z["best_cut"][same])
    if same.all():
        assert res["cut"][0] == float(z["res_cut"]) and res["mean cut"][0] == float(z["res_mean_cut"])
        assert np.array_equal(np.array(raw["sols"][0]), z["best_spins"])


def test_test_network_groups_graphs_and_respects_rng_order():
    from eco_dqn_b200.experiments.utils import test_network
    gs = np.load(os.path.join(GOLDEN, "graphsets.npz"))
    z = load("er20_g0")
    graphs = [g.astype(np.float64) for g in gs["er20"][:5]]
    np.random.seed(5)
    res, raw = test_network(network_for(z), env_args_for(z), graphs, "cuda", 2, n_attempts=6, return_raw=True)
    # expected initial spins: the reference draws N for the constructor's reset, then N per episode, graph by graph
    np.random.seed(5)
    for j in range(5):
        np.random.randint(2, size=20)
        want = np.stack([2 * np.random.randint(2, size=20) - 1 for _ in range(6)])
        assert np.array_equal(np.array(raw["init spins"][j]), want)
    assert (res["cut"] <= gs["er20_opt"][:5]).all() and (res["cut"] >= res["greedy (rand init) mean cut"] - 20).all()
    assert len(res) == 5


def test_spin_system_facade_follows_reference_step_by_step():
    import eco_dqn_b200.envs.core as ising_env
    from eco_dqn_b200.envs.utils import SingleGraphGenerator
    from eco_dqn_b200.agents.solver import Greedy, Network
    z = load("er20_g0")
    J, T, n = z["J"].astype(np.float64), int(z["T"]), int(z["n"])
    np.random.seed(int(z["seed"]))
    env = ising_env.make("SpinSystem", SingleGraphGenerator(J), T, **env_args_for(z))
    assert env.n_spins == n and env.action_space.n == n and env.observation_space.shape == [n, 7]
    assert env.scorer._max_local_reward == float(z["mlr"]) and env.scorer._lower_bound == float(z["lb"])
    assert env.scorer._solution_quality_normalizer == float(z["qn"])
    obs = env.reset()                                   # same RNG position as the reference's first episode
    assert np.array_equal(env.state[0], z["init_spins"][0])
    assert obs.shape == (7 + n, n) and np.array_equal(obs[7:], J)
    genv = deepcopy(env)
    k = list(z["obs_steps"])
    for t in range(T):
        if t in k:
            assert np.array_equal(obs[:7].astype(np.float32), z["obs"][0, k.index(t)])
        obs, rew, done, info = env.step(int(z["actions"][0, t]))
        assert np.float64(rew).view(np.uint64) == z["rewards"][0, t].view(np.uint64)
        assert env.score == z["scores"][0, t + 1] and done == bool(z["dones"][0, t]) and info is None
    assert env.best_solution == z["best_cut"][0] and np.array_equal(env.best_spins, z["best_spins"][0])
    assert env.current_step == T
    with pytest.raises(NotImplementedError):
        env.step(0)
    # the deep copy taken at reset is an independent episode: greedy from the same start
    Greedy(genv).solve()
    assert genv.best_solution == z["greedy_cuts"][0] and np.array_equal(genv.best_spins, z["greedy_spins"][0])
    assert genv.current_step == z["greedy_steps"][0]
    # Network solver, one env: follows the reference's action sequence for this episode
    net_env = ising_env.make("SpinSystem", SingleGraphGenerator(J), T, **env_args_for(z))
    agent = Network(network_for(z), net_env)
    agent.reset(spins=z["init_spins"][1])
    agent.solve()
    assert [int(h[0]) for h in agent.history[1:]] == z["actions"][1].tolist()
    assert net_env.best_solution == z["best_cut"][1]
    with pytest.raises(Exception):
        env.reset(spins=np.zeros(n))


def test_make_rejects_configurations_outside_the_path():
    import eco_dqn_b200.envs.core as ising_env
    from eco_dqn_b200.envs.utils import SingleGraphGenerator, OptimisationTarget, SpinBasis
    z = load("er20_g0")
    gg = SingleGraphGenerator(z["J"].astype(np.float64))
    with pytest.raises(NotImplementedError):
        ising_env.make("SpinSystem", gg, 40)                      # factory defaults: ENERGY target (score_solver.py:885)
    for key, bad in (("optimisation_target", OptimisationTarget.MAX_CLIQUE), ("spin_basis", SpinBasis.BINARY)):
        with pytest.raises(NotImplementedError):
            ising_env.make("SpinSystem", gg, 40, **dict(env_args_for(z), **{key: bad}))
    with pytest.raises(NotImplementedError):
        ising_env.make("Other")


# ---------------------------------------------------------------------------------------------------------------
# S2V-DQN configuration through the reference-facing surface (experiments/pretrained_agent/test_s2v.py, SIGNED basis)
# ---------------------------------------------------------------------------------------------------------------
def s2v_env_args():
    from eco_dqn_b200.envs.utils import Observable, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis
    return {'observables': [Observable.SPIN_STATE], 'reward_signal': RewardSignal.DENSE, 'extra_action': ExtraAction.NONE,
            'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
            'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': None,
            'reversible_spins': False}


def s2v_network_for(z):
    from eco_dqn_b200.networks.mpnn import MPNN
    from oracle.mpnn import weights_from_npz
    net = MPNN(n_obs_in=1, n_layers=3, n_features=64, n_hid_readout=[], tied_weights=False)
    net.load_state_dict({k: torch.tensor(v) for k, v in weights_from_npz(z).items()})
    return net.cuda().eval()


@pytest.mark.parametrize("name", ["s2v_er20_g0", "s2v_ba40_g1"])
def test_s2v_test_network_matches_reference(name):
    from eco_dqn_b200.experiments.utils import test_network
    z = load(name)
    res, raw, hist = test_network(s2v_network_for(z), s2v_env_args(), [z["J"].astype(np.float64)], "cuda", 1,
                                  n_attempts=50, return_raw=True, return_history=True)
    assert res["cut"][0] == float(z["best_cut"]) and res["mean cut"][0] == float(z["best_cut"])
    assert res["greedy (+1 init) cut"][0] == float(z["greedy_cut"])
    assert res["greedy (rand init) cut"][0] == float(z["greedy_cut"])          # experiments/utils.py:257-260
    assert np.array_equal(res["sol"][0], z["best_spins"].astype(np.float64))
    assert hist["actions"][0][0][1:] == [int(a) for a in z["actions"]]
    assert np.array_equal(np.array(hist["rewards"][0][0][1:]), z["rewards"])
    assert np.array_equal(np.array(hist["scores"][0][0]), z["scores"])
    assert len(raw["cuts"][0]) == 1 and raw["greedy cuts"][0] == []


def test_s2v_facade_and_solvers_follow_the_reference():
    import eco_dqn_b200.envs.core as ising_env
    from eco_dqn_b200.envs.utils import SingleGraphGenerator
    from eco_dqn_b200.agents.solver import Greedy, Network
    z = load("s2v_er20_g5")
    n = int(z["n"])
    env = ising_env.make("SpinSystem", SingleGraphGenerator(z["J"].astype(np.float64)), n, **s2v_env_args())
    obs = env.reset()
    assert obs.shape == (1 + n, n) and np.array_equal(obs[0], -np.ones(n)) and env.get_allowed_action_states() == -1
    assert env.observation_space.shape[1] == 1
    for t, a in enumerate(z["actions"]):
        obs, r, d, _ = env.step(int(a))
        assert r == z["rewards"][t] and d == bool(z["dones"][t]) and env.score == z["scores"][t + 1]
        assert np.array_equal(obs[0], z["spins"][t + 1].astype(np.float64))
    assert env.best_solution == float(z["best_cut"])
    g = deepcopy(env)
    g.reset(spins=np.array([-1] * n))
    Greedy(g).solve()
    assert g.best_solution == float(z["greedy_cut"]) and np.array_equal(g.best_spins, z["greedy_spins"].astype(np.float64))
    g2 = deepcopy(env)                       # step-by-step greedy (the host decides): same descent
    g2.reset()
    agent = Greedy(g2)
    done = False
    while not done:
        _, done = agent.step()
    assert g2.best_solution == float(z["greedy_cut"])
    e3 = deepcopy(env)
    net_agent = Network(s2v_network_for(z), e3)
    net_agent.reset()
    done = False
    acts = []
    while not done:
        _, done = net_agent.step()
        acts.append(int(net_agent.history[-1][0]))
    assert acts == [int(a) for a in z["actions"]] and e3.best_solution == float(z["best_cut"])


@pytest.mark.parametrize("name", ["mincut_er20_g0", "mincut_ba40u_g1"])
def test_mincut_test_network_matches_reference_frames(name):
    """OptimisationTarget.MIN_CUT through test_network: same seed -> the reference's result frame."""
    from eco_dqn_b200.experiments.utils import test_network
    from eco_dqn_b200.envs.utils import OptimisationTarget
    z = load(name)
    args = dict(env_args_for(z), optimisation_target=OptimisationTarget.MIN_CUT)
    np.random.seed(int(z["seed"]))
    res, raw = test_network(network_for(z), args, [z["J"].astype(np.float64)], "cuda", 2,
                            n_attempts=z["init_spins"].shape[0], return_raw=True)
    assert np.array_equal(np.array(raw["init spins"][0]).astype(np.int8), z["init_spins"])
    assert np.array_equal(np.array(raw["cuts"][0]), z["best_cut"])
    assert res["cut"][0] == float(z["res_cut"]) and res["mean cut"][0] == float(z["res_mean_cut"])
    assert res["greedy (+1 init) cut"][0] == float(z["greedy_single_cut"])
    assert res["greedy (rand init) cut"][0] == float(z["res_greedy_rand_cut"])
    assert res["greedy (rand init) mean cut"][0] == float(z["res_greedy_rand_mean_cut"])
