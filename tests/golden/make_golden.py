"""Generate golden vectors by running the UNMODIFIED reference (BetterBelle/eco-dqn) on CPU.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Outputs (committed): tests/golden/*.npz.  Nothing at test time reads /root/reference.

What is recorded ("golden-vector protocol"):
  * the graph (int8), the env config, the pretrained weights used (12 tensors, fp32);
  * per episode: init spins, the greedy-Q action sequence the reference took, per-step reward (fp64),
    per-step score (fp64), final best_solution / best_spins;
  * for the first few episodes: observation rows 0..6 cast to fp32 exactly as the reference's driver
    casts them (experiments/utils.py:174), and the reference MPNN's Q-values on those observations;
  * the reference's own `test_network` result frames for the same seed (cut / greedy columns).

The driver loop below mirrors experiments/utils.py:125-207 but calls only reference objects
(`make`, `env.reset/step`, `MPNN.forward`); it additionally checks itself against the reference's
own `test_network` output for the same seed so the recorded trajectories are the reference's.
"""
import os
import sys
import types
import pickle
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# docplex is absent and only used by CplexSolver (src/agents/solver.py:9) -> stub it.
for name in ("docplex", "docplex.mp", "docplex.mp.model"):
    m = types.ModuleType(name)
    m.Model = object
    sys.modules[name] = m
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings("ignore")

import torch  # noqa: E402

import src.envs.core as ising_env  # noqa: E402
from src.envs.utils import (SingleGraphGenerator, DEFAULT_OBSERVABLES, RewardSignal, ExtraAction,  # noqa: E402
                            OptimisationTarget, SpinBasis, Stopping)
from src.networks.mpnn import MPNN  # noqa: E402
from src.agents.solver import Greedy  # noqa: E402
from experiments.utils import test_network, load_graph_set  # noqa: E402

torch.set_num_threads(8)


def env_args_for(n, basin=True, target=OptimisationTarget.CUT):
    return {'observables': DEFAULT_OBSERVABLES,
            'reward_signal': RewardSignal.BLS,
            'extra_action': ExtraAction.NONE,
            'optimisation_target': target,
            'spin_basis': SpinBasis.SIGNED,
            'norm_rewards': True,
            'memory_length': None,
            'horizon_length': None,
            'stag_punishment': None,
            'basin_reward': (1. / n) if basin else None,
            'reversible_spins': True,
            'stopping': Stopping.NORMAL}


def load_net(path):
    net = MPNN(n_obs_in=7, n_layers=3, n_features=64, n_hid_readout=[], tied_weights=False)
    sd = torch.load(path, map_location="cpu")
    net.load_state_dict(sd)
    net.eval()
    for p in net.parameters():
        p.requires_grad = False
    return net, {k: v.numpy().astype(np.float32) for k, v in sd.items()}


def run_case(name, graph, net, weights, seed, n_attempts, n_obs_eps, obs_steps, basin=True, target=OptimisationTarget.CUT):
    """Reference rollout on one graph: B=n_attempts episodes, greedy-Q, T=2N steps."""
    from copy import deepcopy
    n = graph.shape[0]
    T = 2 * n
    env_args = env_args_for(n, basin, target)

    # --- 1. the reference's own test_network for this seed ------------------------------------
    np.random.seed(seed)
    with open(os.devnull, "w") as devnull:
        old = sys.stdout
        sys.stdout = devnull
        try:
            res, raw, hist = test_network(net, env_args, [graph], "cpu", 2, n_attempts=n_attempts,
                                          return_raw=True, return_history=True)
        finally:
            sys.stdout = old

    # --- 2. our instrumented driver, same seed -> must reproduce (1) ---------------------------
    np.random.seed(seed)
    test_env = ising_env.make("SpinSystem", SingleGraphGenerator(graph), T, **env_args)
    g_env = deepcopy(test_env)
    g_env.reset(spins=np.array([-1] * n))
    Greedy(g_env).solve()
    greedy_single_cut = g_env.best_solution
    greedy_single_spins = np.array(g_env.best_spins)

    envs, greedy_envs, obs = [], [], []
    for _ in range(n_attempts):
        e = deepcopy(test_env)
        obs.append(e.reset())
        envs.append(e)
        greedy_envs.append(deepcopy(e))
    init_spins = np.stack([e.best_spins.copy() for e in envs]).astype(np.int8)
    init_score = np.array([e.score for e in envs], dtype=np.float64)
    init_cut = np.array([e.best_solution for e in envs], dtype=np.float64)

    actions = np.zeros((n_attempts, T), dtype=np.int32)
    rewards = np.zeros((n_attempts, T), dtype=np.float64)
    scores = np.zeros((n_attempts, T + 1), dtype=np.float64)
    scores[:, 0] = init_score
    dones = np.zeros((n_attempts, T), dtype=np.uint8)
    best_scores = np.zeros((n_attempts, T + 1), dtype=np.float64)
    best_scores[:, 0] = init_score
    obs_steps = sorted(set(s for s in obs_steps if s <= T))
    obs_rec = np.zeros((n_obs_eps, len(obs_steps), 7, n), dtype=np.float32)
    q_rec = np.zeros((n_obs_eps, len(obs_steps), n), dtype=np.float32)
    for t in range(T):
        ob = torch.FloatTensor(np.array(obs))           # experiments/utils.py:174
        if t in obs_steps:
            obs_rec[:, obs_steps.index(t)] = ob[:n_obs_eps, :7, :].numpy()
        qs = net(ob.clone())                            # forward transposes its argument in place
        if t in obs_steps:
            q_rec[:, obs_steps.index(t)] = qs[:n_obs_eps].numpy()
        acts = qs.argmax(1, True).squeeze(1).numpy()    # experiments/utils.py:65
        obs = []
        for i, (e, a) in enumerate(zip(envs, acts)):
            o, r, d, _ = e.step(a)
            actions[i, t] = a
            rewards[i, t] = r
            scores[i, t + 1] = e.score
            best_scores[i, t + 1] = e.best_score
            dones[i, t] = d
            obs.append(o)
    if T in obs_steps:
        ob = torch.FloatTensor(np.array(obs))
        obs_rec[:, obs_steps.index(T)] = ob[:n_obs_eps, :7, :].numpy()
        q_rec[:, obs_steps.index(T)] = net(ob.clone())[:n_obs_eps].numpy()
    best_cut = np.array([e.best_solution for e in envs], dtype=np.float64)
    best_spins = np.stack([e.best_spins for e in envs]).astype(np.int8)
    final_spins = np.stack([e.state[0, :n] for e in envs]).astype(np.int8)

    greedy_cuts, greedy_spins, greedy_steps = [], [], []
    for e in greedy_envs:
        Greedy(e).solve()
        greedy_cuts.append(e.best_solution)
        greedy_spins.append(np.array(e.best_spins))
        greedy_steps.append(e.current_step)

    # --- 3. cross-check driver (2) against the reference's own loop (1) -----------------------
    assert np.array_equal(np.array(raw["init spins"][0]).astype(np.int8), init_spins)
    assert np.array_equal(np.array(raw["cuts"][0], dtype=np.float64), best_cut)
    assert np.array_equal(np.array(raw["greedy cuts"][0], dtype=np.float64), np.array(greedy_cuts))
    h_actions = np.array([[int(a) for a in row[1:]] for row in hist["actions"][0]])
    assert np.array_equal(h_actions, actions), "driver diverged from reference test_network"
    h_rewards = np.array([[float(x) for x in row[1:]] for row in hist["rewards"][0]])
    assert np.array_equal(h_rewards, rewards)
    assert res["greedy (+1 init) cut"][0] == greedy_single_cut

    sc = test_env.scorer
    out = dict(
        J=graph.astype(np.int8), n=np.int32(n), T=np.int32(T), seed=np.int32(seed), min_cut=np.int32(target == OptimisationTarget.MIN_CUT),
        basin_reward=np.float64(1. / n if basin else -1.0),
        mlr=np.float64(sc._max_local_reward), qn=np.float64(sc._solution_quality_normalizer),
        lb=np.float64(sc._lower_bound),
        init_spins=init_spins, init_score=init_score, init_cut=init_cut,
        actions=actions, rewards=rewards, scores=scores, best_scores=best_scores, dones=dones,
        best_cut=best_cut, best_spins=best_spins, final_spins=final_spins,
        obs_steps=np.array(obs_steps, dtype=np.int32), obs=obs_rec, q=q_rec,
        greedy_single_cut=np.float64(greedy_single_cut), greedy_single_spins=greedy_single_spins.astype(np.int8),
        greedy_cuts=np.array(greedy_cuts, dtype=np.float64), greedy_spins=np.stack(greedy_spins).astype(np.int8),
        greedy_steps=np.array(greedy_steps, dtype=np.int32),
        res_cut=np.float64(res["cut"][0]), res_mean_cut=np.float64(res["mean cut"][0]),
        res_greedy_rand_cut=np.float64(res["greedy (rand init) cut"][0]),
        res_greedy_rand_mean_cut=np.float64(res["greedy (rand init) mean cut"][0]),
    )
    for k, v in weights.items():
        out["w::" + k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f kB): best cuts %s" % (path, os.path.getsize(path) / 1e3, best_cut[:6]))


def run_case_s2v(name, graph, net_path):
    """S2V-DQN configuration (reference experiments/pretrained_agent/test_s2v.py, with SpinBasis.SIGNED because the
    BINARY basis is broken in the reference's own driver): irreversible spins starting at -1, one observable (the spin),
    dense reward, masked argmax (experiments/utils.py:67-74), T = N steps, a single attempt."""
    from copy import deepcopy
    from src.envs.utils import Observable
    n = graph.shape[0]
    T = n
    env_args = {'observables': [Observable.SPIN_STATE], 'reward_signal': RewardSignal.DENSE,
                'extra_action': ExtraAction.NONE, 'optimisation_target': OptimisationTarget.CUT,
                'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True, 'memory_length': None, 'horizon_length': None,
                'stag_punishment': None, 'basin_reward': None, 'reversible_spins': False}
    net = MPNN(n_obs_in=1, n_layers=3, n_features=64, n_hid_readout=[], tied_weights=False)
    sd = torch.load(net_path, map_location="cpu")
    net.load_state_dict(sd)
    net.eval()
    for p in net.parameters():
        p.requires_grad = False
    weights = {k: v.numpy().astype(np.float32) for k, v in sd.items()}
    with open(os.devnull, "w") as devnull:
        old = sys.stdout
        sys.stdout = devnull
        try:
            res, raw, hist = test_network(net, env_args, [graph], "cpu", 1, n_attempts=50, return_raw=True,
                                          return_history=True)
        finally:
            sys.stdout = old
    env = ising_env.make("SpinSystem", SingleGraphGenerator(graph), T, **env_args)
    g_env = deepcopy(env)
    g_env.reset(spins=np.array([-1] * n))
    Greedy(g_env).solve()
    obs = env.reset()
    init_score = env.score
    actions, rewards, scores, dones, qs_rec, spins_rec = [], [], [env.score], [], [], [obs[0].copy()]
    done = False
    while not done:
        ob = torch.FloatTensor(np.array([obs]))
        qs = net(ob)                                     # transposes `ob` in place (mpnn.py:44)
        qs_rec.append(qs[0].numpy().copy() if qs.dim() == 2 else qs.numpy().copy())
        q2 = qs if qs.dim() == 2 else qs[None]
        mask = (ob[:, :, 0] != -1)                       # experiments/utils.py:71-73
        a = int(q2.masked_fill(mask, -1000).argmax(1, True).squeeze(1).numpy()[0])
        obs, r, done, _ = env.step(a)
        actions.append(a); rewards.append(r); scores.append(env.score); dones.append(done)
        spins_rec.append(obs[0].copy())
    assert [int(a) for a in hist["actions"][0][0][1:]] == actions, "driver diverged from reference test_network"
    assert np.array_equal(np.array([float(x) for x in hist["rewards"][0][0][1:]]), np.array(rewards))
    assert res["cut"][0] == env.best_solution and res["greedy (+1 init) cut"][0] == g_env.best_solution
    sc = env.scorer
    out = dict(J=graph.astype(np.int8), n=np.int32(n), T=np.int32(T),
               mlr=np.float64(sc._max_local_reward), qn=np.float64(sc._solution_quality_normalizer),
               lb=np.float64(sc._lower_bound), init_score=np.float64(init_score),
               actions=np.array(actions, dtype=np.int32), rewards=np.array(rewards, dtype=np.float64),
               scores=np.array(scores, dtype=np.float64), dones=np.array(dones, dtype=np.uint8),
               q=np.stack(qs_rec).astype(np.float32), spins=np.stack(spins_rec).astype(np.int8),
               best_cut=np.float64(env.best_solution), best_spins=np.array(env.best_spins).astype(np.int8),
               greedy_cut=np.float64(g_env.best_solution), greedy_spins=np.array(g_env.best_spins).astype(np.int8),
               greedy_steps=np.int32(g_env.current_step))
    for k, v in weights.items():
        out["w::" + k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d steps, best cut %s, greedy %s" % (path, len(actions), env.best_solution, g_env.best_solution))


def main_mincut():
    """ECO-DQN configuration with OptimisationTarget.MIN_CUT (score_solver.py:423-505): same driver, same networks."""
    nets = os.path.join(REF, "experiments/pretrained_agent/networks/eco")
    val = os.path.join(REF, "_graphs/validation")
    with open(os.devnull, "w") as dn:
        old = sys.stdout
        sys.stdout = dn
        try:
            er20 = load_graph_set(os.path.join(val, "ER_20spin_p15_100graphs.pkl"))
            ba40 = load_graph_set(os.path.join(val, "BA_40spin_m4_100graphs.pkl"))
            bau = load_graph_set(os.path.join(val, "BA_40spin_m4_uniform_100graphs.pkl"))
        finally:
            sys.stdout = old
    net20, w20 = load_net(os.path.join(nets, "network_best_ER_20spin.pth"))
    netb40, wb40 = load_net(os.path.join(nets, "network_best_BA_40spin.pth"))
    run_case("mincut_er20_g0", er20[0], net20, w20, seed=300, n_attempts=6, n_obs_eps=3, obs_steps=range(0, 41),
             target=OptimisationTarget.MIN_CUT)
    run_case("mincut_ba40_g2", ba40[2], netb40, wb40, seed=301, n_attempts=4, n_obs_eps=2, obs_steps=range(0, 81, 4),
             target=OptimisationTarget.MIN_CUT)
    run_case("mincut_ba40u_g1", bau[1], netb40, wb40, seed=302, n_attempts=4, n_obs_eps=2, obs_steps=range(0, 81, 4),
             target=OptimisationTarget.MIN_CUT)      # all weights +1: negative max_local_reward, quality normaliser 1


def main_s2v():
    s2v = os.path.join(REF, "experiments/pretrained_agent/networks/s2v")
    val = os.path.join(REF, "_graphs/validation")
    with open(os.devnull, "w") as dn:
        old = sys.stdout
        sys.stdout = dn
        try:
            er20 = load_graph_set(os.path.join(val, "ER_20spin_p15_100graphs.pkl"))
            ba40 = load_graph_set(os.path.join(val, "BA_40spin_m4_100graphs.pkl"))
            er200 = load_graph_set(os.path.join(REF, "_graphs/testing/ER_200spin_p15_50graphs.pkl"))
        finally:
            sys.stdout = old
    run_case_s2v("s2v_er20_g0", er20[0], os.path.join(s2v, "network_best_ER_20spin.pth"))
    run_case_s2v("s2v_er20_g5", er20[5], os.path.join(s2v, "network_best_ER_20spin.pth"))
    run_case_s2v("s2v_ba40_g1", ba40[1], os.path.join(s2v, "network_best_BA_40spin.pth"))
    run_case_s2v("s2v_er200_g0", er200[0], os.path.join(s2v, "network_best_ER_200spin.pth"))


def main():
    nets = os.path.join(REF, "experiments/pretrained_agent/networks/eco")
    val = os.path.join(REF, "_graphs/validation")
    tst = os.path.join(REF, "_graphs/testing")

    def graphs(p):
        with open(os.devnull, "w") as dn:
            old = sys.stdout
            sys.stdout = dn
            try:
                return load_graph_set(p)
            finally:
                sys.stdout = old

    er20 = graphs(os.path.join(val, "ER_20spin_p15_100graphs.pkl"))
    net20, w20 = load_net(os.path.join(nets, "network_best_ER_20spin.pth"))
    for gi in (0, 1, 7):
        run_case("er20_g%d" % gi, er20[gi], net20, w20, seed=100 + gi, n_attempts=8, n_obs_eps=4,
                 obs_steps=range(0, 41))
    run_case("er20_g3_nobasin", er20[3], net20, w20, seed=7, n_attempts=6, n_obs_eps=2,
             obs_steps=range(0, 41), basin=False)

    er40 = graphs(os.path.join(val, "ER_40spin_p15_100graphs.pkl"))
    net40, w40 = load_net(os.path.join(nets, "network_best_ER_40spin.pth"))
    run_case("er40_g0", er40[0], net40, w40, seed=40, n_attempts=6, n_obs_eps=2, obs_steps=range(0, 81, 2))

    bau = graphs(os.path.join(val, "BA_40spin_m4_uniform_100graphs.pkl"))
    netb40, wb40 = load_net(os.path.join(nets, "network_best_BA_40spin.pth"))
    run_case("ba40u_g0", bau[0], netb40, wb40, seed=41, n_attempts=4, n_obs_eps=2, obs_steps=range(0, 81, 4))

    ba60 = graphs(os.path.join(val, "BA_60spin_m4_100graphs.pkl"))
    netb60, wb60 = load_net(os.path.join(nets, "network_best_BA_60spin.pth"))
    run_case("ba60_g2", ba60[2], netb60, wb60, seed=60, n_attempts=4, n_obs_eps=2, obs_steps=range(0, 121, 8))

    er200 = graphs(os.path.join(tst, "ER_200spin_p15_50graphs.pkl"))
    net200, w200 = load_net(os.path.join(nets, "network_best_ER_200spin.pth"))
    run_case("er200_g0", er200[0], net200, w200, seed=200, n_attempts=4, n_obs_eps=2,
             obs_steps=list(range(0, 12)) + list(range(20, 401, 20)))

    ba200 = graphs(os.path.join(val, "BA_200spin_m4_100graphs.pkl"))
    netb200, wb200 = load_net(os.path.join(nets, "network_best_BA_200spin.pth"))
    run_case("ba200_g0", ba200[0], netb200, wb200, seed=201, n_attempts=4, n_obs_eps=2,
             obs_steps=list(range(0, 12)) + list(range(20, 401, 20)))

    # A handful of extra BA-200 validation graphs (int8) for the multi-graph GPU tests and known-answer
    # upper bounds (README.md:82: opts/cuts_* are best-known cut values).
    with open(os.path.join(val, "opts/cuts_BA_200spin_m4_100graphs.pkl"), "rb") as f:
        cuts200 = np.array(pickle.load(f), dtype=np.float64)
    with open(os.path.join(val, "opts/cuts_ER_20spin_p15_100graphs.pkl"), "rb") as f:
        cuts20 = np.array(pickle.load(f), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "graphsets.npz"),
                        ba200=np.stack(ba200[:8]).astype(np.int8), ba200_opt=cuts200[:8],
                        er20=np.stack(er20[:16]).astype(np.int8), er20_opt=cuts20[:16])
    print("wrote graphsets.npz")

    # Graph generators (reference src/envs/utils.py:165-236): same seeds -> same graphs in the host-side mirror.
    import random
    from src.envs.utils import RandomErdosRenyiGraphGenerator, RandomBarabasiAlbertGraphGenerator, EdgeType
    out = {}
    for tag, make_gen in (("er20", lambda: RandomErdosRenyiGraphGenerator(20, 0.15, EdgeType.DISCRETE)),
                          ("er40u", lambda: RandomErdosRenyiGraphGenerator(40, [0.15, 0.02], EdgeType.UNIFORM)),
                          ("ba20", lambda: RandomBarabasiAlbertGraphGenerator(20, 4, EdgeType.DISCRETE)),
                          ("ba60u", lambda: RandomBarabasiAlbertGraphGenerator(60, 4, EdgeType.UNIFORM))):
        np.random.seed(123)
        random.seed(123)
        gen = make_gen()
        out[tag] = np.stack([gen.get() for _ in range(3)]).astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "generators.npz"), **out)
    print("wrote generators.npz")

    dqn_case(er40[:3], os.path.join(nets, "network_best_ER_40spin.pth"))


def dqn_case(graphs, net_path):
    """One reference DQN.train_step (dqn.py:403-451) on a hand-built minibatch: loss, gradients, updated weights."""
    import tempfile
    from src.agents.dqn.dqn import DQN
    from src.agents.dqn.utils import TestMetric
    from src.envs.utils import SetGraphGenerator
    n, T, B = 40, 10, 16
    env = ising_env.make("SpinSystem", SetGraphGenerator(list(graphs), ordered=True), T, **env_args_for(n))
    tmp = tempfile.mkdtemp()
    agent = DQN([env], lambda: MPNN(n_obs_in=7, n_layers=3, n_features=64, n_hid_readout=[], tied_weights=False),
                init_network_params=net_path, double_dqn=True, gamma=0.95, update_learning_rate=False,
                initial_learning_rate=1e-4, minibatch_size=B, logging=False, seed=3, adam_epsilon=1e-8,
                test_save_path=os.path.join(tmp, "t"), network_save_path=os.path.join(tmp, "n"),
                test_metric=TestMetric.BEST)
    w0 = {k: v.clone().numpy() for k, v in agent.network.state_dict().items()}
    # perturb the target network so that online != target (as after some training)
    torch.manual_seed(0)
    with torch.no_grad():
        for p_ in agent.target_network.parameters():
            p_.add_(0.01 * torch.randn_like(p_))
    wt = {k: v.clone().numpy() for k, v in agent.target_network.state_dict().items()}
    rng = np.random.RandomState(11)
    trans, gidx = [], []
    ep = 0
    state = torch.as_tensor(env.reset())
    g_of_episode = [i for i in range(len(graphs)) if np.array_equal(graphs[i], env.matrix)][0]
    while len(trans) < B:
        a = int(rng.randint(n)) if rng.rand() < 0.5 else int(np.argmax(env.scorer.get_score_mask(env.state[0], env.matrix)))
        nxt, r, d, _ = env.step(a)
        if env.current_step in (1, 4, 7, 10):       # a spread of steps, including the terminal one
            trans.append((state, torch.as_tensor([a], dtype=torch.long), torch.as_tensor([r], dtype=torch.float),
                          torch.as_tensor(nxt), torch.as_tensor([d], dtype=torch.float)))
            gidx.append(g_of_episode)
        if d:
            state = torch.as_tensor(env.reset())
            g_of_episode = [i for i in range(len(graphs)) if np.array_equal(graphs[i], env.matrix)][0]
        else:
            state = torch.as_tensor(nxt)
    batch = [torch.stack(t) for t in zip(*trans)]
    loss = agent.train_step(batch)
    grads = {k: p_.grad.clone().numpy() for k, p_ in agent.network.named_parameters()}
    w1 = {k: v.clone().numpy() for k, v in agent.network.state_dict().items()}
    out = dict(graphs=np.stack(graphs).astype(np.int8), graph_idx=np.array(gidx, dtype=np.int32),
               rows=batch[0][:, :7, :].float().numpy(), rows_next=batch[3][:, :7, :].float().numpy(),
               actions=batch[1][:, 0].numpy().astype(np.int64), rewards=batch[2][:, 0].numpy(),
               dones=batch[4][:, 0].numpy(), loss=np.float64(loss), gamma=np.float64(0.95), lr=np.float64(1e-4))
    for k in w0:
        out["w::" + k], out["wt::" + k], out["g::" + k], out["w1::" + k] = w0[k], wt[k], grads[k], w1[k]
    np.savez_compressed(os.path.join(HERE, "dqn_er40.npz"), **out)
    print("wrote dqn_er40.npz, loss %.6g, dones %s" % (loss, batch[4][:, 0].tolist()))


# ------------------------------------------------------------------------------------------------------------------
# round 2: multi-graph test_network, DQN.evaluate_agent, epsilon-greedy acting
# ------------------------------------------------------------------------------------------------------------------
def _quiet(fn, *a, **k):
    with open(os.devnull, "w") as dn:
        old = sys.stdout
        sys.stdout = dn
        try:
            return fn(*a, **k)
        finally:
            sys.stdout = old


def multi_graph_case(name, graphs, net, weights, seed, n_attempts):
    """The reference's test_network over SEVERAL same-sized graphs with different maximum degrees: it batches the
    attempts of one graph at a time, so `norm / norm.max()` (mpnn.py:102) is per graph."""
    n = graphs[0].shape[0]
    np.random.seed(seed)
    res, raw, hist = _quiet(test_network, net, env_args_for(n), list(graphs), "cpu", 2, n_attempts=n_attempts,
                            return_raw=True, return_history=True)
    G = len(graphs)
    out = dict(graphs=np.stack(graphs).astype(np.int8), seed=np.int32(seed), n_attempts=np.int32(n_attempts),
               max_degree=np.array([(g != 0).sum(1).max() for g in graphs], dtype=np.int32),
               init_spins=np.stack([np.array(raw["init spins"][j]) for j in range(G)]).astype(np.int8),
               actions=np.array([[[int(a) for a in row[1:]] for row in hist["actions"][j]] for j in range(G)], dtype=np.int32),
               rewards=np.array([[[float(x) for x in row[1:]] for row in hist["rewards"][j]] for j in range(G)], dtype=np.float64),
               scores=np.array([[[float(x) for x in row] for row in hist["scores"][j]] for j in range(G)], dtype=np.float64),
               cuts=np.array([raw["cuts"][j] for j in range(G)], dtype=np.float64),
               sols=np.array([raw["sols"][j] for j in range(G)]).astype(np.int8),
               greedy_cuts=np.array([raw["greedy cuts"][j] for j in range(G)], dtype=np.float64),
               res_cut=np.array(res["cut"], dtype=np.float64), res_mean_cut=np.array(res["mean cut"], dtype=np.float64),
               res_greedy_single=np.array(res["greedy (+1 init) cut"], dtype=np.float64),
               res_greedy_rand=np.array(res["greedy (rand init) cut"], dtype=np.float64),
               res_greedy_rand_mean=np.array(res["greedy (rand init) mean cut"], dtype=np.float64))
    for k, v in weights.items():
        out["w::" + k] = v
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %s: max degrees %s, cuts %s" % (path, out["max_degree"].tolist(), out["res_cut"].tolist()))


def dqn_eval_case(graphs, net_path):
    """Reference DQN.evaluate_agent (dqn.py:514-602) and DQN.act (dqn.py:453-465) with seeded global RNGs."""
    import random
    import tempfile
    import src.agents.dqn.dqn as dqn_mod
    from src.agents.dqn.dqn import DQN
    from src.agents.dqn.utils import TestMetric
    from src.envs.utils import SetGraphGenerator
    n, T = graphs[0].shape[0], 2 * graphs[0].shape[0]
    gen = SetGraphGenerator(list(graphs), ordered=True)
    log = []
    orig_get = gen.get

    def logged_get(*a, **k):
        m = orig_get(*a, **k)
        log.append([i for i in range(len(graphs)) if m is graphs[i]][0])
        return m

    gen.get = logged_get
    env = ising_env.make("SpinSystem", gen, T, **env_args_for(n))
    tmp = tempfile.mkdtemp()
    out = {}
    for metric, tag in ((TestMetric.BEST, "best"), (TestMetric.FINAL, "final")):
        agent = DQN([env], lambda: MPNN(n_obs_in=7, n_layers=3, n_features=64, n_hid_readout=[], tied_weights=False),
                    init_network_params=net_path, minibatch_size=4, test_episodes=6, logging=False, seed=3,
                    test_save_path=os.path.join(tmp, "t"), network_save_path=os.path.join(tmp, "n"), test_metric=metric)
        gen.i = 1
        del log[:]
        random.seed(77)
        np.random.seed(77)
        captured = []
        real_mean = np.mean

        def spy_mean(x, *a, **k):
            captured.append(np.array(x, dtype=np.float64))
            return real_mean(x, *a, **k)

        dqn_mod.np.mean = spy_mean
        try:
            score, sol = agent.evaluate_agent()
        finally:
            dqn_mod.np.mean = real_mean
        out[tag + "_score"], out[tag + "_solution"] = np.float64(score), np.float64(sol)
        out[tag + "_scores"], out[tag + "_solutions"] = captured[0], captured[1]
        out[tag + "_graph_order"] = np.array(log, dtype=np.int32)
        print("evaluate_agent(%s): score %s solution %s graphs %s" % (tag, score, sol, log))

    # epsilon-greedy acting: reference draws `random.uniform(0, 1)` then, when exploring, `np.random.randint(0, n)`
    agent.epsilon = 0.5
    gen.i = 0
    random.seed(5)
    np.random.seed(5)
    obs = env.reset()
    acts, spins0 = [], env.state[0, :n].copy()
    for t in range(24):
        a = agent.act(torch.FloatTensor(np.array(obs)).clone(), True)     # dqn.py:296
        acts.append(int(a))
        obs, _, _, _ = env.step(int(a))
    out.update(graphs=np.stack(graphs).astype(np.int8), act_spins=spins0.astype(np.int8), act_actions=np.array(acts, dtype=np.int32),
               act_epsilon=np.float64(0.5), act_graph=np.int32(0), seed_eval=np.int32(77), seed_act=np.int32(5),
               gen_start=np.int32(1), minibatch_size=np.int32(4), test_episodes=np.int32(6))
    sd = torch.load(net_path, map_location="cpu")
    for k, v in sd.items():
        out["w::" + k] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "dqn_eval_er40.npz"), **out)
    print("wrote dqn_eval_er40.npz, act actions %s" % acts)


def main_round2():
    nets = os.path.join(REF, "experiments/pretrained_agent/networks/eco")
    val = os.path.join(REF, "_graphs/validation")
    er20 = _quiet(load_graph_set, os.path.join(val, "ER_20spin_p15_100graphs.pkl"))
    er40 = _quiet(load_graph_set, os.path.join(val, "ER_40spin_p15_100graphs.pkl"))
    net20, w20 = load_net(os.path.join(nets, "network_best_ER_20spin.pth"))
    deg = [int((g != 0).sum(1).max()) for g in er20]
    pick, seen = [], set()
    for i, d in enumerate(deg):                 # four graphs with four different maximum degrees
        if d not in seen:
            seen.add(d)
            pick.append(i)
        if len(pick) == 4:
            break
    multi_graph_case("multi_er20", [er20[i] for i in pick], net20, w20, seed=21, n_attempts=4)
    net40, w40 = load_net(os.path.join(nets, "network_best_ER_40spin.pth"))
    deg = [int((g != 0).sum(1).max()) for g in er40]
    pick = [int(np.argmin(deg)), int(np.argmax(deg)), 0]
    multi_graph_case("multi_er40", [er40[i] for i in pick], net40, w40, seed=22, n_attempts=3)
    dqn_eval_case([er40[i] for i in pick], os.path.join(nets, "network_best_ER_40spin.pth"))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "s2v":      # only the S2V cases (added later in round 1)
        main_s2v()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mincut":   # only the Min-Cut cases (added later in round 1)
        main_mincut()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "round2":   # multi-graph test_network, evaluate_agent, act (round 2)
        main_round2()
        sys.exit(0)
    main()
