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


@pytest.mark.parametrize("n,p,B,loss_name,norm_mode", [(17, 0.3, 5, "mse", "batch"), (40, 0.15, 64, "mse", "batch"),
                                                       (40, 0.15, 16, "huber", "set"), (100, 0.08, 7, "mse", "graph"),
                                                       (200, 0.04, 9, "huber", "batch"), (333, 0.03, 3, "mse", "batch")])
def test_grad_kernels_match_autograd(n, p, B, loss_name, norm_mode):
    """eco_mpnn_grad (forward + backward in hand-written kernels) against autograd through the PyTorch module of the
    same network on the same minibatch: loss 1e-5 rel, every gradient tensor 1e-4 of its largest entry; and two calls
    give identical bits."""
    import ctypes as C
    import torch.nn.functional as F
    import eco_dqn_b200.engine as eng
    from eco_dqn_b200 import _lib
    from eco_dqn_b200._lib import lib, check
    from eco_dqn_b200.networks.mpnn import MPNN
    rng = np.random.default_rng(7 * n + B)
    torch.manual_seed(n + B)
    G = 3
    Js = np.zeros((G, n, n), dtype=np.int8)
    for g in range(G):
        up = np.triu(rng.random((n, n)) < p, 1)
        a = (up * np.where(rng.random((n, n)) < 0.5, -1, 1)).astype(np.int8)
        Js[g] = a + a.T
    Js[:, 0, 1] = Js[:, 1, 0] = 1                                   # no empty graph
    dev = torch.device("cuda:0")
    gs = eng.GraphSet(Js, device=dev)
    NP = gs.NP
    net = MPNN().to(dev)
    with torch.no_grad():
        for q in net.parameters():
            q.copy_(torch.randn_like(q) * (0.3 if q.dim() > 1 else 0.1))
    gidx = torch.tensor(rng.integers(0, G, size=B), dtype=torch.int32, device=dev)
    rows = torch.tensor(rng.standard_normal((B, 7, n)).astype(np.float32), device=dev)
    rows[:, 0] = torch.sign(rows[:, 0])
    rows[:, 3:] = rows[:, 3:, :1]                                   # global observables: constant over the vertices
    xn = torch.zeros(B, 3, NP, device=dev)
    xn[:, :, :n] = rows[:, :3]
    xg = rows[:, 3:, 0].contiguous()
    actions = torch.tensor(rng.integers(0, n, size=B), dtype=torch.int32, device=dev)
    targets = torch.tensor(rng.standard_normal(B).astype(np.float32) * 3, device=dev)
    deg = torch.tensor((Js != 0).sum(1).clip(1).astype(np.float32), device=dev)
    norm_max = {"batch": float(deg[gidx.long()].max()), "set": 0.0, "graph": -1.0}[norm_mode]

    # reference: autograd through the module (one episode at a time for per-graph normalisation)
    adj = torch.tensor(Js.astype(np.float32), device=dev)[gidx.long()]
    obs = torch.cat([rows, adj], dim=1)
    if norm_mode == "graph":
        q_all = torch.stack([net(obs[b:b + 1]).reshape(-1) for b in range(B)])
    elif norm_mode == "set":
        # the whole set's max degree: append one (unused) episode per graph so that norm.max() covers the set
        extra = torch.cat([torch.zeros(G, 7, n, device=dev), torch.tensor(Js.astype(np.float32), device=dev)], dim=1)
        q_all = net(torch.cat([obs, extra], dim=0))[:B]
    else:
        q_all = net(obs)
    qv = q_all.gather(1, actions.long().unsqueeze(1)).squeeze(1)
    loss_ref = {"mse": F.mse_loss, "huber": F.smooth_l1_loss}[loss_name](qv, targets, reduction="mean")
    net.zero_grad()
    loss_ref.backward()

    w = net.engine_weights(dev)
    nb = lib().eco_mpnn_grad_scratch_bytes(B, n)
    scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
    outs = []
    for rep in range(2):
        flat = torch.full((_lib.MPNN_N_PARAMS,), float("nan"), device=dev)
        loss = torch.full((1,), float("nan"), device=dev)
        check(lib().eco_mpnn_grad(C.byref(gs.c), C.byref(w.c), B, eng._ptr(gidx), eng._ptr(xn), eng._ptr(xg), norm_max,
                                  eng._ptr(actions), eng._ptr(targets), {"mse": _lib.LOSS_MSE, "huber": _lib.LOSS_HUBER}[loss_name],
                                  eng._ptr(loss), eng._ptr(flat), eng._ptr(scratch), eng._stream()))
        torch.cuda.synchronize()
        outs.append((loss.cpu().numpy().copy(), flat.cpu().numpy().copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    loss_k, flat_k = float(outs[0][0][0]), outs[0][1]
    loss_r = float(loss_ref.detach())
    assert abs(loss_k - loss_r) <= 1e-5 * abs(loss_r) + 1e-7, (loss_k, loss_r)
    params = dict(net.named_parameters())
    off = 0
    for key, shp in zip(eng.STATE_DICT_KEYS, eng.STATE_DICT_SHAPES):
        cnt = int(np.prod(shp))
        gk = flat_k[off:off + cnt].reshape(shp)
        gr = params[key].grad.cpu().numpy()
        assert np.isfinite(gk).all(), key
        assert np.abs(gk - gr).max() <= 1e-4 * np.abs(gr).max() + 1e-8, (key, np.abs(gk - gr).max(), np.abs(gr).max())
        off += cnt


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_kernel_adam_matches_torch_adam(weight_decay):
    """eco_mpnn_adam (one launch over the 12 live parameter tensors) against torch.optim.Adam over several steps with the
    same gradients; the engine's packed bf16 copy follows the update."""
    from eco_dqn_b200.networks.mpnn import MPNN
    from eco_dqn_b200.agents.dqn.utils import KernelAdam
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    a = MPNN().to(dev)
    b = MPNN().to(dev)
    b.load_state_dict(a.state_dict())
    opt_a = KernelAdam(a, lr=3e-3, eps=1e-8, weight_decay=weight_decay)
    opt_b = torch.optim.Adam(b.parameters(), lr=3e-3, eps=1e-8, weight_decay=weight_decay)
    for step in range(6):
        if step == 3:
            opt_a.param_groups[0]['lr'] = opt_b.param_groups[0]['lr'] = 1e-3        # the trainer's lr schedule
        for pa, pb in zip(a.parameters(), b.parameters()):
            g = torch.randn_like(pa) * (10.0 ** (step - 3))
            pa.grad, pb.grad = g.clone(), g.clone()
        opt_a.step()
        opt_b.step()
        for (k, pa), pb in zip(a.named_parameters(), b.parameters()):
            assert torch.allclose(pa, pb, rtol=2e-6, atol=1e-8), (step, k, float((pa - pb).abs().max()))
    # the engine sees the updated values: Q from the kernels equals Q from a freshly built weight set
    import eco_dqn_b200.engine as eng
    w_live = a.engine_weights(dev)
    w_new = eng.MPNNWeights(a.state_dict(), device=dev)
    assert w_live.aliases and torch.equal(w_live._packed, w_new._packed)


def _fill_replay(agent, timesteps):
    """Act (randomly: training has not started) until `timesteps` transitions are stored; no update runs."""
    start = agent.replay_start_size
    agent.replay_start_size = 10 ** 9
    agent.learn(timesteps=timesteps)
    agent.replay_start_size = start


def _opt_state(agent):
    o = agent.optimizer
    return ([p.detach().clone() for p in agent.network.parameters()], o.exp_avg.clone(), o.exp_avg_sq.clone(), o.step_dev.clone())


def _restore(agent, st):
    o = agent.optimizer
    with torch.no_grad():
        for p, q in zip(agent.network.parameters(), st[0]):
            p.copy_(q)
        o.exp_avg.copy_(st[1]); o.exp_avg_sq.copy_(st[2]); o.step_dev.copy_(st[3])
    agent.network.engine_weights(agent.device).repack()


def test_captured_update_equals_eager_update(tmp_path):
    """The update replayed as one CUDA graph (TD target, eco_mpnn_grad, Adam with its step counter and learning rate on the
    device, operand re-pack) gives bit for bit the parameters of the same updates launched eagerly, including across a
    learning-rate change and a target-network sync."""
    gs = np.load(os.path.join(GOLDEN, "graphsets.npz"))
    agent = make_agent(tmp_path, list(gs["er20"][:6]), 40, n_envs=8, init_weight_std=0.05, minibatch_size=16,
                       replay_buffer_size=1000, replay_start_size=64)
    _fill_replay(agent, 320)
    gen = torch.Generator(device="cuda").manual_seed(3)
    idxs = [agent.replay_buffer.sample_indices(16, gen) for _ in range(6)]
    st0 = _opt_state(agent)
    tgt0 = {k: v.clone() for k, v in agent.target_network.state_dict().items()}

    def run(step_fn):
        _restore(agent, st0)
        agent.target_network.load_state_dict(tgt0)
        agent.target_network.engine_weights(agent.device)
        out = []
        for k, idx in enumerate(idxs):
            if k == 2:
                agent.optimizer.param_groups[0]['lr'] = 3e-4
            if k == 4:
                agent.target_network.load_state_dict(agent.network.state_dict())
                agent.target_network.engine_weights(agent.device)
            out.append(float(step_fn(idx)))
        agent.optimizer.param_groups[0]['lr'] = 1e-4
        return out, [p.detach().clone() for p in agent.network.parameters()], int(agent.optimizer.steps)

    l_eager, p_eager, n_eager = run(lambda idx: agent._update(agent.replay_buffer.gather(idx)))
    assert agent._cg is None
    l_graph, p_graph, n_graph = run(agent._train_step_device)
    assert agent._cg is not None and n_eager == n_graph == int(st0[3].item()) + 6
    assert l_eager == l_graph
    for a, b in zip(p_eager, p_graph):
        assert torch.equal(a, b)
    assert not torch.equal(p_eager[2], st0[0][2])


def test_peer_adam_with_one_rank_equals_device_adam():
    """eco_dp_adam with world = 1 (the exchange degenerates to this rank's own slot) equals eco_mpnn_adam_dev."""
    from eco_dqn_b200.networks.mpnn import MPNN
    from eco_dqn_b200.agents.dqn.utils import KernelAdam
    torch.manual_seed(1)
    nets = [MPNN().cuda(), MPNN().cuda()]
    nets[1].load_state_dict(nets[0].state_dict())
    opts = [KernelAdam(n, lr=1e-3, weight_decay=0.01) for n in nets]
    opts[1].attach_peers(1, 0, lambda mine: mine.reshape(1, 64))
    gen = torch.Generator(device="cuda").manual_seed(5)
    for step in range(4):
        grads = [torch.randn(p.shape, device="cuda", generator=gen) * 10.0 ** (step - 2) for p in nets[0].parameters()]
        for net, opt in zip(nets, opts):
            for p, g in zip(net.parameters(), grads):
                p.grad = g.clone()
            opt.step()
        for a, b in zip(nets[0].parameters(), nets[1].parameters()):
            assert torch.equal(a, b)
    assert opts[1].steps == 4 and int(opts[1].err_dev.item()) == 0
    opts[1].close()


def test_captured_acting_step_fills_the_replay_consistently(tmp_path):
    """n_envs > 1: `learn` replays the acting lock-step (observation snapshot, epsilon-greedy action, env step, replay append)
    as one CUDA graph.  Before training starts every transition must still be what the environments produced: within an
    episode the next observation of step t is the observation of step t + 1 of the same environment, the graph index stays,
    `done` is raised on the last step only, actions are valid vertices, and the per-episode reward sums are the scores; the
    host mirrors (replay size / position, env step counter) follow."""
    gs = np.load(os.path.join(GOLDEN, "graphsets.npz"))
    graphs = list(gs["er20"][:6])
    E, T = 4, 40
    agent = make_agent(tmp_path, graphs, T, n_envs=E, update_frequency=8, update_target_frequency=160,
                       replay_start_size=10 ** 6, replay_buffer_size=1000, test_frequency=10 ** 9, test_episodes=6,
                       save_network_frequency=10 ** 9, init_weight_std=0.01, final_exploration_step=600,
                       final_exploration_rate=0.05)
    assert agent._act_graph_ok()
    agent.learn(timesteps=E * T * 2 + E * 7)               # two full episodes and seven steps of a third
    rb = agent.replay_buffer
    n_rows = E * T * 2 + E * 7
    assert len(rb) == n_rows and rb._position == n_rows and int(agent._ag_pos.item()) == n_rows
    assert agent._env.current_step == 7
    xn, xg, xn2, xg2 = rb.xn[:n_rows].cpu(), rb.xg[:n_rows].cpu(), rb.xn_next[:n_rows].cpu(), rb.xg_next[:n_rows].cpu()
    act, done, graph, rew = rb.action[:n_rows].cpu(), rb.done[:n_rows].cpu(), rb.graph[:n_rows].cpu(), rb.reward[:n_rows].cpu()
    assert int(act.min()) >= 0 and int(act.max()) < 20
    for e in range(E):
        rows = torch.arange(e, n_rows, E)                  # environment e, step after step
        for ep in range(2):
            r = rows[ep * T:(ep + 1) * T]
            assert torch.equal(xn2[r[:-1]], xn[r[1:]]) and torch.equal(xg2[r[:-1]], xg[r[1:]])
            assert len(graph[r].unique()) == 1
            assert done[r[:-1]].sum() == 0 and done[r[-1]] == 1
            assert not torch.equal(xn[r[0]], xn[r[-1]])
        assert done[rows[2 * T:]].sum() == 0
    # the running scores of the current (third) episode are its reward sums
    cur = torch.stack([rew[torch.arange(e, n_rows, E)[2 * T:]].double().sum() for e in range(E)])
    assert torch.allclose(agent._scores.cpu(), cur, atol=1e-6)
