"""Oracle: batched greedy-Q rollout on B copies of one graph plus the Greedy baseline.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates reference
experiments/utils.py::__test_network_batched (hot loop :169-207, predict :57-66, greedy baseline :218-227)
with the oracle env and oracle MPNN.  Under Stopping.NORMAL every episode finishes at step T = max_steps,
so the reference's `zip(test_envs, actions)` bookkeeping (quirk A.4-6) reduces to a plain loop.
"""
import time

import numpy as np
import torch

from .spin_env import MaxCutEnv
from .mpnn import as_torch_weights, mpnn_forward


def rollout(J, weights, init_spins, max_steps, basin_reward=None, forced_actions=None, record_obs=False,
            q_hook=None, min_cut=False):
    """Greedy-Q (or teacher-forced) rollout of len(init_spins) episodes on graph J.

    Returns dict(actions[B,T], rewards[B,T] f64, scores[B,T+1] f64, best_cut[B], best_spins[B,N],
    final_spins[B,N], env_steps, seconds) -- `seconds` spans only the step loop, like the reference's
    t_total (experiments/utils.py:164,214)."""
    w = as_torch_weights(weights) if weights is not None else None
    B, T = len(init_spins), int(max_steps)
    envs = [MaxCutEnv(J, T, basin_reward, min_cut=min_cut) for _ in range(B)]
    obs = [e.reset(s) for e, s in zip(envs, init_spins)]
    n = envs[0].n
    actions = np.zeros((B, T), dtype=np.int32)
    rewards = np.zeros((B, T), dtype=np.float64)
    scores = np.zeros((B, T + 1), dtype=np.float64)
    scores[:, 0] = [e.score for e in envs]
    obs_rec = np.zeros((B, T + 1, 7, n), dtype=np.float32) if record_obs else None
    t0 = time.perf_counter()
    for t in range(T):
        if forced_actions is None or record_obs or q_hook is not None:
            ob = torch.FloatTensor(np.array(obs))          # experiments/utils.py:174
            if record_obs:
                obs_rec[:, t] = ob[:, :7, :].numpy()
        if forced_actions is None or q_hook is not None:
            qs = mpnn_forward(w, ob)
            if q_hook is not None:
                q_hook(t, qs)
        if forced_actions is None:
            acts = qs.argmax(1, True).squeeze(1).numpy()   # experiments/utils.py:65 (first max on ties)
        else:
            acts = forced_actions[:, t]
        obs = []
        for i, (e, a) in enumerate(zip(envs, acts)):
            o, r, _, _ = e.step(int(a))
            actions[i, t], rewards[i, t], scores[i, t + 1] = a, r, e.score
            obs.append(o)
    seconds = time.perf_counter() - t0
    if record_obs:
        obs_rec[:, T] = torch.FloatTensor(np.array(obs))[:, :7, :].numpy()
    return dict(actions=actions, rewards=rewards, scores=scores,
                best_cut=np.array([e.best_solution for e in envs], dtype=np.float64),
                best_spins=np.stack([e.best_spins for e in envs]).astype(np.int8),
                final_spins=np.stack([e.spins for e in envs]).astype(np.int8),
                obs=obs_rec, env_steps=B * T, seconds=seconds)


def rollout_s2v(J, weights, max_steps, forced_actions=None, q_hook=None):
    """One S2V-DQN episode (experiments/pretrained_agent/test_s2v.py with SpinBasis.SIGNED): spins start at -1, are
    flipped at most once, observation = [spin row; adjacency], dense reward, argmax over the spins still at -1
    (experiments/utils.py:67-74: the others are filled with -1000), done when none is left or after max_steps."""
    w = as_torch_weights(weights)
    env = MaxCutEnv(J, max_steps, None, reversible=False, dense_reward=True)
    env.reset()
    actions, rewards, scores, dones = [], [], [env.score], []
    done = False
    t = 0
    while not done:
        ob = torch.FloatTensor(np.vstack((env.state[0:1], env.J)))[None]
        if forced_actions is None or q_hook is not None:
            qs = mpnn_forward(w, ob)
            if q_hook is not None:
                q_hook(t, qs)
        if forced_actions is None:
            mask = torch.from_numpy(env.state[0] != -1)[None]
            a = int(qs.masked_fill(mask, -1000).argmax(1, True).squeeze(1).numpy()[0])
        else:
            a = int(forced_actions[t])
        _, r, done, _ = env.step(a)
        actions.append(a); rewards.append(r); scores.append(env.score); dones.append(done)
        t += 1
    return dict(actions=np.array(actions, dtype=np.int32), rewards=np.array(rewards, dtype=np.float64),
                scores=np.array(scores, dtype=np.float64), dones=np.array(dones, dtype=np.uint8),
                best_cut=float(env.best_solution), best_spins=env.best_spins.astype(np.int8))


def greedy_baseline(J, init_spins, max_steps, basin_reward=None, min_cut=False):
    """reference experiments/utils.py:218-227 + src/agents/solver.py:105-131."""
    cuts, spins, steps = [], [], []
    for s in init_spins:
        e = MaxCutEnv(J, max_steps, basin_reward, min_cut=min_cut)
        e.reset(s)
        steps.append(e.greedy_solve())
        cuts.append(e.best_solution)
        spins.append(e.best_spins.astype(np.int8))
    return np.array(cuts, dtype=np.float64), np.stack(spins), np.array(steps, dtype=np.int32)
