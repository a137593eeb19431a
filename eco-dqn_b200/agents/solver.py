"""Single-environment solvers with the reference's interface (reference src/agents/solver.py:11-267): `Greedy` and
`Network`, the two entry points on the rollout path.  They drive a SpinSystem facade one step at a time (each step is a
device launch); batched evaluation lives in experiments.utils.test_network.  The other reference solvers (Random,
CoverMatching, Cplex, NetworkX) belong to problems outside this path."""
from abc import ABC, abstractmethod

import numpy as np
import torch


class SpinSolver(ABC):
    def __init__(self, env, record_cut=False, record_rewards=False, record_qs=False, verbose=False, name=None):
        self.env = env
        self.verbose = verbose
        self.record_solution = record_cut
        self.record_rewards = record_rewards
        self.record_qs = record_qs
        self.name = name
        self.measure = 0
        self.total_reward = 0

    def reset(self, spins=None):
        self.total_reward = 0
        self.env.reset(spins)

    def set_env(self, env):
        self.env = env

    def solve(self, *args):
        done = False
        while not done:
            reward, done = self.step(*args)
            self.total_reward += reward
        self.measure = self.env.scorer.get_solution(self.env.state[0, :self.env.n_spins], self.env.matrix)
        return self.total_reward

    @abstractmethod
    def step(self, *args):
        raise NotImplementedError()


class Greedy(SpinSolver):
    """Flip the vertex with the largest immediate gain; stop when the best gain is negative (solver.py:105-131)."""

    def step(self):
        rewards_available = self.env.scorer.get_score_mask(self.env.state[0, :self.env.n_spins], self.env.matrix)
        if self.env.reversible_spins:
            action = rewards_available.argmax()
        else:                                   # solver.py:116-121: only the spins that have not been flipped yet
            masked = rewards_available.copy()
            np.putmask(masked, self.env.get_observation()[0, :] != self.env.get_allowed_action_states(),
                       np.finfo(np.float64).min)
            action = masked.argmax()
        if rewards_available[action] < 0:
            return 0, True
        _, reward, done, _ = self.env.step(action)
        return reward, done

    def solve(self, *args):
        # fast path: the whole greedy descent in one device call per step, no host decisions
        env = self.env._env
        remaining = self.env.max_steps - env.current_step
        before = env.episodes()["total_reward"][0]
        env.rollout(policy="greedy", n_steps=remaining)
        self.env._ep = None
        self.total_reward += float(env.episodes()["total_reward"][0] - before)
        self.measure = self.env.scorer.get_solution(self.env.state[0, :self.env.n_spins], self.env.matrix)
        return self.total_reward


class Network(SpinSolver):
    """Greedy w.r.t. the Q-network, with the reference's per-step history records (solver.py:161-267)."""

    epsilon = 0

    def __init__(self, network, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.device = torch.device("cuda")
        self.network = network.to(self.device)
        self.network.eval()
        self.record_solution = self.record_qs = self.record_rewards = self.record_spins = True
        self.history = []

    def _record(self, action, reward, qs):
        env = self.env
        spins = env.state[0, :env.n_spins]
        rec = [float(action), float(env.scorer.get_solution(spins, env.matrix)), float(reward), qs, list(spins),
               list(env.scorer.get_score_mask(spins, env.matrix)), env.scorer.is_valid(spins, env.matrix)]
        self.history.append(rec)

    def reset(self, spins=None, clear_history=True):
        self.env.reset(spins)
        self.total_reward = 0
        if clear_history:
            self.history = []
            self._record(0, 0, [0] * self.env.n_spins)

    @torch.no_grad()
    def step(self):
        env = self.env
        qs, act = env._env.q_values(self.network.engine_weights(self.device))
        qs = qs[0]
        if np.random.uniform(0, 1) >= self.epsilon:
            action = int(act[0])                 # (irreversible spins: q_values already masked the flipped ones)
        elif env.reversible_spins:
            action = np.random.randint(0, env.action_space.n)
        else:                                    # solver.py:234-240: a random spin among those still at -1
            x = (env.state[0, :env.n_spins] == env.get_allowed_action_states()).nonzero()[0]
            action = int(x[np.random.randint(0, len(x))])
        _, reward, done, _ = env.step(action)
        self._record(action, reward, qs.tolist())
        return reward, done
