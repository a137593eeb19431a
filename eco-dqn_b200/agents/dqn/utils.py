"""Replay storage and small helpers of the DQN trainer (reference src/agents/dqn/utils.py:14-83).

The reference stores two full observations per transition -- (7 + N) x N fp64 each, adjacency included -- in a
Python dict and stacks a minibatch on a prefetch thread.  Here a transition is the engine's compact feature
format: three per-vertex rows + four global scalars for s and s' (fp32), the graph's slot in the device graph
ring, action, reward, done.  Everything lives in preallocated device tensors; sampling is an index gather.
"""
import random
from enum import Enum

import numpy as np
import torch


class TestMetric(Enum):          # reference dqn/utils.py:14-19
    FINAL = 1
    BEST = 2
    CUMULATIVE_REWARD = 3
    ENERGY_ERROR = 4


def set_global_seed(seed, env=None):   # reference dqn/utils.py:22-26
    torch.manual_seed(seed)
    if env is not None:
        env.set_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


class ReplayBuffer:
    FIELDS = ("xn", "xg", "action", "reward", "xn_next", "xg_next", "done", "graph")

    def __init__(self, capacity, n_padded, device):
        self._capacity = int(capacity)
        self._size = 0
        self._position = 0
        self.device = device
        c, npad = self._capacity, n_padded
        self.xn = torch.zeros(c, 3, npad, dtype=torch.float32, device=device)
        self.xg = torch.zeros(c, 4, dtype=torch.float32, device=device)
        self.xn_next = torch.zeros(c, 3, npad, dtype=torch.float32, device=device)
        self.xg_next = torch.zeros(c, 4, dtype=torch.float32, device=device)
        self.action = torch.zeros(c, dtype=torch.int64, device=device)
        self.reward = torch.zeros(c, dtype=torch.float32, device=device)
        self.done = torch.zeros(c, dtype=torch.float32, device=device)
        self.graph = torch.zeros(c, dtype=torch.int32, device=device)

    def add(self, xn, xg, action, reward, xn_next, xg_next, done, graph):
        """Append E transitions (first dimension of every argument); the oldest are overwritten when full."""
        e = xn.shape[0]
        idx = (self._position + torch.arange(e, device=self.device)) % self._capacity
        self.xn[idx], self.xg[idx] = xn, xg
        self.xn_next[idx], self.xg_next[idx] = xn_next, xg_next
        self.action[idx] = action.to(torch.int64)
        self.reward[idx] = reward.to(torch.float32)          # reference stores rewards as fp32 (dqn.py:299)
        self.done[idx] = done.to(torch.float32)
        self.graph[idx] = graph.to(torch.int32)
        self._position = (self._position + e) % self._capacity
        self._size = min(self._size + e, self._capacity)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self._size, (batch_size,), device=self.device, generator=generator)
        return {k: getattr(self, k)[idx] for k in self.FIELDS}

    def __len__(self):
        return self._size
