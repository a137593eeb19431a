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
        """`batch_size` distinct transitions, like the reference's `random.sample(memory, batch_size)`
        (dqn/utils.py:45-48: without replacement)."""
        if batch_size > self._size:
            raise ValueError("Sample larger than population or is negative")     # what random.sample raises
        idx = torch.randperm(self._size, device=self.device, generator=generator)[:batch_size]
        return {k: getattr(self, k)[idx] for k in self.FIELDS}

    def sample_indices(self, batch_size, generator=None):
        """Row numbers of `batch_size` distinct stored transitions (device int64)."""
        if batch_size > self._size:
            raise ValueError("Sample larger than population or is negative")
        return torch.randperm(self._size, device=self.device, generator=generator)[:batch_size]

    def gather(self, idx):
        return {k: getattr(self, k)[idx] for k in self.FIELDS}

    def __len__(self):
        return self._size


class KernelAdam:
    """torch.optim.Adam for the MPNN's 12 tensors in one kernel launch (csrc/dp_update.cu): the update of reference
    dqn.py:449.  Keeps the small part of the torch optimizer interface the trainer uses (`param_groups[i]['lr']`,
    `zero_grad`, `step`).  The moments live in two flat fp32 buffers in state_dict order; the step counter and the learning
    rate live on the DEVICE (eco_mpnn_adam_dev), so a whole update can be captured in a CUDA graph -- call `sync_lr()`
    outside the graph after changing `param_groups[0]['lr']`.

    `attach_peers(world, rank, all_gather_bytes)`: data-parallel training on the GPUs of one box -- `step()` then runs
    eco_dp_adam, which sums the ranks' gradients straight out of peer memory (CUDA IPC over NVLink) and applies Adam in the
    same launch; no separate all-reduce."""

    def __init__(self, network, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        from ... import engine, _lib
        self._engine, self._lib = engine, _lib
        self.network = network
        self.params = list(network.parameters())
        self.param_groups = [{'lr': float(lr), 'betas': tuple(betas), 'eps': float(eps),
                              'weight_decay': float(weight_decay), 'params': self.params}]
        dev = self.params[0].device
        self.exp_avg = torch.zeros(_lib.MPNN_N_PARAMS, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(_lib.MPNN_N_PARAMS, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.err_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lr_on_device = float(lr)
        self.grad_scale = 1.0          # 1 / world after a SUM all-reduce of the gradients (folded into the kernel)
        self._dp = None

    @property
    def steps(self):
        return int(self.step_dev.item())

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def sync_lr(self):
        """Bring the device copy of the learning rate up to date (a fill kernel: keep it outside captured graphs)."""
        lr = float(self.param_groups[0]['lr'])
        if lr != self._lr_on_device:
            self.lr_dev.fill_(lr)
            self._lr_on_device = lr

    def attach_peers(self, world, rank, all_gather_bytes):
        """all_gather_bytes(uint8 tensor [64]) -> uint8 tensor [world, 64] with every rank's handle, in rank order."""
        import ctypes as C
        L = self._lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.exp_avg.device):
            self._lib.check(L.eco_dp_create(C.byref(h), int(world), int(rank)))
            mine = (C.c_ubyte * 64)()
            self._lib.check(L.eco_dp_handle(h, mine))
            handles = all_gather_bytes(torch.tensor(list(mine), dtype=torch.uint8)).cpu().contiguous()
            assert tuple(handles.shape) == (world, 64)
            buf = (C.c_ubyte * (64 * world)).from_buffer_copy(handles.numpy().tobytes())
            self._lib.check(L.eco_dp_open(h, buf))
        self._dp = h

    def close(self):
        if self._dp is not None:
            self._lib.lib().eco_dp_destroy(self._dp)
            self._dp = None

    def _flat_grad(self):
        """The gradients as one flat buffer in state_dict order (no copy when they are views of eco_mpnn_grad's output)."""
        named = dict(self.network.named_parameters())
        grads = [named[k].grad for k in self._engine.STATE_DICT_KEYS]
        base, off, flat_ok = grads[0], 0, True
        for g in grads:
            flat_ok = flat_ok and g.is_contiguous() and g.untyped_storage().data_ptr() == base.untyped_storage().data_ptr() \
                and g.storage_offset() == base.storage_offset() + off
            off += g.numel()
        if flat_ok:
            return torch.as_strided(base, (off,), (1,), base.storage_offset())
        return torch.cat([g.reshape(-1) for g in grads])

    @torch.no_grad()
    def step(self):
        import ctypes as C
        eng, L = self._engine, self._lib.lib()
        w = self.network.engine_weights(self.params[0].device)
        if not w.aliases:
            raise NotImplementedError("KernelAdam updates the live fp32 parameters in place; this network's engine weights "
                                      "are copies (n_obs_in = 1)")
        g = self.param_groups[0]
        flat = self._flat_grad()
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        with torch.cuda.device(flat.device):
            if self._dp is not None:
                self._lib.check(L.eco_dp_adam(self._dp, C.byref(w.c), eng._ptr(flat), eng._ptr(self.exp_avg),
                                              eng._ptr(self.exp_avg_sq), eng._ptr(self.step_dev), eng._ptr(self.lr_dev),
                                              g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'],
                                              eng._ptr(self.err_dev), eng._stream()))
            else:
                self._lib.check(L.eco_mpnn_adam_dev(C.byref(w.c), eng._ptr(flat), eng._ptr(self.exp_avg),
                                                    eng._ptr(self.exp_avg_sq), eng._ptr(self.step_dev), eng._ptr(self.lr_dev),
                                                    g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'],
                                                    float(self.grad_scale), eng._stream()))
        w.repack()                # the parameters changed behind autograd's version counters
