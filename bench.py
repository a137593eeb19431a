#!/usr/bin/env python
"""bench.py -- env-steps/sec incl. MPNN Q-eval for the batched ECO-DQN Max-Cut rollout (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # reference arm: the CPU oracle port on host cores

One bench "step" = one full greedy-Q rollout of a batch: reset + T = 2N env steps for B episodes, each env step
being one MPNN Q-evaluation + argmax + one fused env-step kernel.  Workload = BASELINE.json configs[1]:
BA-200 (m=4, +-1 weights), B = 4096 concurrent episodes per GPU, G = 4096 distinct graphs per GPU (J int8 is
164 MB > the 126 MB L2, so no L2 flush is needed between iterations), pretrained BA-200 weights.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec incl. MPNN Q-eval"
N_SPINS = 200
BA_M = 4
B_PER_GPU = 4096


def flops_mpnn(n):
    """algorithmic flops of one Q-evaluation of one episode (+-1 graphs)."""
    return 636 * n * n + 108530 * n + 8192


def bytes_env(n):
    """algorithmic bytes of one env step of one episode (features materialised)."""
    return 20.25 * n + 96


def ba_graphs(count, n, m, seed):
    """BA(n, m) graphs with +-1 weights, like RandomBarabasiAlbertGraphGenerator (reference src/envs/utils.py:204-236)."""
    import networkx as nx
    np.random.seed(seed)
    random.seed(seed)
    out = np.zeros((count, n, n), dtype=np.int8)
    for i in range(count):
        g = nx.barabasi_albert_graph(n, m)
        adj = nx.to_numpy_array(g)
        mask = 2. * np.random.randint(2, size=(n, n)) - 1.
        mask = np.tril(mask) + np.triu(mask.T, 1)
        a = adj * mask
        np.fill_diagonal(a, 0)
        out[i] = a.astype(np.int8)
    return out


def er_graphs(count, n, p, seed):
    """G(n, p) graphs with +-1 weights, like RandomErdosRenyiGraphGenerator with EdgeType.DISCRETE (reference
    src/envs/utils.py:165-202: a symmetric Bernoulli(p) mask times a symmetric +-1 matrix, zero diagonal)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((count, n, n), dtype=np.int8)
    iu = np.triu_indices(n, 1)
    for i in range(count):
        keep = rng.random(len(iu[0])) < p
        sign = (2 * rng.integers(0, 2, size=len(iu[0])) - 1).astype(np.int8)
        out[i, iu[0], iu[1]] = keep * sign
        out[i] += out[i].T
    return out


def load_weights():
    """The reference's pretrained eco/network_best_BA_200spin checkpoint, as recorded in the golden fixture (`w::<key>`)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ba200_g0.npz"))
    return {k[3:]: np.asarray(z[k], dtype=np.float32) for k in z.files if k.startswith("w::")}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_sample(J, weights, n_eps, n_steps, seed=0):
    """The reference's CPU path (oracle port of __test_network_batched) on a bounded sample of the workload:
    n_eps episodes of ONE graph (the reference batches copies of one graph), n_steps steady-state env steps,
    torch CPU threads = all host cores.  Returns env-steps/s of the step loop (the span the reference times)."""
    import torch
    from oracle.rollout import rollout as cpu_rollout
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(seed)
    n = J.shape[0]
    spins = (2 * rng.integers(0, 2, size=(n_eps, n)) - 1).astype(np.int8)
    cpu_rollout(J.astype(np.float64), weights, spins[:2], 2, 1.0 / n)            # warm-up (thread pools, allocs)
    out = cpu_rollout(J.astype(np.float64), weights, spins, n_steps, 1.0 / n)
    # max_steps = n_steps here only bounds the loop; every step before the last is a normal steady-state step
    return out["env_steps"] / out["seconds"], cores, out["seconds"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = N_SPINS
    weights = load_weights()
    J = ba_graphs(1, n, BA_M, seed=0)[0]
    n_eps, n_steps = 32, 100
    vals = []
    for i in range(args.warmup):
        cpu_reference_sample(J, weights, 4, 2, seed=i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        v, cores, _ = cpu_reference_sample(J, weights, n_eps, n_steps, seed=100 + i)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = "%d episodes x %d env steps of one BA-%d graph per bench step (full step would be %d x %d)" % (
        n_eps, n_steps, n, B_PER_GPU, 2 * n)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(1),
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def dqn_update_sample(timesteps=3000, n_envs=16):
    """C5 of BASELINE.json: DQN.learn on ER-40 with the hyper-parameters of experiments/train_eco.py (minibatch 64,
    update every 32 steps, replay 5000).  Returns ms per train_step (TD target + eco_mpnn_grad + Adam) and ms per 1000
    environment timesteps of the whole learn loop.  A reported side number, not the bench metric."""
    import contextlib
    import tempfile
    import torch
    import eco_dqn_b200.envs.core as ising_env
    from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,
                                         Stopping, RandomErdosRenyiGraphGenerator, EdgeType)
    from eco_dqn_b200.networks.mpnn import MPNN
    from eco_dqn_b200.agents.dqn.dqn import DQN
    from eco_dqn_b200.agents.dqn.utils import TestMetric
    n = 40
    env_args = {'observables': DEFAULT_OBSERVABLES, 'reward_signal': RewardSignal.BLS, 'extra_action': ExtraAction.NONE,
                'optimisation_target': OptimisationTarget.CUT, 'spin_basis': SpinBasis.SIGNED, 'norm_rewards': True,
                'memory_length': None, 'horizon_length': None, 'stag_punishment': None, 'basin_reward': 1. / n,
                'reversible_spins': True, 'stopping': Stopping.NORMAL}
    with contextlib.redirect_stdout(sys.stderr):
        env = ising_env.make("SpinSystem", RandomErdosRenyiGraphGenerator(n, 0.15, EdgeType.DISCRETE), 2 * n, **env_args)
        tmp = tempfile.mkdtemp()
        agent = DQN([env], lambda: MPNN(), init_weight_std=0.01, double_dqn=True, gamma=0.95, update_learning_rate=False,
                    initial_learning_rate=1e-4, minibatch_size=64, update_frequency=32, update_target_frequency=1000,
                    replay_start_size=500, replay_buffer_size=5000, final_exploration_step=3000,
                    final_exploration_rate=0.05, test_frequency=10 ** 9, save_network_frequency=10 ** 9, logging=False,
                    seed=5, test_metric=TestMetric.BEST, test_save_path=os.path.join(tmp, "s"),
                    network_save_path=os.path.join(tmp, "n"), n_envs=n_envs)
        acc = {"s": 0.0, "n": 0}
        orig = agent.train_step

        def timed(tr):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = orig(tr)
            torch.cuda.synchronize()
            acc["s"] += time.perf_counter() - t0
            acc["n"] += 1
            return out

        agent.train_step = timed
        agent.learn(timesteps=1000)          # warm-up: fills the replay, first updates
        acc["s"], acc["n"] = 0.0, 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        agent.learn(timesteps=timesteps)
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
    return {"workload": "DQN.learn, ER-40 p=0.15, minibatch 64, update every 32 timesteps, %d lock-step environments" % n_envs,
            "train_step_ms": acc["s"] / max(acc["n"], 1) * 1e3, "train_steps": acc["n"],
            "ms_per_1000_timesteps": tot / timesteps * 1e6, "timing": "host clock around synchronised calls"}


def config_dict(world):
    return {"workload": "BA_200spin (m=4, +-1 weights) batched ECO-DQN greedy-Q rollout, %d concurrent episodes per GPU, "
                        "%d distinct graphs per GPU, T=2N=%d env steps per episode" % (B_PER_GPU, B_PER_GPU, 2 * N_SPINS),
            "step": "one full rollout: reset + T x (MPNN Q-eval + argmax + env step) for the whole batch",
            "n_spins": N_SPINS, "episodes_per_gpu": B_PER_GPU, "env_steps_per_episode": 2 * N_SPINS,
            "graphs_per_gpu": B_PER_GPU, "weights": "pretrained eco/network_best_BA_200spin (reference checkpoint)",
            "l2": "inputs larger than L2 (per GPU: bf16 adjacency operand images 708 MB + int8 adjacency 177 MB + 50 MB state); no flush needed",
            "parallelism": "episodes sharded over %d GPU(s), no collective during rollout, all_gather of best cuts" % world}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    import eco_dqn_b200.engine as engine
    from eco_dqn_b200 import _lib
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    n, T, B, G = N_SPINS, 2 * N_SPINS, B_PER_GPU, B_PER_GPU
    impl = {"auto": _lib.MPNN_AUTO, "simt": _lib.MPNN_SIMT, "tc": _lib.MPNN_TCGEN05}[args.mpnn]

    wd = load_weights()
    J = ba_graphs(G, n, BA_M, seed=rank)                      # each rank owns its own graphs + episodes
    rng = np.random.default_rng(1000 + rank)
    gs = engine.GraphSet(J)
    env = engine.BatchedSpinSystem(gs, B, T, 1.0 / n, mpnn_impl=impl)
    w = engine.MPNNWeights(wd)
    used_impl = "tcgen05" if (impl != _lib.MPNN_SIMT and w.c.packed) else "simt"
    spins_dev = torch.from_numpy((2 * rng.integers(0, 2, size=(B, n)) - 1).astype(np.int8)).cuda()
    gidx_dev = torch.arange(B, dtype=torch.int32, device="cuda")
    gathered = [torch.empty(B, dtype=torch.int32, device="cuda") for _ in range(world)] if world > 1 else None

    def one_step():
        env.reset(spins=spins_dev, graph_idx=gidx_dev)
        env.rollout(w)
        bc, _, _ = env.results()
        if world > 1:
            dist.all_gather(gathered, bc)                      # the path's only collective: final best cuts
        return bc

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    barrier()

    # ---- timed region: device timing with CUDA events on the launch stream, max over ranks --------------
    L.eco_launch_count(1)
    L.eco_profile_enable(PROFILE_STRIDE)      # every PROFILE_STRIDE-th launch of each kernel is bracketed by events
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        bc = one_step()
    ev1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = int(L.eco_launch_count(0))
    tot = C.c_double()
    cnt = C.c_int64()
    L.eco_profile_read(0, C.byref(tot), C.byref(cnt))
    mpnn_ms, mpnn_n = tot.value, cnt.value
    L.eco_profile_read(1, C.byref(tot), C.byref(cnt))
    env_ms, env_n = tot.value, cnt.value
    L.eco_profile_enable(0)
    env_steps_total = world * B * T * args.steps
    value = env_steps_total / (ms_total / 1000.0)
    best_mean = float(bc.float().mean().item())

    # ---- e2e: host buffers in, host buffers out, through the C-ABI session (H2D + D2H inside the timing) ----
    sess = engine.HostSession(G, n, B, T, 1.0 / n, wd, impl=impl)
    J_pin = torch.from_numpy(J).pin_memory()
    spins_pin = spins_dev.cpu().pin_memory()
    gidx_pin = torch.arange(B, dtype=torch.int32).pin_memory()
    cut_pin = torch.zeros(B, dtype=torch.int32).pin_memory()
    bs_pin = torch.zeros(B, n, dtype=torch.int8).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    sess.rollout(J_pin, gidx_pin, spins_pin, cut_pin, bs_pin)          # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sess.rollout(J_pin, gidx_pin, spins_pin, cut_pin, bs_pin)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * T * e2e_steps / float(e2e_s.item())
    assert np.array_equal(cut_pin.numpy(), bc.cpu().numpy()), "session result differs from engine result"
    sess.close()

    # ---- env-step kernel alone at a size that is HBM- rather than launch-bound ---------------------------
    env_only = None
    if rank == 0 and not args.skip_env_only and world == 1:
        Bbig = 262144
        envb = engine.BatchedSpinSystem(gs, Bbig, T, 1.0 / n)
        envb.reset(spins=torch.from_numpy((2 * rng.integers(0, 2, size=(Bbig, n)) - 1).astype(np.int8)).cuda())
        gen = torch.Generator(device="cuda").manual_seed(7)
        acts = [torch.randint(0, n, (Bbig,), generator=gen, device="cuda", dtype=torch.int32) for _ in range(24)]
        for a in acts[:4]:
            envb.step(a)
        torch.cuda.synchronize()
        L.eco_profile_enable(1)
        for a in acts[4:]:
            envb.step(a)
        L.eco_profile_read(1, C.byref(tot), C.byref(cnt))
        L.eco_profile_enable(0)
        avg = tot.value / cnt.value / 1000.0
        gbs = bytes_env(n) * Bbig / avg / 1e9
        env_only = {"episodes": Bbig, "avg_launch_us": avg * 1e6, "env_steps_per_s": Bbig / avg,
                    "achieved_gbs": gbs, "peak_gbs": peaks()["hbm_gbs"], "frac": gbs / peaks()["hbm_gbs"],
                    "bytes_per_env_step": bytes_env(n)}
        del envb

    if rank == 0:
        pk = peaks()
        avg_mpnn_s = mpnn_ms / max(mpnn_n, 1) / 1000.0
        ach = flops_mpnn(n) * B / avg_mpnn_s / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if used_impl == "tcgen05" and os.path.exists(tpath):      # dram bytes per launch from the committed ncu capture
            traffic = json.load(open(tpath))["mpnn_tc_kernel"]["dram_bytes_per_launch"]
        roof = {"bound": "tensor", "kernel": "mpnn_forward_argmax (%s)" % used_impl, "achieved": ach,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"], "traffic": traffic,
                "peak_source": pk["source"] + " (bf16 sustained)", "avg_launch_ms": avg_mpnn_s * 1e3,
                "launches_timed": mpnn_n, "launches": T * args.steps,
                "share_of_step": avg_mpnn_s * 1e3 * T * args.steps / ms_total,
                "flops_per_launch": flops_mpnn(n) * B}
        avg_env_s = env_ms / max(env_n, 1) / 1000.0
        roof_env = {"bound": "hbm", "kernel": "env_step", "achieved": bytes_env(n) * B / avg_env_s / 1e9,
                    "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bytes_env(n) * B / avg_env_s / 1e9 / pk["hbm_gbs"],
                    "avg_launch_us": avg_env_s * 1e6, "launches_timed": env_n,
                    "share_of_step": avg_env_s * 1e3 * T * args.steps / ms_total,
                    "note": "B=4096 moves only %.1f MB per launch: launch-latency bound; see env_only" %
                            (bytes_env(n) * B / 1e6)}
        cpu = None
        if not args.skip_cpu and world == 1:          # side numbers: rank 0 at N = 1 only (the other ranks would wait)
            v, cores, secs = cpu_reference_sample(J[0], wd, 32, 2 * n)
            cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                   "sample": "32 complete episodes (400 env steps each) of one BA-200 graph, batched like the "
                             "reference's test_network (%.1f s of CPU work)" % secs}
        dqn = None
        if not args.skip_dqn and world == 1:
            try:
                dqn = dqn_update_sample()
            except Exception as e:          # a side number must not cost the bench line
                dqn = {"error": "%s: %s" % (type(e).__name__, e)}
        line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16x2-split/f32" if used_impl == "tcgen05" else "f32",
                "data": "synthetic", "config": config_dict(world),
                "e2e": {"value": e2e_value, "unit": "env-steps/s",
                        "h2d_bytes_per_step": int(G * n * n + B * 4 + B * n), "d2h_bytes_per_step": int(B * 4 + B * n),
                        "steps": e2e_steps},
                "gpu_launches": launches, "roofline": roof, "roofline_env_step": roof_env, "env_only": env_only,
                "cpu_baseline": cpu, "clocks": clocks, "mpnn_impl": used_impl, "mean_best_cut": best_mean,
                "dqn_update": dqn}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


PROFILE_STRIDE = 8      # kernel launches bracketed by CUDA events inside the timed region: one in 8 (150 of 1200 per kernel)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mpnn", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-env-only", action="store_true")
    ap.add_argument("--skip-dqn", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
