"""Pin the CPU oracle against trajectories recorded from the unmodified reference (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden_cases
from oracle.spin_env import MaxCutEnv, time_since_flip_table, immanency_table
from oracle.mpnn import weights_from_npz, mpnn_forward
from oracle.rollout import rollout, greedy_baseline


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def basin(z):
    b = float(z["basin_reward"])
    return None if b < 0 else b


@pytest.mark.parametrize("name", golden_cases())
def test_env_teacher_forced_bit_exact(name):
    z = load(name)
    J, T = z["J"].astype(np.float64), int(z["T"])
    out = rollout(J, None, z["init_spins"], T, basin(z), forced_actions=z["actions"], record_obs=True)
    assert np.array_equal(out["rewards"].view(np.uint64), z["rewards"].view(np.uint64)), "fp64 rewards differ"
    assert np.array_equal(out["scores"], z["scores"])
    assert np.array_equal(out["best_cut"], z["best_cut"])
    assert np.array_equal(out["best_spins"], z["best_spins"])
    assert np.array_equal(out["final_spins"], z["final_spins"])
    k = z["obs"].shape[0]
    got = out["obs"][:k][:, z["obs_steps"]]
    # `==` on fp32 (treats -0.0 == 0.0, like the survey's protocol), plus a sign check on the spin row
    assert np.array_equal(got, z["obs"])


@pytest.mark.parametrize("name", golden_cases())
def test_scalars(name):
    z = load(name)
    e = MaxCutEnv(z["J"].astype(np.float64), int(z["T"]), basin(z))
    e.reset(z["init_spins"][0])
    assert e.mlr == float(z["mlr"]) and e.qn == float(z["qn"]) and e.lb == float(z["lb"])
    assert e.score == z["init_score"][0] and e.best_solution == z["init_cut"][0]


@pytest.mark.parametrize("name", golden_cases())
def test_mpnn_q_values(name):
    z = load(name)
    w = weights_from_npz(z)
    J = z["J"].astype(np.float32)
    k, ns = z["obs"].shape[:2]
    for si in range(ns):
        obs = np.concatenate([z["obs"][:, si], np.broadcast_to(J, (k,) + J.shape)], axis=1)
        q = mpnn_forward(w, obs).numpy()
        ref = z["q"][:, si]
        # the golden Q was computed with all n_attempts episodes in the batch (same graph => same norm.max())
        assert np.allclose(q, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max()), (name, si)
        assert np.array_equal(q.argmax(1), ref.argmax(1))


@pytest.mark.parametrize("name", ["er20_g0", "er20_g3_nobasin", "ba40u_g0"])
def test_free_running_rollout_matches_reference(name):
    z = load(name)
    w = weights_from_npz(z)
    torch.set_num_threads(4)
    out = rollout(z["J"].astype(np.float64), w, z["init_spins"], int(z["T"]), basin(z))
    assert np.array_equal(out["actions"], z["actions"])
    assert np.array_equal(out["best_cut"], z["best_cut"])


@pytest.mark.parametrize("name", golden_cases())
def test_greedy_baseline(name):
    z = load(name)
    cuts, spins, steps = greedy_baseline(z["J"].astype(np.float64), z["init_spins"], int(z["T"]), basin(z))
    assert np.array_equal(cuts, z["greedy_cuts"])
    assert np.array_equal(spins, z["greedy_spins"])
    assert np.array_equal(steps, z["greedy_steps"])
    c1, s1, _ = greedy_baseline(z["J"].astype(np.float64), -np.ones((1, int(z["n"])), dtype=np.int8), int(z["T"]), basin(z))
    assert c1[0] == float(z["greedy_single_cut"]) and np.array_equal(s1[0], z["greedy_single_spins"])


def test_tables_match_repeated_addition():
    for T in (40, 80, 400):
        t = time_since_flip_table(T)
        acc = 0.0
        for k in range(1, T + 1):
            acc += 1. / T
            assert t[k] == acc
        im = immanency_table(T)
        assert im[0] == 0 and im[T] == 1.0 and im[1] == max(0, ((1 - T) / T) + 1)


def test_known_answer_upper_bounds():
    """reference README.md:82 -- opts/cuts_* are best-known cuts: no rollout may exceed them."""
    gs = np.load(os.path.join(GOLDEN, "graphsets.npz"))
    for name in golden_cases():
        z = load(name)
        if name.startswith("er20_g") and "nobasin" not in name:
            gi = int(name.split("_g")[1])
            assert np.array_equal(gs["er20"][gi], z["J"])
            assert z["best_cut"].max() <= gs["er20_opt"][gi]


# ---------------------------------------------------------------------------------------------------------------
# S2V-DQN configuration (irreversible spins, spin-only observation, dense reward): SURVEY.md section 8(f)3
# ---------------------------------------------------------------------------------------------------------------
from conftest import s2v_cases            # noqa: E402
from oracle.rollout import rollout_s2v   # noqa: E402


@pytest.mark.parametrize("name", s2v_cases())
def test_s2v_teacher_forced_bit_exact(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    qs = []
    out = rollout_s2v(z["J"].astype(np.float64), weights_from_npz(z), int(z["T"]), forced_actions=z["actions"],
                      q_hook=lambda t, q: qs.append(q.numpy()[0]))
    assert np.array_equal(out["rewards"], z["rewards"])          # numerically (a zero-gain flip gives -0.0 in both)
    assert np.array_equal(np.signbit(out["rewards"]), np.signbit(z["rewards"]))
    assert np.array_equal(out["scores"], z["scores"]) and np.array_equal(out["dones"], z["dones"])
    assert out["best_cut"] == float(z["best_cut"]) and np.array_equal(out["best_spins"], z["best_spins"])
    q = np.stack(qs)
    assert np.allclose(q, z["q"], rtol=1e-4, atol=1e-5 * np.abs(z["q"]).max())


@pytest.mark.parametrize("name", s2v_cases())
def test_s2v_free_running_and_greedy(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = rollout_s2v(z["J"].astype(np.float64), weights_from_npz(z), int(z["T"]))
    assert np.array_equal(out["actions"], z["actions"]) and out["best_cut"] == float(z["best_cut"])
    env = MaxCutEnv(z["J"].astype(np.float64), int(z["T"]), None, reversible=False, dense_reward=True)
    env.reset(np.array([-1] * int(z["n"])))
    steps = env.greedy_solve()
    assert env.best_solution == float(z["greedy_cut"]) and steps == int(z["greedy_steps"])
    assert np.array_equal(env.best_spins.astype(np.int8), z["greedy_spins"])


# ---------------------------------------------------------------------------------------------------------------
# OptimisationTarget.MIN_CUT: ECO-DQN configuration, minimisation scorer
# ---------------------------------------------------------------------------------------------------------------
from conftest import mincut_cases            # noqa: E402


@pytest.mark.parametrize("name", mincut_cases())
def test_mincut_env_scalars_greedy_and_rollout(name):
    z = load(name)
    assert int(z["min_cut"]) == 1
    J, T = z["J"].astype(np.float64), int(z["T"])
    e = MaxCutEnv(J, T, basin(z), min_cut=True)
    e.reset(z["init_spins"][0])
    assert e.mlr == float(z["mlr"]) and e.qn == float(z["qn"]) and e.lb == float(z["lb"])
    assert e.score == z["init_score"][0] and e.best_solution == z["init_cut"][0]
    out = rollout(J, None, z["init_spins"], T, basin(z), forced_actions=z["actions"], record_obs=True, min_cut=True)
    assert np.array_equal(out["rewards"].view(np.uint64), z["rewards"].view(np.uint64)), "fp64 rewards differ"
    assert np.array_equal(out["scores"], z["scores"]) and np.array_equal(out["best_cut"], z["best_cut"])
    assert np.array_equal(out["best_spins"], z["best_spins"])
    k = z["obs"].shape[0]
    assert np.array_equal(out["obs"][:k][:, z["obs_steps"]], z["obs"])
    cuts, spins, steps = greedy_baseline(J, z["init_spins"], T, basin(z), min_cut=True)
    assert np.array_equal(cuts, z["greedy_cuts"]) and np.array_equal(spins, z["greedy_spins"])
    assert np.array_equal(steps, z["greedy_steps"])
    c1, s1, _ = greedy_baseline(J, -np.ones((1, int(z["n"])), dtype=np.int8), T, basin(z), min_cut=True)
    assert c1[0] == float(z["greedy_single_cut"]) and np.array_equal(s1[0], z["greedy_single_spins"])
    free = rollout(J, weights_from_npz(z), z["init_spins"], T, basin(z), min_cut=True)
    assert np.array_equal(free["actions"], z["actions"]) and np.array_equal(free["best_cut"], z["best_cut"])


# ---- round 2 -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["multi_er20", "multi_er40"])
def test_oracle_reproduces_reference_test_network_over_several_graphs(name):
    """The reference batches the attempts of ONE graph at a time, so the degree feature is normalised per graph
    (mpnn.py:102); the oracle rollout, called per graph, follows the reference's actions / rewards / cuts bit for bit."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    w = weights_from_npz(z)
    assert len(set(z["max_degree"].tolist())) > 1
    for j, J in enumerate(z["graphs"]):
        n = J.shape[0]
        out = rollout(J.astype(np.float64), w, z["init_spins"][j], 2 * n, 1.0 / n)
        assert np.array_equal(out["actions"], z["actions"][j])
        assert np.array_equal(out["rewards"].view(np.uint64), z["rewards"][j].view(np.uint64))
        assert np.array_equal(out["scores"], z["scores"][j])
        assert np.array_equal(out["best_cut"], z["cuts"][j])
        assert np.array_equal(out["best_spins"], z["sols"][j])
        assert out["best_cut"].max() == z["res_cut"][j] and out["best_cut"].mean() == z["res_mean_cut"][j]


def test_blocked_oracle_forward_equals_dense_forward():
    """mpnn_forward_blocked (used for N > 500 on the GPU box) against the as-written dense forward."""
    from oracle.mpnn import mpnn_forward_blocked
    for name in ("er200_g0", "ba60_g2", "er20_g0"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        w = weights_from_npz(z)
        J = z["J"].astype(np.float32)
        for s in (0, z["obs"].shape[1] // 2):
            rows = z["obs"][0, s]                                       # [7, n]
            dense = mpnn_forward(w, np.concatenate([rows, J], axis=0)[None]).numpy()[0]
            for rb in (7, 64):
                blocked = mpnn_forward_blocked(w, rows.T, J, rows_per_block=rb).numpy()
                assert np.abs(blocked - dense).max() <= 2e-6 * np.abs(dense).max(), (name, s, rb)
            assert np.abs(dense - z["q"][0, s]).max() <= 1e-5 * np.abs(dense).max()
