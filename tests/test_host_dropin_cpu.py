"""CPU checks of the host-side mirror of the reference API (no device work)."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN


def test_generators_reproduce_reference_graphs_for_the_same_seeds():
    from eco_dqn_b200.envs.utils import (RandomErdosRenyiGraphGenerator, RandomBarabasiAlbertGraphGenerator, EdgeType,
                                         SingleGraphGenerator, SetGraphGenerator)
    z = np.load(os.path.join(GOLDEN, "generators.npz"))
    makers = {"er20": lambda: RandomErdosRenyiGraphGenerator(20, 0.15, EdgeType.DISCRETE),
              "er40u": lambda: RandomErdosRenyiGraphGenerator(40, [0.15, 0.02], EdgeType.UNIFORM),
              "ba20": lambda: RandomBarabasiAlbertGraphGenerator(20, 4, EdgeType.DISCRETE),
              "ba60u": lambda: RandomBarabasiAlbertGraphGenerator(60, 4, EdgeType.UNIFORM)}
    for tag, mk in makers.items():
        np.random.seed(123)
        random.seed(123)
        gen = mk()
        got = np.stack([gen.get() for _ in range(3)]).astype(np.int8)
        assert np.array_equal(got, z[tag]), tag
    g = SingleGraphGenerator(z["er20"][0].astype(float))
    assert g.edge_type == EdgeType.DISCRETE and g.n_spins == 20 and g.get() is g.matrix
    assert SingleGraphGenerator(z["ba60u"][0].astype(float)).edge_type == EdgeType.UNIFORM
    s = SetGraphGenerator([m.astype(float) for m in z["er20"]], ordered=True)
    assert [s.get() is s.graphs[i % 3] for i in range(4)] == [True] * 4
    with pytest.raises(NotImplementedError):
        SetGraphGenerator([np.zeros((3, 3)), np.zeros((4, 4))])


def test_mpnn_module_is_checkpoint_compatible_and_matches_golden_q():
    from eco_dqn_b200.networks.mpnn import MPNN
    from oracle.mpnn import KEYS, weights_from_npz
    for name in ("er20_g0", "ba40u_g0", "er200_g0"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        net = MPNN(n_obs_in=7, n_layers=3, n_features=64, tied_weights=False, n_hid_readout=[])
        assert tuple(net.state_dict().keys()) == KEYS
        net.load_state_dict({k: torch.tensor(v) for k, v in weights_from_npz(z).items()})
        net.eval()
        k, J = z["obs"].shape[0], z["J"].astype(np.float32)
        for si in (0, z["obs"].shape[1] // 2):
            obs = torch.tensor(np.concatenate([z["obs"][:, si], np.broadcast_to(J, (k,) + J.shape)], axis=1))
            keep = obs.clone()
            with torch.no_grad():
                q = net(obs).numpy()
            assert torch.equal(obs, keep), "forward must not mutate its argument"
            ref = z["q"][:, si]
            assert np.allclose(q, ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())
    assert net(torch.tensor(np.concatenate([z["obs"][0, 0], J], axis=0))).shape == (J.shape[0],)   # mpnn.py:75 squeeze


def test_mpnn_module_general_weights_route_matches_dense_definition():
    from eco_dqn_b200.networks.mpnn import MPNN
    from oracle.mpnn import mpnn_forward
    torch.manual_seed(0)
    net = MPNN()
    n, B = 12, 3
    A = torch.randn(B, n, n) * (torch.rand(B, n, n) < 0.4)
    A = torch.triu(A, 1)
    A = A + A.transpose(1, 2)
    obs = torch.cat([torch.randn(B, 7, n), A], dim=1)
    ref = mpnn_forward({k: v.detach().numpy() for k, v in net.state_dict().items()}, obs.numpy())
    with torch.no_grad():
        assert torch.allclose(net(obs), ref, rtol=1e-4, atol=1e-5)


def test_unsupported_configurations_raise_at_construction():
    from eco_dqn_b200.envs.spinsystem import check_supported
    from eco_dqn_b200.envs.utils import (DEFAULT_OBSERVABLES, RewardSignal, ExtraAction, OptimisationTarget, SpinBasis,
                                         Stopping, Observable)
    ok = dict(observables=DEFAULT_OBSERVABLES, reward_signal=RewardSignal.BLS, extra_action=ExtraAction.NONE,
              optimisation_target=OptimisationTarget.CUT, spin_basis=SpinBasis.SIGNED, norm_rewards=True,
              memory_length=None, horizon_length=None, stag_punishment=None, reversible_spins=True, init_snap=None,
              stopping=Stopping.NORMAL, max_steps=40)
    check_supported(**ok)
    for key, bad in (("optimisation_target", OptimisationTarget.ENERGY), ("optimisation_target", OptimisationTarget.MIN_COVER),
                     ("reward_signal", RewardSignal.DENSE), ("extra_action", ExtraAction.PASS),
                     ("spin_basis", SpinBasis.BINARY), ("norm_rewards", False), ("memory_length", 10),
                     ("stag_punishment", 0.1), ("reversible_spins", False), ("stopping", Stopping.EARLY),
                     ("observables", [Observable.SPIN_STATE])):
        with pytest.raises(NotImplementedError):
            check_supported(**dict(ok, **{key: bad}))
    with pytest.raises(AssertionError):
        check_supported(**dict(ok, observables=[Observable.TIME_SINCE_FLIP]))


def test_graph_loaders_mc_text_and_pickled_sets(tmp_path):
    """load_graph (GSet `.mc` text + best-known files) and load_graph_set (ndarray / networkx / scipy-sparse pickles) return
    what the reference's loaders return (experiments/utils.py:391-432), including the extra random bit on the solution."""
    import pickle
    import networkx as nx
    import scipy.sparse
    from eco_dqn_b200.experiments.utils import load_graph, load_graph_set, read_mc_instance
    for sub in ("instances", "bkvl", "bksol"):
        (tmp_path / sub).mkdir()
    edges = [(1, 2, 1), (1, 5, -1), (2, 3, 1), (4, 5, -1), (3, 5, 1)]
    (tmp_path / "instances" / "toy.mc").write_text("5 5\n" + "".join("%d %d %d\n" % e for e in edges))
    (tmp_path / "bkvl" / "toy.bkvl").write_text("3\n")
    (tmp_path / "bksol" / "toy.bksol").write_text("0110\n")
    np.random.seed(11)
    extra = np.random.choice([0, 1])
    np.random.seed(11)
    g = load_graph(str(tmp_path), "toy")
    want = np.zeros((5, 5))
    for i, j, w in edges:
        want[i - 1, j - 1] = want[j - 1, i - 1] = w
    assert (g.name, g.n_vertices, g.n_edges, g.bk_val) == ("toy", 5, 5, 3.0)
    assert np.array_equal(g.matrix, want) and g.matrix.dtype == np.float64
    assert g.bk_sol.tolist() == [0, 1, 1, 0, int(extra)]
    n, m, r, c, w = read_mc_instance(str(tmp_path / "instances" / "toy.mc"))
    assert (n, m) == (5, 5) and r.tolist() == [0, 0, 1, 3, 2] and c.tolist() == [1, 4, 2, 4, 4] and w.tolist() == [1, -1, 1, -1, 1]
    (tmp_path / "instances" / "bad.mc").write_text("5\n1 2 1\n")
    with pytest.raises(AssertionError):
        read_mc_instance(str(tmp_path / "instances" / "bad.mc"))
    nxg = nx.Graph()
    nxg.add_nodes_from(range(5))
    nxg.add_weighted_edges_from([(i - 1, j - 1, w) for i, j, w in edges])
    with open(tmp_path / "set.pkl", "wb") as f:
        pickle.dump([want, nxg, scipy.sparse.csr_matrix(want)], f)
    out = load_graph_set(str(tmp_path / "set.pkl"))
    assert len(out) == 3 and all(np.array_equal(o, want) for o in out)
