"""Host-side mirror of reference src/envs/utils.py for the rollout path: the enums every caller passes as env
kwargs, the observable list, and the graph generators that supply dense adjacency matrices.

Graph supply stays on the host: generators return numpy arrays exactly as the
reference's do and draw from the same global RNGs (numpy for weights / ER probability, python `random` inside
networkx), so a script that seeds them gets the same graphs.  Only the generators the reference's scripts use
are provided: Erdos-Renyi, Barabasi-Albert, Single and Set (reference utils.py:165-236, 319-382).
"""
import random
from abc import ABC, abstractmethod
from enum import Enum

import networkx as nx
import numpy as np


class Stopping(Enum):            # reference utils.py:10-14
    NORMAL = 1
    QUARTER = 2
    EARLY = 3


class EdgeType(Enum):            # :16-19
    UNIFORM = 1
    DISCRETE = 2
    RANDOM = 3


class RewardSignal(Enum):        # :21-26
    DENSE = 1
    BLS = 2
    SINGLE = 3
    CUSTOM_BLS = 4


class ExtraAction(Enum):         # :28-32
    PASS = 1
    RANDOMISE = 2
    NONE = 3


class OptimisationTarget(Enum):  # :34-41
    CUT = 1
    ENERGY = 2
    MIN_COVER = 3
    MIN_CUT = 4
    MAX_IND_SET = 5
    MAX_CLIQUE = 6
    MIN_DOM_SET = 7


class SpinBasis(Enum):           # :43-46
    SIGNED = 1
    BINARY = 2


class Observable(Enum):          # :48-66
    SPIN_STATE = 1
    IMMEDIATE_QUALITY_CHANGE = 2
    IMMEDIATE_VALIDITY_DIFFERENCE = 3
    IMMEDIATE_VALIDITY_CHANGE = 4
    TIME_SINCE_FLIP = 5
    EPISODE_TIME = 6
    TERMINATION_IMMANENCY = 7
    NUMBER_OF_QUALITY_IMPROVEMENTS = 8
    NUMBER_OF_VALIDITY_IMPROVEMENTS = 9
    DISTANCE_FROM_BEST_SOLUTION = 10
    DISTANCE_FROM_BEST_STATE = 11
    GLOBAL_VALIDITY_DIFFERENCE = 12
    VALIDITY_BIT = 13


DEFAULT_OBSERVABLES = [Observable.SPIN_STATE,                      # :68-74 -- the 7 rows the kernels produce
                       Observable.IMMEDIATE_QUALITY_CHANGE,
                       Observable.TIME_SINCE_FLIP,
                       Observable.DISTANCE_FROM_BEST_SOLUTION,
                       Observable.DISTANCE_FROM_BEST_STATE,
                       Observable.NUMBER_OF_QUALITY_IMPROVEMENTS,
                       Observable.TERMINATION_IMMANENCY]


def calculate_cut(spins, matrix):
    """Cut value of a +-1 assignment (reference utils.py:90-94), computed as (sum J - s^T J s) / 4."""
    spins = np.asarray(spins, dtype=np.float64)
    matrix = np.asarray(matrix, dtype=np.float64)
    return 0.25 * (matrix.sum() - spins @ matrix @ spins)


def calculate_cut_changes(spins, matrix):
    """Change of the cut when vertex i is flipped: s_i (J s)_i (reference utils.py:97-102)."""
    spins = np.asarray(spins, dtype=np.float64)
    return spins * (np.asarray(matrix, dtype=np.float64) @ spins)


class GraphGenerator(ABC):
    """reference utils.py:105-123 (padding is a no-op there, utils.py:112-116, and is not offered here)."""

    def __init__(self, n_spins, edge_type, biased=False):
        self.n_spins = n_spins
        self.edge_type = edge_type
        self.biased = biased

    @abstractmethod
    def get(self, with_padding=False):
        raise NotImplementedError


def _weight_mask(n, edge_type):
    """Symmetric weight pattern multiplied onto a 0/1 adjacency; RNG consumption as the reference's
    get_connection_mask closures (utils.py:175-190)."""
    if edge_type == EdgeType.UNIFORM:
        return np.ones((n, n))
    if edge_type == EdgeType.DISCRETE:
        m = 2. * np.random.randint(2, size=(n, n)) - 1.
    elif edge_type == EdgeType.RANDOM:
        m = 2. * np.random.rand(n, n) - 1
    else:
        raise NotImplementedError()
    return np.tril(m) + np.triu(m.T, 1)


def _weighted(graph, n, edge_type):
    adj = np.multiply(nx.to_numpy_array(graph), _weight_mask(n, edge_type))
    np.fill_diagonal(adj, 0)
    return adj


class RandomErdosRenyiGraphGenerator(GraphGenerator):
    """reference utils.py:165-202."""

    def __init__(self, n_spins=20, p_connection=[0.1, 0], edge_type=EdgeType.DISCRETE):
        super().__init__(n_spins, edge_type, False)
        if type(p_connection) not in [list, tuple]:
            p_connection = [p_connection, 0]
        assert len(p_connection) == 2, "p_connection must have length 2"
        self.p_connection = p_connection

    def get(self, with_padding=False):
        p = np.clip(np.random.normal(*self.p_connection), 0, 1)
        return _weighted(nx.erdos_renyi_graph(self.n_spins, p), self.n_spins, self.edge_type)


class RandomBarabasiAlbertGraphGenerator(GraphGenerator):
    """reference utils.py:204-236."""

    def __init__(self, n_spins=20, m_insertion_edges=4, edge_type=EdgeType.DISCRETE):
        super().__init__(n_spins, edge_type, False)
        self.m_insertion_edges = m_insertion_edges

    def get(self, with_padding=False):
        return _weighted(nx.barabasi_albert_graph(self.n_spins, self.m_insertion_edges), self.n_spins, self.edge_type)


def _classify(matrices):
    if all(np.isin(m, [0, 1]).all() for m in matrices):
        return EdgeType.UNIFORM
    if all(np.isin(m, [0, -1, 1]).all() for m in matrices):
        return EdgeType.DISCRETE
    return EdgeType.RANDOM


class SingleGraphGenerator(GraphGenerator):
    """reference utils.py:319-345 (unbiased graphs only on this path)."""

    def __init__(self, matrix, bias=None):
        if bias is not None:
            raise NotImplementedError("biased graphs are outside the accelerated path")
        super().__init__(matrix.shape[0], _classify([matrix]), False)
        self.matrix = matrix
        self.bias = None

    def get(self, with_padding=False):
        return self.matrix


class SetGraphGenerator(GraphGenerator):
    """reference utils.py:347-382: hands out the graphs of a fixed set in order, or one at random per call."""

    def __init__(self, matrices, biases=None, ordered=False):
        if biases is not None:
            raise NotImplementedError("biased graphs are outside the accelerated path")
        if len(set(m.shape[0] for m in matrices)) != 1:
            raise NotImplementedError("All graphs in SetGraphGenerator must have the same dimension.")
        super().__init__(matrices[0].shape[0], _classify(matrices), False)
        self.graphs = matrices
        self.ordered = ordered
        if ordered:
            self.i = 0

    def get(self, with_padding=False):
        if self.ordered:
            m = self.graphs[self.i]
            self.i = (self.i + 1) % len(self.graphs)
            return m
        return random.sample(self.graphs, k=1)[0]


class HistoryBuffer:
    """Host-side visited-configuration set with the reference's semantics (utils.py:438-464); the device keeps the
    same information as a 128-bit Zobrist key per configuration (env_kernels.cu)."""

    def __init__(self):
        self._seen = set()
        self._current = frozenset()

    def update(self, action):
        self._current = self._current ^ {action}
        if self._current in self._seen:
            return False
        self._seen.add(self._current)
        return True
