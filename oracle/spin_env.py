"""Oracle: single-episode Max-Cut spin environment (ECO-DQN configuration), dense numpy, fp64.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, in the reference's own O(N^2)-per-step
arithmetic and operation order (so that fp64 results are bit-identical and so that timing it is a fair
"reference CPU path" baseline):

  reference src/envs/spinsystem.py:183-259  (reset)         -> MaxCutEnv.reset
  reference src/envs/spinsystem.py:283-330  (_reset_state)  -> MaxCutEnv.reset
  reference src/envs/spinsystem.py:355-559  (step)          -> MaxCutEnv.step
  reference src/envs/spinsystem.py:561-574  (get_observation) -> MaxCutEnv.observation
  reference src/envs/score_solver.py:343-419 (MaximumCutUnbiasedScorer) -> the _cut/_gain helpers + normalisers
  reference src/envs/score_solver.py:175-200 (MaximizationProblem score / quality)
  reference src/envs/utils.py:90-102        (calculate_cut / calculate_cut_changes)
  reference src/envs/utils.py:438-464       (HistoryBuffer)  -> VisitedSets

Only the configuration every reference script uses is covered: DEFAULT_OBSERVABLES, RewardSignal.BLS,
norm_rewards=True, ExtraAction.NONE, OptimisationTarget.CUT, SpinBasis.SIGNED, infinite memory,
horizon = max_steps, no stagnation punishment, reversible spins, Stopping.NORMAL; basin_reward is a
parameter (1/N in the experiments, None in the pretrained-agent script).
"""
import numpy as np

N_OBS = 7  # DEFAULT_OBSERVABLES, reference src/envs/utils.py:68-74


def cut_value(spins, J):
    """reference utils.py:90-94: 1/4 * sum_ij J_ij (1 - s_i s_j), with the N x N outer product."""
    return (1 / 4) * np.sum(np.multiply(J, 1 - np.outer(spins, spins)))


def flip_gains(spins, J):
    """reference utils.py:97-102 (numba there): s * (J @ s) = change of the cut if vertex i is flipped."""
    return spins * (J @ spins)


class VisitedSets:
    """reference utils.py:438-464.  Remembers every configuration reached, encoded as the set of vertices
    currently flipped w.r.t. the episode's initial spins.  The initial (empty) set is NOT recorded up
    front (utils.py:440-442), so the first return to it counts as new."""

    def __init__(self):
        self.by_size = {}
        self.current = frozenset()

    def visit(self, vertex):
        nxt = self.current ^ {vertex}
        self.current = nxt
        bucket = self.by_size.setdefault(len(nxt), set())
        if nxt in bucket:
            return False
        bucket.add(nxt)
        return True


class MaxCutEnv:
    def __init__(self, J, max_steps, basin_reward=None, reversible=True, dense_reward=False, min_cut=False):
        """reversible=False, dense_reward=True, basin_reward=None is the S2V-DQN configuration of
        experiments/pretrained_agent/test_s2v.py (observables=[SPIN_STATE], RewardSignal.DENSE, irreversible spins)."""
        self.J = np.asarray(J, dtype=np.float64)
        self.n = self.J.shape[0]
        self.max_steps = int(max_steps)
        self.horizon = self.max_steps                       # spinsystem.py:163
        self.basin_reward = basin_reward
        self.reversible = reversible
        self.dense_reward = dense_reward
        # OptimisationTarget.MIN_CUT (score_solver.py:423-505): every mask is the negated cut change, the quality is
        # normaliser - cut, the normaliser is |sum of the negative weights| over the whole matrix
        self.min_cut = min_cut
        self.state = None

    def _gains(self, spins):
        """get_score_mask / get_solution_quality_mask (score_solver.py:389-413 resp. :472-490)."""
        g = flip_gains(spins, self.J)
        return -g if self.min_cut else g

    # ------------------------------------------------------------------ scorer pieces
    def _quality(self, spins):
        if self.min_cut:                                    # score_solver.py:219-222: max(0, normaliser) - measure
            return max(0, self.qn) - cut_value(spins, self.J)
        # score_solver.py:196-200: measure + |min(0, lower_bound)|
        return cut_value(spins, self.J) + abs(min(0, self.lb))

    def _score(self, spins):
        # score_solver.py:182-188 with is_valid == True and invalidity == 0 for Max-Cut
        return True * self._quality(spins) - 0

    def _nscore(self, spins):
        # score_solver.py:190-194
        return True * self._quality(spins) / self.qn - 0 / 1

    # ------------------------------------------------------------------ reset
    def reset(self, spins=None):
        n, J = self.n, self.J
        self.step_count = 0
        empty = np.array([-1] * n, dtype=np.float64)
        g0 = self._gains(empty)                             # spinsystem.py:200-206
        nz = g0[np.nonzero(g0)]
        if nz.size == 0:
            raise ValueError("graph has no non-zero weighted degree (the reference recurses forever here)")
        self.mlr = np.max(nz)                               # score_solver.py:367-375

        state = np.zeros((N_OBS, n))                        # spinsystem.py:289
        if spins is None:
            if self.reversible:
                state[0, :] = 2 * np.random.randint(2, size=n) - 1    # spinsystem.py:294
            else:
                state[0, :] = -1                                       # spinsystem.py:296-297
        else:
            spins = np.asarray(spins)
            if not np.isin(spins, [-1, 1]).all():           # spinsystem.py:604-606
                raise Exception("SpinSystem is configured for signed spins ([-1,1]).")
            state[0, :] = spins
        gains = self._gains(state[0])
        state[1, :] = gains / self.mlr                      # spinsystem.py:311-312
        state[5, :] = np.sum(gains > 0) / n                 # spinsystem.py:321-322
        self.state = state

        if self.min_cut:
            self.qn = max(1, abs(np.sum(np.multiply(J, (J < 0)))))   # score_solver.py:439-443
        else:
            self.qn = max(1, np.sum(np.multiply(J, (J > 0))) / 2)   # score_solver.py:353-357
        self.lb = min(0, np.sum(np.multiply(J, (J < 0))) / 2)   # score_solver.py:359-365

        s = state[0]
        self.score = self._score(s)                         # spinsystem.py:224-226
        self.nscore = self._nscore(s)
        self.best_score = self.best_obs_score = self.score  # spinsystem.py:234-237
        self.best_nscore = self.best_obs_nscore = self.nscore
        self.best_solution = cut_value(s, J)                # spinsystem.py:241
        self.best_spins = s.copy()
        self.best_obs_spins = s.copy()
        self.visited = VisitedSets() if self.basin_reward is not None else None   # spinsystem.py:256-257
        return self.observation()

    # ------------------------------------------------------------------ step
    def step(self, action):
        n, J = self.n, self.J
        rew = 0
        self.step_count += 1
        if self.step_count > self.max_steps:                # spinsystem.py:365-367
            raise NotImplementedError("environment already done")
        new_state = np.copy(self.state)

        old = self.state[0]
        delta = self._gains(old)[action]                    # spinsystem.py:393
        delta_n = (self._gains(old) / self.qn)[action]      # spinsystem.py:394
        new_state[0, action] = -old[action]
        self.score += delta                                 # spinsystem.py:399-400
        self.nscore += delta_n
        self.state = new_state
        s = new_state[0]
        gains = self._gains(s)                              # spinsystem.py:414-416

        if self.score > self.best_obs_score and not self.dense_reward:   # spinsystem.py:418-424 (BLS, norm_rewards)
            rew = self.nscore - self.best_obs_nscore
        if self.dense_reward:                               # spinsystem.py:435-436 (norm_rewards)
            rew = delta_n
        if self.basin_reward is not None:                   # spinsystem.py:443-457
            is_new = self.visited.visit(int(action))
            if np.all(gains <= 0) and is_new:
                rew += self.basin_reward
        if self.score > self.best_score:                    # spinsystem.py:459-463
            self.best_score = self.score
            self.best_nscore = self.nscore
            self.best_spins = s.copy()
            self.best_solution = cut_value(self.best_spins, J)
        self.best_obs_score = self.best_score               # spinsystem.py:473-477 (infinite memory)
        self.best_obs_nscore = self.best_nscore
        self.best_obs_spins = self.best_spins.copy()

        st = self.state                                     # spinsystem.py:486-535
        st[1, :] = gains / self.mlr
        st[2, :] += (1. / self.max_steps)
        st[2, action] = 0
        st[3, :] = np.abs(self._quality(s) - self._quality(self.best_spins)) / self.mlr
        st[4, :] = np.count_nonzero(self.best_obs_spins - s)
        st[5, :] = np.sum(gains > 0) / n
        st[6, :] = max(0, ((self.step_count - self.max_steps) / self.horizon) + 1)

        done = self.step_count == self.max_steps            # spinsystem.py:541-544
        if not self.reversible and np.count_nonzero(s < 0) == 0:         # spinsystem.py:552-556
            done = True
        return self.observation(), rew, done, None

    # ------------------------------------------------------------------ views
    def observation(self):
        """spinsystem.py:561-574 (SIGNED basis): vstack of the 7 feature rows and the adjacency."""
        return np.vstack((self.state.copy(), self.J))

    @property
    def spins(self):
        return self.state[0]

    def greedy_solve(self):
        """reference src/agents/solver.py:46-67,105-131: flip argmax of the gains until the best gain is < 0
        (zero-gain moves are taken) or the step budget ends.  Returns number of steps taken."""
        done = False
        while not done:
            gains = self._gains(self.state[0])
            if self.reversible:
                a = gains.argmax()
            else:                                           # solver.py:116-121: only spins still at -1
                masked = gains.copy()
                np.putmask(masked, self.state[0] != -1, np.finfo(np.float64).min)
                a = masked.argmax()
            if gains[a] < 0:
                break
            _, _, done, _ = self.step(a)
        return self.step_count


def time_since_flip_table(max_steps):
    """k-fold fp64 accumulation of 1/max_steps starting from 0 (spinsystem.py:493): entry k is the value
    row 2 holds for a vertex whose last reset-to-zero was k steps ago."""
    t = np.zeros(max_steps + 1, dtype=np.float64)
    acc = np.float64(0.0)
    inc = 1. / max_steps
    for k in range(1, max_steps + 1):
        acc = acc + inc
        t[k] = acc
    return t


def immanency_table(max_steps, horizon=None):
    """spinsystem.py:509-511 evaluated for step = 0..max_steps (entry 0 is the reset value 0)."""
    horizon = max_steps if horizon is None else horizon
    t = np.zeros(max_steps + 1, dtype=np.float64)
    for k in range(1, max_steps + 1):
        t[k] = max(0, ((k - max_steps) / horizon) + 1)
    return t
