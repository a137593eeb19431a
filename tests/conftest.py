import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden_cases():
    """ECO-DQN rollout cases (one graph, several attempts)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("s2v_", "mincut_"))
                  and not f.startswith(("multi_", "dqn_")) and f not in ("graphsets.npz", "generators.npz"))


def mincut_cases():
    """ECO-DQN configuration with OptimisationTarget.MIN_CUT (same file layout as golden_cases)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("mincut_"))


def s2v_cases():
    """S2V-DQN cases: irreversible spins, spin-only observation, dense reward, one attempt from all -1."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("s2v_"))
