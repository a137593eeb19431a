"""eco_dqn_b200 -- B200-native batched ECO-DQN Max-Cut rollout engine (drop-in for the reference's rollout path).

The directory is named `eco-dqn_b200/`; import it as `eco_dqn_b200` through the shim module at the repo root.
Sub-modules mirror the reference's layout for the rollout path only:

    eco_dqn_b200.envs.core / envs.utils / envs.spinsystem   <- src/envs/*
    eco_dqn_b200.networks.mpnn                              <- src/networks/mpnn.py
    eco_dqn_b200.agents.solver / agents.dqn                 <- src/agents/solver.py, src/agents/dqn/*
    eco_dqn_b200.experiments.utils                          <- experiments/utils.py
    eco_dqn_b200.engine                                     <- the batched device engine underneath
"""
from . import _lib                                    # noqa: F401
from ._lib import lib, LIB_PATH, EcoError             # noqa: F401

__all__ = ["lib", "LIB_PATH", "EcoError"]
