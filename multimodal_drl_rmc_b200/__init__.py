"""B200-native drop-in for the DQN learner hot path of youcefMehamlia/Multimodal-DRL-RMC.

Same surface as the reference's ``dqn`` package for this path (dqn/__init__.py:1-6):

    from multimodal_drl_rmc_b200 import Agents, Networks
    agent = Agents.PerDuelingDoubleDQNAgent(n_env=..., ...)      # dqn/agent.py:275-320
    agent.store_transitions(...); agent.learn(); agent.update_target_network()

Compute lives in ``librmc_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/rmc_b200.h``); importing this package never falls back to a CPU implementation.
"""
from . import agent as Agents
from . import network as Networks
from .replay_memory import ReplayMemoryNaive, ReplayMemoryPrioritized, SumTree
from ._lib import build_library, lib

__all__ = ["Agents", "Networks", "ReplayMemoryNaive", "ReplayMemoryPrioritized", "SumTree", "build_library", "lib"]
