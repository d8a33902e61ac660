"""The trainer loop around the learner hot path: ``train.py:83-105`` with the GPU learner step
overlapped with the environment step.

``strict_loop`` is the reference's loop verbatim (choose -> env.step -> store -> learn -> update_target -> log -> save).
``overlapped_loop`` drives a vectorised environment with the asynchronous interface of the reference's ``SubprocVecEnv``
(``dqn/utils/baselines_wrappers/subproc_vec_env.py:71-87`` ``step_async`` / ``step_wait``): after the actions have been
handed to the environment workers, the learner step of this iteration is ENQUEUED on the GPU (``learn()`` +
``update_target_network()`` issue one launch and return; nothing synchronises the stream), so the fused kernel runs while
the simulator steps on the CPU; the transitions are stored when the workers answer.  Per iteration the host waits for the
device exactly once -- for the actions (``choose_actions``), which the environment needs on the host.

What changes with respect to the strict order: the learner step of iteration t draws its minibatch before the transitions of
iteration t are stored (they enter the replay one iteration later: one env step out of >= 100,000 in the buffer).
Everything else -- ``agent.step`` bookkeeping, epsilon, beta, target updates, logging, checkpoints -- is the reference's.
"""
from __future__ import annotations

import itertools


def init_replay_memory_buffer(agent, env, sample_action):
    """train.py:63-81: ``min_buffer_size // n_env`` environment steps with random actions (the last ``resume_step`` of them
    greedy) before learning starts.  ``sample_action()`` stands for ``env.action_space.sample()``."""
    obses = env.reset()
    n = agent.min_buffer_size // agent.n_env
    for t in range(n):
        if t >= n - agent.resume_step:
            actions = agent.choose_actions(obses)
        else:
            actions = [sample_action() for _ in range(agent.n_env)]
        new_obses, rews, dones, _ = env.step(actions)
        agent.store_transitions(obses, actions, rews, dones, new_obses, None)
        obses = new_obses
    return n * agent.n_env


def strict_loop(agent, env, n_iterations, start_step=None):
    """train.py:87-105, bounded to ``n_iterations`` iterations; returns the number of environment steps taken."""
    obses = env.reset()
    first = agent.resume_step if start_step is None else int(start_step)
    for step in itertools.islice(itertools.count(start=first), int(n_iterations)):
        agent.step = step
        actions = agent.choose_actions(obses)
        new_obses, rews, dones, infos = env.step(actions)
        agent.store_transitions(obses, actions, rews, dones, new_obses, infos)
        obses = new_obses
        agent.learn()
        agent.update_target_network()
        agent.log()
        agent.save_model()
    return int(n_iterations) * agent.n_env


def overlapped_loop(agent, env, n_iterations, start_step=None, learn_every=1):
    """The same iteration with the learner step in flight while the environments step (see the module docstring).
    ``learn_every``: learner steps are issued on every ``learn_every``-th iteration (1 = the reference's schedule)."""
    obses = env.reset()
    first = agent.resume_step if start_step is None else int(start_step)
    for k, step in enumerate(itertools.islice(itertools.count(start=first), int(n_iterations))):
        agent.step = step
        actions = agent.choose_actions(obses)            # the iteration's only host<->device synchronisation
        env.step_async(actions)                          # workers step the simulator ...
        if k % learn_every == 0:
            agent.learn()                                # ... while the learner step of this iteration is launched
            agent.update_target_network()                #     (one launch, asynchronous)
        new_obses, rews, dones, infos = env.step_wait()
        agent.store_transitions(obses, actions, rews, dones, new_obses, infos)
        obses = new_obses
        agent.log()
        agent.save_model()
    return int(n_iterations) * agent.n_env
