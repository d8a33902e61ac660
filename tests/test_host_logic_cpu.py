"""CPU tests of host-side arithmetic that must equal the reference's bit for bit."""
import warnings

import numpy as np
import pytest

from multimodal_drl_rmc_b200.agent import epsilon_value


def _reference_epsilon(x, start, end, decay, exp_decay):          # dqn/agent.py:86-90, verbatim arithmetic
    if exp_decay:
        return np.exp(np.interp(x, [0, decay], [np.log(start), np.log(end)]))
    return np.interp(x, [0, decay], [start, end])


def test_epsilon_schedule_is_bit_identical_to_the_reference_formula():
    warnings.simplefilter("ignore")          # log(0) in the degenerate schedules
    rng = np.random.default_rng(0)
    for start, end, decay in ((1.0, 0.01, 2_000_000), (1.0, 0.05, 500_000), (0.9, 0.1, 12_345), (1.0, 0.01, 7), (0.0, 0.0, 1000), (1.0, 0.0, 1000),
                              (1.0, 1.0, 1000)):
        xs = [0, 1, 2, decay - 1, decay, decay + 1, 10 * decay] + [int(v) for v in rng.integers(0, decay + 1, 3000)]
        for exp_decay in (True, False):
            for x in xs:
                got, ref = epsilon_value(x, start, end, decay, exp_decay), _reference_epsilon(x, start, end, decay, exp_decay)
                assert np.float64(got).tobytes() == np.float64(ref).tobytes(), (x, start, end, decay, exp_decay, got, ref)


def test_package_synthetic_generator_equals_the_oracles():
    """bench.py's GPU arm draws its inputs from the package (it must not import oracle/); both generators are one recipe."""
    from multimodal_drl_rmc_b200.synthetic import synthetic_transitions
    from oracle.dqn_oracle import synthetic_transitions as oracle_gen
    for n, d, seed in ((1, 14, 0), (1000, 14, 20251018), (257, 8, 5), (300, 284, 9)):
        for a, b in zip(synthetic_transitions(n, d, seed), oracle_gen(n, d, seed)):
            assert a.dtype == b.dtype and np.array_equal(a, b)


def test_synthetic_vec_env_follows_the_subproc_vec_env_protocol():
    from multimodal_drl_rmc_b200.synthetic import SyntheticVecEnv
    env = SyntheticVecEnv(3, 14, episode_len=4, seed=2)
    obs = env.reset()
    assert obs.shape == (3, 14) and obs.dtype == np.float32
    with pytest.raises(RuntimeError):
        env.step_wait()
    n_done = 0
    for t in range(9):
        env.step_async([1, 2, 3])
        with pytest.raises(RuntimeError):
            env.step_async([0, 0, 0])
        new_obs, rew, done, infos = env.step_wait()
        assert new_obs.shape == (3, 14) and rew.shape == (3,) and done.shape == (3,) and len(infos) == 3
        assert np.allclose(new_obs[:, -1], [2 / 8, 3 / 8, 4 / 8])
        for e in range(3):
            assert (("r" in infos[e]) and infos[e]["l"] == 4) == bool(done[e])
        n_done += int(done.sum())
    assert n_done == 6
    env.close()


def test_bench_gpu_arm_does_not_import_the_oracle():
    """The oracle is test infrastructure: bench.py may execute it only in the cpu_baseline / --impl reference legs."""
    import ast
    import os
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"build_cpu_learner"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        uses = [n for n in ast.walk(fn) if isinstance(n, ast.ImportFrom) and (n.module or "").startswith("oracle")]
        assert not uses or fn.name in allowed, "bench.py::%s imports oracle/" % fn.name
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]
    assert not top
