"""CPU tests of host-side arithmetic that must equal the reference's bit for bit."""
import warnings

import numpy as np

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
