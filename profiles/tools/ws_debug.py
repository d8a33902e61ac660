"""Which seeds of the large-batch parity cases are free of ReLU mask flips (a pre-activation within fp32 rounding of zero whose
sign differs between two summation orders changes one sample's contribution to a whole gradient column by O(1/B))?"""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from tests import parity_utils as PU
for algo, D, B, soft, tf in (("DQNAgent", 20, 5000, False, 2), ("PerDuelingDoubleDQNAgent", 14, 4808, True, 30000)):
    for seed in (11, 12, 13, 14, 15):
        res = PU.run_parity_case(algo, D, B, 8192, 8192, 2, seed=seed, soft=soft, target_freq=tf)
        print(algo, B, "seed", seed, "grads %.2g (%s) weights %.2g q %.2g" % (res["max_rel_grads"], res["worst_grad"], res["max_rel_weights"], res["max_rel_q"]), flush=True)
