"""Synthetic workloads (configs, length distribution, batches, candidates, ground truth) live in the
product package (`repurpose_b200/synth.py`) because bench.py and the tools generate their inputs with
them; the tests reach them through this alias next to the oracle they feed."""
from repurpose_b200.synth import *  # noqa: F401,F403
from repurpose_b200.synth import (MAX_SEQ_LEN, MODEL_CFG, TEST_CFG, bias_reg_head, make_batch,  # noqa: F401
                                  make_candidates, make_gt_segments, max_seg_num, sample_lengths)
