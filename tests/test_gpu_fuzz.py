"""Random small networks with adversarial weights, biases, cutoffs and frames through the C ABI on the GPU against the plain-C
restatement (tools/gpu_fuzz.py): depths from two int8 layers (which the reference itself cannot run, dnn.cc:199) to five, output
widths 1…299, all-zero layers, weights of 1e30, rounding ties, hostile frames.  Last-hidden bytes and logits bit for bit, NaN rows in
the same places, scores within the stated tolerance.  The first 1055 networks of seed 1 ran green on a B200
(profiles/r2_gpu_fuzz.json); this test replays the first 150 of them."""
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.gpu
def test_random_networks_through_the_c_abi_equal_the_oracle():
    import gpu_fuzz

    stats = gpu_fuzz.run(seed=1, seconds=120.0, max_networks=150)
    assert stats["errors_count"] == 0, stats["errors"]
    assert stats["networks"] == 150
    for kind in ("hidden_mismatch", "logits_mismatch", "nan_pattern_mismatch", "score_tolerance"):
        assert stats[kind + "_count"] == 0, (kind, stats[kind])
