#!/usr/bin/env python
"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference/src/cpp by oracle/Makefile).  Run here, where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no model and no known-answer vectors, so these fixtures are
outputs of the reference itself on the seeded synthetic networks / frames / masks of
fast-dnn_b200/synth.py.  Inputs are not stored: they are regenerated from the same seeds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import synth  # noqa: E402
import oracle_py  # noqa: E402

# name → (shape, stress, frames, frame seed)
CASES = {
    "tiny": ("tiny", False, 9, 7),
    "tiny_stress": ("tiny", True, 9, 9),
    "ragged": ("ragged", False, 13, 7),
    "S": ("S", False, 12, 7),
    "S_stress": ("S", True, 6, 9),
}


def main():
    oracle_py.build(port=False, ref=True)
    for name, (shape, stress, n, seed) in CASES.items():
        path = synth.network_file(shape, stress=stress)
        ref = oracle_py.Ref(path)
        frames = synth.make_frames(n, ref.input_dim, seed=seed)
        masks = synth.make_masks(n, ref.output_dim, seed=11)
        trace = ref.hidden_trace(frames)
        ctx = ref.lazy_context(n)
        ctx.until_output(frames)
        lin = ctx.output_linear()
        lazy = np.stack([ctx.lazy(i, masks[i]) for i in range(n)])
        ctx.close()
        out = {
            "multipliers": np.array([ref.qlayer(i)[2] for i in range(ref.qlayer_count)], dtype=np.float32),
            "weights_crc": np.array([int(np.frombuffer(ref.qlayer(i)[0].tobytes(), dtype=np.uint8).astype(np.uint64).dot(
                np.arange(1, ref.qlayer(i)[0].size + 1, dtype=np.uint64) % np.uint64(65521)) % np.uint64(2 ** 61 - 1))
                for i in range(ref.qlayer_count)], dtype=np.uint64),
            "hidden_first": trace[0],
            "hidden_last": trace[-1],
            "output_linear": lin,
            "softmax": ref.calculate(frames),
            "lazy": lazy,
        }
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
    np.save(os.path.join(HERE, "sigmoid_lut.npy"), oracle_py.Ref.sigmoid_lut())


if __name__ == "__main__":
    main()
