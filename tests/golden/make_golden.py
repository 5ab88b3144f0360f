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
    make_cfg0()


def make_cfg0():
    """BASELINE configs[0]: the reference's shipped FEATURE files (data/8khz.aligned.bin rows 0-99, data/16khz.bin; both
    432-dim, real speech features, range ≈ [−73, 126] with three zero pad columns) through the synthetic 432-input network P
    on the unmodified reference.  The frames themselves are stored (they cannot be regenerated on the GPU box, where
    /root/reference does not exist) together with the reference's bytes for them."""
    from fast_dnn_b200 import formats
    ref = oracle_py.Ref(synth.network_file("P"))
    out = {}
    for key, name, rows in (("khz8", "8khz.aligned.bin", 100), ("khz16", "16khz.bin", 100)):
        x = formats.read_feature_bin(os.path.join("/root/reference/data", name))[:rows]
        trace = ref.hidden_trace(x)
        ctx = ref.lazy_context(rows)
        ctx.until_output(x)
        lin = ctx.output_linear()
        ctx.close()
        sm = ref.calculate(x)
        out[key] = x
        out[key + "_hidden_first"] = trace[0]
        out[key + "_hidden_last"] = trace[-1]
        out[key + "_argmax"] = sm.argmax(axis=1).astype(np.int32)
        out[key + "_softmax_rows"] = sm[::10]            # every tenth row in full
        out[key + "_linear_rows"] = lin[::10]            # pre-bias dequantized output-layer sums of the same rows
    np.savez_compressed(os.path.join(HERE, "cfg0.npz"), **out)
    print("cfg0", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
