"""The CPU oracle is pinned before it is trusted: the plain-C restatement (oracle/fdnn_oracle.c)
against (a) the committed golden vectors, which are outputs of the unmodified reference, and
(b) the compiled reference itself when /root/reference is present (this container)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE_ROOT
from fast_dnn_b200 import formats, synth
import oracle_py

sys_path_golden = os.path.join(GOLDEN, "make_golden.py")


def golden_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", sys_path_golden)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES


CASES = golden_cases()


def weights_crc(w):
    b = np.frombuffer(w.tobytes(), dtype=np.uint8).astype(np.uint64)
    return int(b.dot(np.arange(1, b.size + 1, dtype=np.uint64) % np.uint64(65521)) % np.uint64(2 ** 61 - 1))


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_reproduces_reference_golden_vectors(name, net_file):
    shape, stress, n, seed = CASES[name]
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    port = oracle_py.Port(net_file(shape, stress=stress))
    frames = synth.make_frames(n, port.input_dim, seed=seed)
    masks = synth.make_masks(n, port.output_dim, seed=11)
    for i in range(port.qlayer_count):
        w, _, m = port.qlayer(i)
        assert m == g["multipliers"][i]
        assert weights_crc(w) == int(g["weights_crc"][i])
    trace = port.hidden_trace(frames)
    assert np.array_equal(trace[0], g["hidden_first"])
    assert np.array_equal(trace[-1], g["hidden_last"])
    assert np.array_equal(port.until_output(frames, threads=3), g["hidden_last"])
    lin = port.output_linear(trace[-1])
    assert np.array_equal(lin.view(np.uint32), g["output_linear"].view(np.uint32))
    # glibc expf on both sides: the port's softmax is bit-identical to the reference's
    assert np.array_equal(port.calculate(frames).view(np.uint32), g["softmax"].view(np.uint32))
    lazy = np.stack([port.lazy(trace[-1][i], masks[i]) for i in range(n)])
    assert np.array_equal(lazy.view(np.uint32), g["lazy"].view(np.uint32))


def test_lut_matches_reference_golden():
    assert np.array_equal(oracle_py.Port.sigmoid_lut(), np.load(os.path.join(GOLDEN, "sigmoid_lut.npy")))


def test_qsigmoid_edges():
    q = oracle_py.Port.qsigmoid
    lut = oracle_py.Port.sigmoid_lut()
    assert q(-6.40) == 0 and q(-6.395) == 0 and q(6.40) == 255 and q(6.395) == 255 and q(1e9) == 0  # int overflow → INT_MIN bucket
    assert q(float("nan")) == 0 and q(float("inf")) == 0 and q(float("-inf")) == 0
    assert q(0.0) == lut[640] == 128
    assert q(0.005) == lut[641]          # round half away from zero: 0.5 → 1
    assert q(-0.005) == lut[639]
    assert q(0.0049) == lut[640]
    for k in range(-645, 646, 7):
        x = np.float32(k) / np.float32(100.0)
        kk = int(np.round(np.float32(x * np.float32(100.0)) + (0.5 if x >= 0 else -0.5) - (0.5 if x >= 0 else -0.5)))
        kk = int(np.trunc(np.float32(x * np.float32(100.0)) + np.copysign(0.5, x)))
        want = 0 if kk <= -640 else 255 if kk >= 640 else lut[kk + 640]
        assert q(float(x)) == want, (k, float(x))


def test_node_sum_saturates_like_pmaddubsw():
    a = np.full(16, 255, np.uint8)
    w = np.full(16, 127, np.int8)
    assert oracle_py.Port.node_sum(a, w) == 8 * 32767
    assert oracle_py.Port.node_sum(a, w, saturate=False) == 16 * 255 * 127
    w[:] = -128
    assert oracle_py.Port.node_sum(a, w) == 8 * -32768
    a[:] = 1
    assert oracle_py.Port.node_sum(a, w) == -16 * 128


def test_saturation_fires_on_synthetic_networks(net_file):
    """the int16 clamp is live on our fixtures, so parity tests do exercise the fix-up path"""
    for shape, stress, floor in (("S", True, 1000), ("S", False, 1)):
        port = oracle_py.Port(net_file(shape, stress=stress))
        frames = synth.make_frames(16, port.input_dim, seed=9)
        trace = port.hidden_trace(frames)
        events = 0
        for layer in range(port.qlayer_count):
            w = port.qlayer(layer)[0].astype(np.int32)
            for f in range(16):
                prod = trace[layer][f].astype(np.int32)[None, :] * w
                pair = prod[:, 0::2] + prod[:, 1::2]
                events += int(((pair > 32767) | (pair < -32768)).sum())
        assert events >= floor, (shape, stress, events)


needs_ref = pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="/root/reference not present (GPU box)")


@needs_ref
@pytest.mark.parametrize("shape,stress,n", [("tiny", False, 21), ("tiny", True, 21), ("ragged", False, 40), ("S", False, 24), ("S", True, 10)])
def test_port_equals_compiled_reference(shape, stress, n, net_file):
    path = net_file(shape, stress=stress)
    port, ref = oracle_py.Port(path), oracle_py.Ref(path)
    assert (port.input_dim, port.output_dim, port.hidden_dim, port.qlayer_count) == (ref.input_dim, ref.output_dim, ref.hidden_dim, ref.qlayer_count)
    for a, b in zip(port.input_layer(), ref.input_layer()):
        assert np.array_equal(a, b)
    for i in range(port.qlayer_count):
        (w, b, m), (w2, b2, m2) = port.qlayer(i), ref.qlayer(i)
        assert np.array_equal(w, w2) and np.array_equal(b, b2) and m == m2
    frames = synth.make_frames(n, port.input_dim, seed=5)
    assert np.array_equal(port.hidden_trace(frames), ref.hidden_trace(frames, batch=10))
    for batch in (1, 8, 10, 64):
        assert np.array_equal(port.calculate(frames).view(np.uint32), ref.calculate(frames, batch=batch).view(np.uint32))
    masks = synth.make_masks(n, port.output_dim, seed=3)
    ctx = ref.lazy_context(n)
    ctx.until_output(frames)
    hidden = ctx.hidden()
    assert np.array_equal(hidden, port.until_output(frames))
    assert np.array_equal(ctx.output_linear().view(np.uint32), port.output_linear(hidden).view(np.uint32))
    for i in range(0, n, 3):
        assert np.array_equal(ctx.lazy(i, masks[i]).view(np.uint32), port.lazy(hidden[i], masks[i]).view(np.uint32))
    ctx.close()


@needs_ref
@pytest.mark.parametrize("shape", ["tiny", "S"])
def test_port_equals_compiled_reference_on_hostile_frames(shape, net_file):
    """NaN, ±inf, huge, tiny, zero and constant rows: what the reference does with them is x86 arithmetic (cvttss2si's INT_MIN for NaN
    and out-of-range products → bucket 0, dnn.h:35-42); the restatement must do the same, byte for byte and NaN for NaN"""
    path = net_file(shape)
    port, ref = oracle_py.Port(path), oracle_py.Ref(path)
    frames = synth.make_hostile_frames(40, port.input_dim)
    assert np.array_equal(port.hidden_trace(frames), ref.hidden_trace(frames, batch=10))
    got, want = port.calculate(frames), ref.calculate(frames, batch=10)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)].view(np.uint32), want[~np.isnan(want)].view(np.uint32))


@needs_ref
def test_lut_and_qsigmoid_equal_compiled_reference():
    assert np.array_equal(oracle_py.Port.sigmoid_lut(), oracle_py.Ref.sigmoid_lut())
    xs = np.concatenate([np.linspace(-7, 7, 2801, dtype=np.float32), np.float32([1e9, -1e9, 3e7, 2.1474836e7, np.inf, -np.inf, np.nan, 0.005, -0.005, 0.015])])
    for x in xs:
        assert oracle_py.Port.qsigmoid(float(x)) == oracle_py.Ref.qsigmoid(float(x)), float(x)


@needs_ref
def test_config0_shipped_features_through_synthetic_network(net_file):
    """BASELINE configs[0]: data/8khz.aligned.bin holds 389×432 FEATURES (not a network); first 100
    frames through the synthetic 432-input network, reference vs port."""
    feats = formats.read_feature_bin(os.path.join(REFERENCE_ROOT, "data", "8khz.aligned.bin"))
    assert feats.shape == (389, 432)
    path = net_file("P")
    port, ref = oracle_py.Port(path), oracle_py.Ref(path)
    x = feats[:100]
    assert np.array_equal(port.calculate(x).view(np.uint32), ref.calculate(x, batch=10).view(np.uint32))
    short = formats.read_feature_bin(os.path.join(REFERENCE_ROOT, "data", "16khz.bin"))
    assert short.shape == (100, 432)  # the file holds a 101st row that the header does not count


@pytest.mark.parametrize("key", ["khz8", "khz16"])
def test_config0_golden_real_features(key, net_file):
    """BASELINE configs[0] without /root/reference: the committed shipped-feature rows and the bytes the unmodified
    reference produced for them (tests/golden/cfg0.npz, make_golden.py) against the port"""
    g = np.load(os.path.join(GOLDEN, "cfg0.npz"))
    x = g[key]
    port = oracle_py.Port(net_file("P"))
    trace = port.hidden_trace(x)
    assert np.array_equal(trace[0], g[key + "_hidden_first"]) and np.array_equal(trace[-1], g[key + "_hidden_last"])
    assert np.array_equal(port.output_linear(trace[-1])[::10].view(np.uint32), g[key + "_linear_rows"].view(np.uint32))
    sm = port.calculate(x)
    assert np.array_equal(sm[::10].view(np.uint32), g[key + "_softmax_rows"].view(np.uint32))
    assert np.array_equal(sm.argmax(axis=1), g[key + "_argmax"])
