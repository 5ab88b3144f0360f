"""Seeded fuzz of the two pins the parity claims rest on, on networks the fixed fixtures do not reach: the host packer's
quantizer (csrc/model_host.cc) against the compiled reference's (dnn.cc:460-509), and the plain-C restatement
(oracle/fdnn_oracle.c) against the compiled reference's forward pass (dnn.cc:402-454) — random small aligned networks with
all-zero layers, grid-valued weights (rounding ties), weights at exactly ±cutoff, outliers of 1e30, denormals, cutoffs
from 1e-6 to 1e6, wide bias ranges, hostile frames, every reference batch size.  CPU only; needs /root/reference.

The reference sizes its scratch by layers()[1]->node_count() (dnn.cc:199), so a network needs at least three int8 layers
(two hidden + output) for that to be a hidden width: with fewer the reference overruns its own heap (found by this fuzz),
which is why depths start at three hidden layers here."""
import numpy as np
import pytest

from fast_dnn_b200 import blob as B
from fast_dnn_b200 import formats, synth
from fast_dnn_b200 import quantized_dnn as qd
import oracle_py


def _random_network(rng, wide_scales: bool):
    I = int(rng.choice([4, 8, 12, 20]))
    H = int(rng.choice([16, 32, 48]))
    O = int(rng.integers(1, 40))
    dims = [I] + [H] * int(rng.integers(3, 6)) + [O]
    layers = []
    for j in range(len(dims) - 1):
        k, n = dims[j], dims[j + 1]
        mode = int(rng.integers(0, 8))
        sigma = 10.0 ** (rng.uniform(-4, 1.5) if wide_scales else rng.uniform(-3, 1.0))
        w = rng.normal(0, sigma, (n, k)).astype(np.float32)
        if mode == 1:
            w[rng.random((n, k)) < 0.05] *= 50
        elif mode == 2:
            w[:] = 0  # multiplier = round(127 / 0)
        elif mode == 3:
            w = np.round(w * 4) / 4
        elif mode == 4:
            w[rng.random((n, k)) < 0.1] = np.float32(3.0) * rng.choice([-1, 1])
        elif mode == 5:
            w = rng.integers(-300, 300, (n, k)) / np.float32(rng.choice([1, 2, 7, 63.5, 127]))
        elif mode == 6 and (wide_scales or j > 0):
            w[0, 0] = np.float32(rng.choice([1e30, -1e30]))
        elif mode == 7 and wide_scales:
            w = w * 0 + np.float32(rng.choice([1e-30, -1e-38, 1e-45]))
        bias = rng.normal(0, 10.0 ** rng.uniform(-2, 1.5), n).astype(np.float32)
        layers.append((np.asarray(w, dtype=np.float32), bias))
    shift = rng.normal(0, 0.1, I).astype(np.float32)
    scale = rng.uniform(0.05, 0.1, I).astype(np.float32)
    return dims, layers, shift, scale


@pytest.mark.parametrize("seed", [1, 12])
def test_packer_quantizer_equals_compiled_reference_on_random_networks(seed, tmp_path, have_reference):
    if not have_reference:
        pytest.skip("needs the compiled reference")
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "net.bin")
    for _ in range(40):
        dims, layers, shift, scale = _random_network(rng, wide_scales=True)
        cutoff = float(rng.choice([3.0, 3.0, 0.01, 0.5, 1.0, 10.0, 100.0, 1e-6, 1e6]))
        formats.write_dnn_bin(path, layers, shift, scale)
        ref = oracle_py.Ref(path, cutoff)
        b = B.Blob(qd.pack(path, cutoff))
        assert len(b.qlayers) == ref.qlayer_count
        for i in range(ref.qlayer_count):
            w, bias, mult = b.qlayer(i)
            w2, bias2, mult2 = ref.qlayer(i)
            assert np.array_equal(w, w2) and np.array_equal(bias, bias2), (dims, cutoff, i)
            assert mult == mult2 or (np.isnan(mult) and np.isnan(mult2)), (dims, cutoff, i, mult, mult2)
        ref.close()


@pytest.mark.parametrize("seed", [2, 11])
def test_restatement_equals_compiled_reference_on_random_networks(seed, tmp_path, have_reference):
    if not have_reference:
        pytest.skip("needs the compiled reference")
    rng = np.random.default_rng(seed)
    path = str(tmp_path / "net.bin")
    for _ in range(40):
        dims, layers, shift, scale = _random_network(rng, wide_scales=False)
        cutoff = float(rng.choice([3.0, 3.0, 0.5, 1.0, 10.0]))
        formats.write_dnn_bin(path, layers, shift, scale)
        n = int(rng.integers(1, 40))
        if rng.random() < 0.3:
            frames = synth.make_hostile_frames(n, dims[0], seed=int(rng.integers(1, 1000)))
        else:
            frames = rng.normal(0, 15, (n, dims[0])).astype(np.float32)
        ref, port = oracle_py.Ref(path, cutoff), oracle_py.Port(path, cutoff)
        for x, y in zip(ref.hidden_trace(frames.copy()), port.hidden_trace(frames.copy())):
            assert np.array_equal(x, y), (dims, cutoff, n)
        want = ref.calculate(frames.copy(), batch=int(rng.choice([1, 8, 10, 64])))
        got = port.calculate(frames.copy())
        # bit for bit, NaN rows (a frame with NaN/inf) included
        assert np.array_equal(np.isnan(want), np.isnan(got)), (dims, cutoff, n)
        ok = ~np.isnan(want)
        assert np.array_equal(want[ok].view(np.uint32), got[ok].view(np.uint32)), (dims, cutoff, n)
        ref.close()
        port.close()
