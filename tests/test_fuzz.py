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
        # lazy masked output (dnn.cc:355-392): any non-zero byte is "active", inactive nodes enter the softmax as exp(0)
        rc = ref.lazy_context(n, batch=int(rng.choice([1, 8, 10])))
        rc.until_output(frames.copy())
        hidden = port.until_output(frames.copy())
        assert np.array_equal(rc.hidden(), hidden), (dims, cutoff, n)
        for idx in rng.choice(n, size=min(n, 3), replace=False):
            mask = rng.choice(np.array([0, 0, 1, 7, -1, -128], dtype=np.int8), size=dims[-1])
            a, b = rc.lazy(int(idx), mask), port.lazy(hidden[int(idx)], mask)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32), b[~np.isnan(b)].view(np.uint32)), (dims, cutoff, n, int(idx))
        rc.close()
        ref.close()
        port.close()


def _mutate(rng, b, texty):
    b = b.copy()
    mode = int(rng.integers(0, 6))
    if mode == 0:  # header bytes
        for _ in range(int(rng.integers(1, 8))):
            b[rng.integers(0, min(len(b), 64))] = rng.integers(0, 256)
    elif mode == 1:  # bytes anywhere (text: characters that mean something to the parser)
        alphabet = np.frombuffer(b"[]<> \n0123456789.e-+xA", dtype=np.uint8)
        for _ in range(int(rng.integers(1, 30))):
            b[rng.integers(0, len(b))] = rng.choice(alphabet) if texty else rng.integers(0, 256)
    elif mode == 2:  # truncation
        b = b[: rng.integers(0, len(b))]
    elif mode == 3 and not texty:  # an extreme big-endian word among the first forty
        w = int(rng.integers(0, min(len(b) // 4, 40)))
        word = np.array([rng.choice([0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, 0x40000000, 0, 1, 0x00FFFFFF, 16, 32])], dtype=">u4")
        b[4 * w:4 * w + 4] = np.frombuffer(word.tobytes(), dtype=np.uint8)
    elif mode == 3:  # a run of text removed
        i = int(rng.integers(0, len(b)))
        b = np.concatenate([b[:i], b[min(len(b), i + int(rng.integers(1, 400))):]])
    elif mode == 4:  # trailing garbage
        b = np.concatenate([b, rng.integers(0, 256, int(rng.integers(1, 100)), dtype=np.uint8)])
    return b  # mode 5: unchanged


def test_host_parsers_under_sanitizers(tmp_path):
    """csrc/model_host.cc (everything the library does with files and foreign blobs before a byte reaches the GPU) built with
    -fsanitize=address,undefined and run over mutated dnn.bin, feature, nnet1 and feature-transform files and corrupted blobs:
    status codes only, no sanitizer report."""
    import shutil
    import subprocess
    from conftest import ROOT
    import os
    import test_formats as tf

    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    exe = str(tmp_path / "host_fuzz_driver")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           os.path.join(ROOT, "tests", "host_fuzz_driver.cc"), os.path.join(ROOT, "fast-dnn_b200", "csrc", "model_host.cc"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr and "cannot find" in r.stderr:
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr[-2000:]
    rng = np.random.default_rng(1)
    layers, shift, scale = synth.make_network((12, 32, 3, 7), seed=5)
    d = tmp_path / "corpus"
    d.mkdir()
    formats.write_dnn_bin(str(d / "base.bin"), layers, shift, scale)
    formats.write_feature_bin(str(d / "fbase.bin"), synth.make_frames(9, 12, seed=1))
    net = np.frombuffer((d / "base.bin").read_bytes(), dtype=np.uint8)
    feat = np.frombuffer((d / "fbase.bin").read_bytes(), dtype=np.uint8)
    nnet = np.frombuffer(tf._kaldi_text(layers).encode(), dtype=np.uint8)
    trans = np.frombuffer(tf._transform_text(shift, scale, True).encode(), dtype=np.uint8)
    count = 300
    for i in range(count):
        (d / f"net_{i}.bin").write_bytes(_mutate(rng, net, False).tobytes())
        (d / f"feat_{i}.bin").write_bytes(_mutate(rng, feat, False).tobytes())
        (d / f"nnet_{i}.txt").write_bytes(_mutate(rng, nnet, True).tobytes())
        (d / f"trans_{i}.txt").write_bytes(_mutate(rng, trans, True).tobytes())
    r = subprocess.run([exe, str(d), str(count)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    assert r.stdout.startswith("accepted: pack "), r.stdout
    # the corpus is not all rejects: unchanged and lightly damaged files still load
    assert int(r.stdout.split()[2]) > count // 4


def test_aligner_refuses_absurd_alignments(tmp_path):
    """an alignment of 2^20 would pad a 40-wide layer into 4 TB of zeros (found by the fuzz: the process grew until the host
    ran out of memory); FeedForwardNetwork.align is only ever called with SIMD widths (FeedForwardNetwork.java:50-58)"""
    layers, shift, scale = synth.make_network((10, 40, 3, 7), seed=5)
    raw, out = str(tmp_path / "raw.bin"), str(tmp_path / "out.bin")
    formats.write_dnn_bin(raw, layers, shift, scale)
    L = qd.lib()
    for ia, ha in [(1 << 20, 16), (4, 1 << 20), (4097, 16), (0, 16), (4, -1)]:
        assert L.fdnn_align_dnn_bin(raw.encode(), out.encode(), ia, ha) == qd.FDNN_EINVAL
    assert L.fdnn_align_dnn_bin(raw.encode(), out.encode(), 4096, 16) == qd.FDNN_OK  # the cap itself is still taken
    got, sh, _ = formats.read_dnn_bin(out)
    assert [w.shape for w, _ in got] == [(48, 4096), (48, 48), (48, 48), (7, 48)] and sh.shape == (4096,)
