"""GPU parity: the CUDA path, called through the C ABI (libfast-dnn.so), against the CPU oracle on
the same seeded inputs.  Bit-exact for every integer/byte stage (quantized weights, u8 activations
of every layer) and for the fp32 logits; stated tolerance for the softmax scores."""
import os
import threading

import numpy as np
import pytest

from conftest import softmax_close
from fast_dnn_b200 import quantized_dnn as qd
from fast_dnn_b200 import synth
import oracle_py

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def loaded(net_file):
    cache = {}

    def get(shape, stress=False):
        key = (shape, stress)
        if key not in cache:
            path = net_file(shape, stress=stress)
            cache[key] = (qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path))
        return cache[key]

    yield get
    for dnn, _ in cache.values():
        dnn.delete()


def stage_parity(dnn, port, frames):
    """every stage of one forward pass, via a traced context"""
    n = frames.shape[0]
    ctx = dnn.get_new_lazy_context(n)
    try:
        ctx.set_trace(True)
        ctx.calculate_until_output(frames)
        want = port.hidden_trace(frames)  # [qlayers][n][H]; slot j = after hidden layer j
        for layer in range(port.qlayer_count):
            got = ctx.hidden(layer)
            diff = np.flatnonzero(got != want[layer])
            assert diff.size == 0, f"u8 activations after hidden layer {layer}: {diff.size} of {got.size} differ"
        lin = port.output_linear(want[-1])
        _, bias, _ = port.qlayer(port.qlayer_count - 1)
        logits = ctx.logits()
        assert np.array_equal(logits.view(np.uint32), (lin + bias).astype(np.float32).view(np.uint32)), "logits not bit-exact"
        return logits
    finally:
        ctx.delete()


@pytest.mark.parametrize("shape,n", [("tiny", 1), ("tiny", 37), ("ragged", 130), ("S", 128), ("P", 100)])
def test_stages_bit_exact(loaded, shape, n):
    dnn, port = loaded(shape)
    frames = synth.make_frames(n, dnn.input_dimension(), seed=7)
    stage_parity(dnn, port, frames)


@pytest.mark.parametrize("shape,n", [("tiny", 33), ("S", 70)])
def test_stress_network_bit_exact(loaded, shape, n):
    """heavy-tailed weights: int8 wrap-around in the quantizer and dense pmaddubsw saturation"""
    dnn, port = loaded(shape, stress=True)
    assert sum(dnn.fixup_count(i) for i in range(port.qlayer_count)) > 0
    frames = synth.make_frames(n, dnn.input_dimension(), seed=9)
    stage_parity(dnn, port, frames)


@pytest.mark.parametrize("shape,n", [("tiny", 5), ("ragged", 64), ("S", 128), ("P", 100)])
def test_calculate_matches_oracle(loaded, shape, n):
    dnn, port = loaded(shape)
    frames = synth.make_frames(n, dnn.input_dimension(), seed=3)
    before = frames.copy()
    got = dnn.calculate(frames)
    assert np.array_equal(frames, before), "caller's input buffer was modified"
    want = port.calculate(frames)
    softmax_close(got, want)
    assert np.array_equal(np.argmax(got, axis=1), np.argmax(want, axis=1))
    np.testing.assert_allclose(got.sum(axis=1), 1.0, atol=2e-5)


def test_calculate_batch_hint_irrelevant_and_empty(loaded):
    dnn, _ = loaded("tiny")
    frames = synth.make_frames(19, dnn.input_dimension(), seed=5)
    a, b = dnn.calculate(frames, 1), dnn.calculate(frames, 512)
    assert np.array_equal(a, b)
    assert dnn.calculate(np.zeros((0, dnn.input_dimension()), np.float32)).shape == (0, 0)
    with pytest.raises(ValueError):
        dnn.calculate(np.zeros((3, dnn.input_dimension() + 4), np.float32))


def test_chunked_streaming_equals_single_pass(loaded, monkeypatch):
    """n larger than the streaming chunk: two contexts on two streams, ragged last chunk"""
    dnn, port = loaded("S")
    n = 4096 + 4096 + 300  # default chunk is 4096 frames
    frames = synth.make_frames(n, dnn.input_dimension(), seed=21)
    got = dnn.calculate(frames)
    idx = np.r_[0:40, 4090:4110, 8180:8200, n - 20:n]
    want = port.calculate(frames[idx])
    softmax_close(got[idx], want)
    ctx = dnn.get_new_lazy_context(n)
    try:
        ctx.calculate_until_output(frames)
        assert np.array_equal(ctx.hidden(), port.until_output(frames))
    finally:
        ctx.delete()


def test_lazy_context_matches_oracle(loaded):
    dnn, port = loaded("S")
    n = 24
    frames = synth.make_frames(n, dnn.input_dimension(), seed=13)
    masks = synth.make_masks(n, dnn.output_dimension(), ratio=0.40, drift=0.03, seed=11)
    masks[3][masks[3] != 0] = 7  # any non-zero byte is "active" (dnn.cc:369)
    hidden = port.until_output(frames)
    ctx = dnn.get_new_lazy_context(n)
    try:
        ctx.calculate_until_output(frames)
        assert np.array_equal(ctx.hidden(), hidden)
        rows = [ctx.calculate_for_output_nodes(masks[i]) for i in range(n)]
        batch = ctx.calculate_for_output_nodes_batch(masks)
    finally:
        ctx.delete()
    for i in range(n):
        want = port.lazy(hidden[i], masks[i])
        softmax_close(rows[i], want)
        assert np.array_equal(rows[i], batch[i])
        inactive = masks[i] == 0
        # inactive outputs come back as 1/total, not 0 (dnn.cc:367-370 + 534-544)
        assert np.all(rows[i][inactive] == rows[i][inactive][0]) and rows[i][inactive][0] > 0
        active = np.flatnonzero(~inactive)
        assert np.array_equal(np.argsort(-rows[i][active], kind="stable")[:10], np.argsort(-want[active], kind="stable")[:10])


def test_tensor_core_and_dp4a_kernels_agree(net_file):
    """same layer through tcgen05 and through the dp4a kernel: identical bytes"""
    path = net_file("S")
    frames = synth.make_frames(200, 440, seed=2)
    outs = []
    for force in ("0", "1"):
        os.environ["FDNN_FORCE_SIMT"] = force
        try:
            dnn = qd.QuantizedDnn.load_from_file(path)
        finally:
            os.environ.pop("FDNN_FORCE_SIMT", None)
        assert dnn.uses_tensor_cores(0) == (force == "0")
        ctx = dnn.get_new_lazy_context(200)
        ctx.calculate_until_output(frames)
        outs.append((ctx.hidden(), ctx.logits()))
        ctx.delete()
        dnn.delete()
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))


def test_model_accessors(loaded):
    dnn, port = loaded("S")
    assert (dnn.input_dimension(), dnn.output_dimension(), dnn.layer_count()) == (440, 2000, 5)
    # layerDimension (jni_dnn.cc:135-148): 0 → layer-0 nodes; i ≥ 1 → nodes of file layer i+1
    assert [dnn.layer_dimension(i) for i in range(6)] == [512, 512, 512, 2000, -1, -1]
    for i in range(port.qlayer_count):
        w, b, m = dnn.qlayer(i)
        w2, b2, m2 = port.qlayer(i)
        assert np.array_equal(w, w2) and np.array_equal(b, b2) and m == m2


def test_shared_model_many_threads(loaded):
    """MultiThreadedStressTest.java:48-67: one immutable model, concurrent calculate() calls"""
    dnn, port = loaded("S")
    frames = synth.make_frames(96, dnn.input_dimension(), seed=17)
    want = dnn.calculate(frames)
    errors = []

    def work(seed):
        try:
            rng = np.random.default_rng(seed)
            for _ in range(6):
                perm = rng.permutation(96)[: int(rng.integers(1, 96))]
                got = dnn.calculate(frames[perm], 10)
                if not np.array_equal(got, want[perm]):
                    errors.append((seed, "mismatch"))
        except Exception as e:  # surfaced by the assert below instead of dying silently in the thread
            errors.append((seed, repr(e)))

    threads = [threading.Thread(target=work, args=(s,)) for s in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors


def test_headline_network_batch512(loaded):
    """BASELINE configs[2]: 7×2048 hidden, 8000 outputs, batch 512 — last-hidden bytes and logits
    bit-exact against the oracle, scores within tolerance, plus size-independent properties."""
    dnn, port = loaded("L")
    frames = synth.make_frames(512, 440, seed=7)
    ctx = dnn.get_new_lazy_context(512)
    try:
        ctx.calculate_until_output(frames)
        hidden = ctx.hidden()
        want_hidden = port.until_output(frames, threads=os.cpu_count() or 8)
        assert np.array_equal(hidden, want_hidden)
        rows = np.r_[0:24, 500:512]
        lin = port.output_linear(want_hidden[rows])
        _, bias, _ = port.qlayer(port.qlayer_count - 1)
        assert np.array_equal(ctx.logits()[rows].view(np.uint32), (lin + bias).astype(np.float32).view(np.uint32))
        masks = synth.make_masks(512, 8000, seed=11)
        lazy = ctx.calculate_for_output_nodes_batch(masks)
        for r in (0, 255, 511):
            softmax_close(lazy[r], port.lazy(want_hidden[r], masks[r]))
    finally:
        ctx.delete()
    got = dnn.calculate(frames)
    np.testing.assert_allclose(got.sum(axis=1), 1.0, atol=5e-5)
    # a frame's result does not depend on its neighbours or its position in the batch
    again = dnn.calculate(frames[::-1].copy())
    assert np.array_equal(again[::-1], got)


def test_cluster_multicast_path_is_bit_exact(net_file):
    """FDNN_CLUSTER=1 (TMA multicast across thread-block clusters, opt-in): same bytes as the default path"""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import fast_dnn_b200\n"
        "from fast_dnn_b200 import quantized_dnn as qd, synth\n"
        "dnn = qd.QuantizedDnn.load_from_file(%r)\n"
        "x = synth.make_frames(300, 440, seed=4)\n"
        "ctx = dnn.get_new_lazy_context(300); ctx.calculate_until_output(x)\n"
        "np.save(sys.argv[1], ctx.hidden()); np.save(sys.argv[2], ctx.logits())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), net_file("S"))
    outs = {}
    for flag in ("0", "1"):
        paths = [f"/tmp/fdnn_cluster_{flag}_{k}.npy" for k in ("h", "l")]
        env = dict(os.environ, FDNN_CLUSTER=flag)
        subprocess.run([sys.executable, "-c", code] + paths, check=True, env=env, timeout=240)
        outs[flag] = [np.load(p) for p in paths]
    assert np.array_equal(outs["0"][0], outs["1"][0])
    assert np.array_equal(outs["0"][1].view(np.uint32), outs["1"][1].view(np.uint32))


@pytest.mark.parametrize("stress", [False, True])
def test_cta_pair_path_is_bit_exact(net_file, stress):
    """FDNN_PAIR=<tile width>: every int8 layer on CTA pairs (tcgen05 cta_group::2, qlayer_pair.cu) — same
    bytes as the single-CTA path (itself checked against the oracle above): one frame, ragged row counts around
    the 128-row CTA and 256-row pair boundaries, several tiles per pair, dense saturation (stress network)"""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import fast_dnn_b200\n"
        "from fast_dnn_b200 import quantized_dnn as qd, synth\n"
        "dnn = qd.QuantizedDnn.load_from_file(%r)\n"
        "hs, ls = [], []\n"
        "for n in (1, 129, 300, 1000):\n"
        "    x = synth.make_frames(n, 440, seed=4 + n)\n"
        "    ctx = dnn.get_new_lazy_context(n); ctx.calculate_until_output(x)\n"
        "    hs.append(ctx.hidden().copy()); ls.append(ctx.logits().copy()); ctx.delete()\n"
        "np.save(sys.argv[1], np.concatenate(hs)); np.save(sys.argv[2], np.concatenate(ls))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), net_file("S", stress=stress))
    outs = {}
    for flag in ("0", "64", "128", "256"):
        paths = [f"/tmp/fdnn_pair_{int(stress)}_{flag}_{k}.npy" for k in ("h", "l")]
        env = dict(os.environ, FDNN_PAIR=flag)
        subprocess.run([sys.executable, "-c", code] + paths, check=True, env=env, timeout=240)
        outs[flag] = [np.load(p) for p in paths]
    for flag in ("64", "128", "256"):
        assert np.array_equal(outs["0"][0], outs[flag][0]), f"pair tile {flag}: last-hidden bytes differ"
        assert np.array_equal(outs["0"][1].view(np.uint32), outs[flag][1].view(np.uint32)), f"pair tile {flag}: logits differ"


def test_headline_network_stream_chunk_uses_pairs(loaded):
    """BASELINE configs[4] regime: a 3100-frame chunk (ragged against the 256-row pair tiles) of the 7×2048/8000
    network — every int8 layer takes the CTA-pair kernel by default here; last-hidden bytes and logits bit-exact
    against the oracle, scores within tolerance."""
    dnn, port = loaded("L")
    n = 3100
    frames = synth.make_frames(n, 440, seed=21)
    ctx = dnn.get_new_lazy_context(n)
    try:
        ctx.calculate_until_output(frames)
        hidden = ctx.hidden()
        want_hidden = port.until_output(frames, threads=os.cpu_count() or 8)
        assert np.array_equal(hidden, want_hidden)
        rows = np.r_[0:8, 250:262, 3090:3100]
        lin = port.output_linear(want_hidden[rows])
        _, bias, _ = port.qlayer(port.qlayer_count - 1)
        assert np.array_equal(ctx.logits()[rows].view(np.uint32), (lin + bias).astype(np.float32).view(np.uint32))
    finally:
        ctx.delete()
    got = dnn.calculate(frames[:700])
    for r in (0, 255, 256, 699):
        softmax_close(got[r], port.calculate(frames[r:r + 1])[0])


def test_throughput_tile_policy_gives_identical_results(net_file):
    """fdnn_set_tile_policy: wider tiles for callers that keep several contexts in flight — same bytes as the default policy"""
    dnn = qd.QuantizedDnn.load_from_file(net_file("L"))
    try:
        frames = synth.make_frames(512, 440, seed=33)
        outs = []
        for policy in ("latency", "throughput"):
            dnn.set_tile_policy(policy)
            ctx = dnn.get_new_lazy_context(512)
            ctx.calculate_until_output(frames)
            outs.append((ctx.hidden().copy(), ctx.logits().copy()))
            ctx.delete()
        assert np.array_equal(outs[0][0], outs[1][0])
        assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32))
        with pytest.raises(qd.FdnnError):
            qd._check(qd.lib().fdnn_set_tile_policy(dnn._h, 7))
    finally:
        dnn.delete()


def test_certified_input_layer_equals_exact_kernel_on_hostile_frames(net_file):
    """csrc/input_tc.cu (tensor-core certificate + exact fix-up) against csrc/input_layer.cu (FDNN_INPUT_TC=0) on frames
    with NaN, ±inf, huge, tiny, zero and constant rows: every layer-0 byte and the logits must be identical, and the
    certificate must have decided most of the ordinary elements itself"""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import fast_dnn_b200\n"
        "from fast_dnn_b200 import quantized_dnn as qd, synth\n"
        "dnn = qd.QuantizedDnn.load_from_file(%r)\n"
        "x = synth.make_frames(300, 440, seed=12)\n"
        "x[3, 7] = np.nan; x[5, :] = 0.0; x[9, 100] = np.inf; x[11, 5] = -np.inf; x[13, :] = 1e30; x[17, :] = 1e-30\n"
        "x[19, :] = 7.25; x[23, 0] = 3e38; x[29, ::2] = -1e-38; x[31, :] *= 1e4; x[37, :] *= 1e-6\n"
        "ctx = dnn.get_new_lazy_context(300); ctx.set_trace(True); ctx.calculate_until_output(x)\n"
        "u = ctx.input_undecided()\n"
        "np.save(sys.argv[1], ctx.hidden(0)); np.save(sys.argv[2], ctx.logits()); np.save(sys.argv[3], np.array([-1 if u is None else u]))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), net_file("S"))
    outs = {}
    for flag in ("0", "1"):
        paths = [f"/tmp/fdnn_itc_{flag}_{k}.npy" for k in ("h0", "l", "u")]
        env = dict(os.environ, FDNN_INPUT_TC=flag)
        subprocess.run([sys.executable, "-c", code] + paths, check=True, env=env, timeout=240)
        outs[flag] = [np.load(p) for p in paths]
    assert int(outs["0"][2][0]) == -1 and int(outs["1"][2][0]) >= 0, "the two runs did not take the two different paths"
    assert np.array_equal(outs["0"][0], outs["1"][0]), "layer-0 bytes differ"
    assert np.array_equal(outs["0"][1].view(np.uint32), outs["1"][1].view(np.uint32)), "logits differ"
    assert int(outs["1"][2][0]) < 0.1 * 300 * 512, "the certificate left more than 10 % of the elements to the exact path"


def test_block_fixup_wide_layer_long_batch_dense_blocks(net_file):
    """The block fix-up kernel's corner cases against the exact kernel (FDNN_INPUT_TC=0) and against the warp-per-frame
    fix-up (FDNN_FIXUP=warp): a hidden layer wider than one compaction pass (2304 > 2048 nodes), a batch long enough for a
    CTA to own all of them (4800 frames ≥ 32 per SM), a ragged last block, single NaN frames (every node listed with one
    element: the maximum number of batches) and a whole block of NaN frames (every word all ones: full batches of 64
    chunks, nothing for the producer to stand in for)."""
    import subprocess
    import sys
    shape = (40, 2304, 2, 64)
    n = 4800 + 7
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); import fast_dnn_b200\n"
        "from fast_dnn_b200 import quantized_dnn as qd, synth\n"
        "dnn = qd.QuantizedDnn.load_from_file(%r)\n"
        "x = synth.make_frames(%d, 40, seed=21)\n"
        "x[5, 3] = np.nan; x[40, 0] = np.inf; x[64:96, 1] = np.nan; x[4799, 2] = np.nan; x[4806, :] = 0.0\n"
        "ctx = dnn.get_new_lazy_context(%d); ctx.set_trace(True); ctx.calculate_until_output(x)\n"
        "np.save(sys.argv[1], ctx.hidden(0)); np.save(sys.argv[2], ctx.logits())\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), net_file(shape), n, n)
    outs = {}
    for name, env_add in (("exact", {"FDNN_INPUT_TC": "0"}), ("block", {}), ("warp", {"FDNN_FIXUP": "warp"})):
        paths = [f"/tmp/fdnn_fb_{name}_{k}.npy" for k in ("h0", "l")]
        subprocess.run([sys.executable, "-c", code] + paths, check=True, env=dict(os.environ, **env_add), timeout=240)
        outs[name] = [np.load(p) for p in paths]
    for name in ("block", "warp"):
        rows, cols = np.nonzero(outs["exact"][0] != outs[name][0])
        assert rows.size == 0, (f"layer-0 bytes differ ({name}): {rows.size} elements, rows {np.unique(rows)[:20]} "
                                f"(blocks {np.unique(rows // 32)[:20]}), cols {np.unique(cols)[:20]} … {np.unique(cols)[-5:]}")
        assert np.array_equal(outs["exact"][1].view(np.uint32), outs[name][1].view(np.uint32)), f"logits differ ({name})"


@pytest.mark.parametrize("hidden", [128, 256, 384, 768])
@pytest.mark.parametrize("stress", [False, True])
def test_pipeline_geometries_bit_exact(net_file, hidden, stress):
    """Hidden widths of 1, 2, 3 and 6 K blocks of 128 bytes: one K block per tile (shorter than any ring), the fused kernel's
    two-K-block stages with a single stage (256) and with an odd number of stages (768), and a width the fused kernel does not take
    (384 = 1.5 stages → layer by layer).  Traced stages against the oracle with the layer-by-layer kernels; the fused kernel and
    CTA pairs, where they apply, against those; dense pmaddubsw saturation with the heavy-tailed weights."""
    shape = (40, hidden, 3, 300)
    path = net_file(shape, stress=stress)
    dnn, port = qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path)
    try:
        assert all(dnn.uses_tensor_cores(i) for i in range(dnn.layer_count() - 1))
        for n in (1, 129, 640):
            frames = synth.make_frames(n, dnn.input_dimension(), seed=700 + n)
            os.environ["FDNN_FUSED"] = "0"
            logits = stage_parity(dnn, port, frames)
            scores = dnn.calculate(frames)
            softmax_close(scores, port.calculate(frames))
            for env in ({"FDNN_FUSED": "2"}, {"FDNN_FUSED": "0", "FDNN_PAIR": "64"}, {"FDNN_FUSED": "0", "FDNN_PAIR": "256"}):
                os.environ.update(env)
                ctx = dnn.get_new_lazy_context(n)
                try:
                    ctx.calculate_until_output(frames)
                    assert np.array_equal(ctx.hidden(), port.until_output(frames)), f"{env}: last-hidden bytes, {n} frames"
                    assert np.array_equal(ctx.logits().view(np.uint32), logits.view(np.uint32)), f"{env}: logits, {n} frames"
                finally:
                    ctx.delete()
                assert np.array_equal(dnn.calculate(frames).view(np.uint32), scores.view(np.uint32)), f"{env}: scores, {n} frames"
                os.environ.pop("FDNN_PAIR", None)
    finally:
        os.environ.pop("FDNN_FUSED", None)
        os.environ.pop("FDNN_PAIR", None)
        dnn.delete()


@pytest.mark.parametrize("shape,policy", [("tiny", "latency"), ("S", "latency"), ("L", "latency"), ("L", "throughput")])
def test_hostile_frames_match_oracle(net_file, shape, policy):
    """NaN, ±inf, huge, tiny, zero and constant rows against the oracle (which tests/test_oracle.py pins on the compiled reference for
    the same frames): every hidden layer's bytes, the logits' bits (NaN for NaN), the scores within tolerance where they are finite"""
    path = net_file(shape)
    dnn, port = qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path)
    dnn.set_tile_policy(policy)
    try:
        frames = synth.make_hostile_frames(40, dnn.input_dimension())
        ctx = dnn.get_new_lazy_context(frames.shape[0])
        try:
            ctx.calculate_until_output(frames)
            hidden = port.until_output(frames)
            assert np.array_equal(ctx.hidden(), hidden)
            _, bias, _ = port.qlayer(port.qlayer_count - 1)
            want_logits = (port.output_linear(hidden) + bias).astype(np.float32)
            got_logits = ctx.logits()
            assert np.array_equal(np.isnan(got_logits), np.isnan(want_logits))
            ok = ~np.isnan(want_logits)
            assert np.array_equal(got_logits[ok].view(np.uint32), want_logits[ok].view(np.uint32))
        finally:
            ctx.delete()
        got, want = dnn.calculate(frames), port.calculate(frames)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        softmax_close(got[ok], want[ok])
    finally:
        dnn.delete()


def test_lazy_mask_edge_cases(loaded):
    """LazyOutputActivations (dnn.cc:355-392) at the edges of its mask argument: nothing active (every score 1/O exactly as the
    reference computes it: exp(0) summed O times), everything active (the same scores as calculate()), negative and large mask
    bytes (any non-zero byte is active), a single active node, and the per-frame index running through the whole context"""
    dnn, port = loaded("S")
    n, O = 6, dnn.output_dimension()
    frames = synth.make_frames(n, dnn.input_dimension(), seed=17)
    hidden = port.until_output(frames)
    masks = np.zeros((n, O), np.int8)
    masks[1, :] = 1
    masks[2, :] = -1
    masks[3, ::3] = -128
    masks[3, 1::3] = 127
    masks[4, 1234] = 1
    masks[5, :] = np.where(np.arange(O) % 2 == 0, 5, 0)
    full = dnn.calculate(frames)
    ctx = dnn.get_new_lazy_context(n)
    try:
        ctx.calculate_until_output(frames)
        rows = [ctx.calculate_for_output_nodes(masks[i]) for i in range(n)]
        batch = ctx.calculate_for_output_nodes_batch(masks)
    finally:
        ctx.delete()
    for i in range(n):
        want = port.lazy(hidden[i], masks[i])
        softmax_close(rows[i], want)
        assert np.array_equal(rows[i].view(np.uint32), batch[i].view(np.uint32))
    assert np.all(rows[0] == rows[0][0]) and abs(float(rows[0][0]) * O - 1.0) < 1e-5
    assert np.array_equal(rows[1].view(np.uint32), full[1].view(np.uint32)), "all-active mask must give calculate()'s scores"
    assert np.array_equal(rows[2].view(np.uint32), full[2].view(np.uint32)), "negative mask bytes are active"
