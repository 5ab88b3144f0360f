"""GPU parity on the configurations BASELINE.json names, row by row (through the C ABI, against the CPU oracle and the
committed reference outputs), plus the soak that pins the certified input layer and the multi-GPU parity cases.

  configs[0]  shipped real features (tests/golden/cfg0.npz) through the 432-input network
  configs[2]  7×2048 / 8000, batch 512: every logits row bit-exact, every softmax row within the stated tolerance
  configs[3]  the same batch through the lazy API with the drifting 40 % masks, per frame and batched, every row
  configs[4]  frame shards over ranks / over the devices of one handle give the bytes of the single-GPU result
"""
import ctypes as C
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, softmax_close
from fast_dnn_b200 import quantized_dnn as qd
from fast_dnn_b200 import synth
import jni_fake
import oracle_py

pytestmark = pytest.mark.gpu

CORES = os.cpu_count() or 8


def order_consistent(got, want, k=10):
    """the k best outputs come in the reference's order wherever the reference separates neighbours by more than the
    stated tolerance (the logits are bit-exact, so only exp/sum rounding can reorder near-ties)"""
    order = np.argsort(-want, kind="stable")[: k + 1]
    for a, b in zip(order, order[1:]):
        if want[a] - want[b] > 1e-9 + 4e-5 * want[a]:
            assert got[a] > got[b], (a, b, got[a], got[b], want[a], want[b])


@pytest.fixture(scope="module")
def headline(net_file):
    path = net_file("L")
    dnn = qd.QuantizedDnn.load_from_file(path)
    port = oracle_py.Port(path)
    frames = synth.make_frames(512, 440, seed=7)
    hidden = port.until_output(frames, threads=CORES)
    _, bias, _ = port.qlayer(port.qlayer_count - 1)
    logits = (port.output_linear(hidden, threads=CORES) + bias).astype(np.float32)
    yield dnn, port, frames, hidden, logits
    dnn.delete()


def test_config2_every_row_of_the_headline_batch(headline):
    dnn, port, frames, hidden, logits = headline
    ctx = dnn.get_new_lazy_context(512)
    try:
        ctx.calculate_until_output(frames)
        assert np.array_equal(ctx.hidden(), hidden)
        got_logits = ctx.logits()
    finally:
        ctx.delete()
    assert np.array_equal(got_logits.view(np.uint32), logits.view(np.uint32)), "logits of the 512×8000 batch are not bit-exact"
    got = dnn.calculate(frames)
    want = port.calculate(frames, threads=CORES)
    softmax_close(got, want)
    assert np.array_equal(got.argmax(axis=1), want.argmax(axis=1))
    for r in range(512):
        order_consistent(got[r], want[r])
    np.testing.assert_allclose(got.sum(axis=1), 1.0, atol=5e-5)
    # the pooled workspace is reused: same bytes on the second and third call (the third replays a captured graph)
    assert np.array_equal(dnn.calculate(frames), got) and np.array_equal(dnn.calculate(frames), got)


def test_config3_lazy_every_row_both_entry_points(headline):
    """LazyOutputActivations (dnn.cc:355-392) with the mask protocol of FuncTest.java:121-154: 40 % active, 3 % drift"""
    dnn, port, frames, hidden, logits = headline
    masks = synth.make_masks(512, 8000, ratio=0.40, drift=0.03, seed=11)
    masks[7][masks[7] != 0] = -3  # any non-zero byte is "active" (dnn.cc:369)
    ctx = dnn.get_new_lazy_context(512)
    try:
        ctx.calculate_until_output(frames)
        rows = np.stack([ctx.calculate_for_output_nodes(masks[i]) for i in range(512)])
        batch = ctx.calculate_for_output_nodes_batch(masks)
        assert ctx.current_vector_index == 512
    finally:
        ctx.delete()
    assert np.array_equal(rows.view(np.uint32), batch.view(np.uint32)), "per-frame and batched lazy outputs differ"
    for i in range(512):
        want = port.lazy(hidden[i], masks[i])
        softmax_close(rows[i], want)
        inactive = masks[i] == 0
        assert inactive.sum() == 8000 - 3200
        assert np.all(rows[i][inactive] == rows[i][inactive][0]) and rows[i][inactive][0] > 0  # 1/total, not 0
        active = np.flatnonzero(~inactive)
        assert active[np.argmax(rows[i][active])] == active[np.argmax(want[active])]
        order_consistent(rows[i][active], want[active])


@pytest.mark.parametrize("key", ["khz8", "khz16"])
def test_config0_shipped_features_stage_exact(net_file, key):
    """rows 0-99 of data/8khz.aligned.bin and all of data/16khz.bin (real speech features, range ≈ [−73, 126], three zero pad
    columns) through the 432-input network: every stage against the port, and against the bytes the unmodified reference
    produced for them (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLDEN, "cfg0.npz"))
    x = g[key]
    assert x.shape == (100, 432) and not x[:, 429:].any()
    path = net_file("P")
    dnn, port = qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path)
    try:
        ctx = dnn.get_new_lazy_context(100)
        ctx.set_trace(True)
        ctx.calculate_until_output(x)
        want = port.hidden_trace(x)
        for layer in range(port.qlayer_count):
            assert np.array_equal(ctx.hidden(layer), want[layer]), f"u8 activations after hidden layer {layer}"
        assert np.array_equal(ctx.hidden(0), g[key + "_hidden_first"]) and np.array_equal(ctx.hidden(), g[key + "_hidden_last"])
        _, bias, _ = port.qlayer(port.qlayer_count - 1)
        logits = ctx.logits()
        assert np.array_equal(logits.view(np.uint32), (port.output_linear(want[-1]) + bias).astype(np.float32).view(np.uint32))
        assert np.array_equal(logits[::10].view(np.uint32), (g[key + "_linear_rows"] + bias).astype(np.float32).view(np.uint32))
        ctx.delete()
        got = dnn.calculate(x)
        softmax_close(got, port.calculate(x))
        softmax_close(got[::10], g[key + "_softmax_rows"])
        assert np.array_equal(got.argmax(axis=1), g[key + "_argmax"])
    finally:
        dnn.delete()


def _load_both_input_paths(path):
    """the same network twice: certified tensor-core input layer (default) and the exact CUDA-core kernel"""
    prev = os.environ.get("FDNN_INPUT_TC")
    try:
        os.environ["FDNN_INPUT_TC"] = "1"
        tc = qd.QuantizedDnn.load_from_file(path)
        os.environ["FDNN_INPUT_TC"] = "0"
        exact = qd.QuantizedDnn.load_from_file(path)
    finally:
        if prev is None:
            os.environ.pop("FDNN_INPUT_TC", None)
        else:
            os.environ["FDNN_INPUT_TC"] = prev
    return tc, exact


def test_certified_input_layer_soak_one_million_frames(net_file):
    """csrc/input_tc.cu decides ≈ 97 % of the layer-0 bytes from an error-bound certificate instead of the reference's
    arithmetic; a wrong constant in that bound would flip about one byte in 1e8.  1 048 576 frames × 2048 nodes = 2.1e9
    elements through both paths (FDNN_INPUT_TC=1 / =0), byte-identical chunk by chunk (device-side checksums of layer 0 and of
    the last hidden layer): synthetic N(0, 15²) frames at scales 1e-3 … 1e3, offsets, the shipped real features tiled at several
    scales, and hostile rows (NaN, ±inf, huge, denormal, constant)."""
    import torch
    dev = torch.device("cuda", 0)
    tc, exact = _load_both_input_paths(net_file("L"))
    chunk, n_chunks = 16384, 64
    base = torch.from_numpy(synth.make_frames(chunk, 440, seed=77)).to(dev)
    g = np.load(os.path.join(GOLDEN, "cfg0.npz"))
    real = np.zeros((200, 440), np.float32)
    real[:100, :432], real[100:, :432] = g["khz8"], g["khz16"]
    real = torch.from_numpy(np.tile(real, (chunk // 200 + 1, 1))[:chunk]).to(dev)
    row_gain = torch.linspace(0.5, 2.0, chunk, device=dev).unsqueeze(1)
    scales = [1.0, 1e-3, 1e-2, 0.1, 0.5, 2.0, 10.0, 100.0, 1e3, 0.03, 0.3, 3.0, 30.0, 7e-4, 1.7, 0.77]
    ctx_tc, ctx_ex = tc.get_new_lazy_context(chunk), exact.get_new_lazy_context(chunk)
    undecided = 0
    try:
        ctx_tc.set_trace(True)
        ctx_ex.set_trace(True)
        for c in range(n_chunks):
            s = scales[c % len(scales)]
            if c % 4 == 3:  # the 200 real frames repeat within a chunk: a per-row gain makes every copy a different frame
                x = real * (s * row_gain) + (0.0 if c % 8 == 3 else 0.37 * c)
            else:
                x = base * s + (c // len(scales)) * 2.5 * s
                x = torch.roll(x, shifts=c, dims=1)  # a different frame ↔ weight-row pairing in every chunk
            if c == 5:
                x = x.clone()
                x[3, 7], x[9, 100], x[11, 5] = float("nan"), float("inf"), float("-inf")
                x[13, :], x[17, :], x[19, :], x[23, 0] = 1e30, 1e-30, 7.25, 3e38
                x[29, ::2] = -1e-38
                x[64:96, 1] = float("nan")
            x = x.contiguous()
            ctx_tc.until_output_device(x.data_ptr(), chunk)
            ctx_ex.until_output_device(x.data_ptr(), chunk)
            torch.cuda.synchronize()
            und = ctx_tc.input_undecided()
            assert und is not None and ctx_ex.input_undecided() is None, "the two models did not take the two different paths"
            undecided += und
            for layer in (0, None):
                a, b = ctx_tc.hidden_digest(layer), ctx_ex.hidden_digest(layer)
                if a != b:
                    ha, hb = ctx_tc.hidden(layer), ctx_ex.hidden(layer)
                    rows, cols = np.nonzero(ha != hb)
                    pytest.fail(f"chunk {c} (scale {s}) layer {layer}: {rows.size} bytes differ, first (frame {rows[0]}, node {cols[0]}): "
                                f"{ha[rows[0], cols[0]]} vs {hb[rows[0], cols[0]]}")
    finally:
        ctx_tc.delete()
        ctx_ex.delete()
        tc.delete()
        exact.delete()
    frac = undecided / (n_chunks * chunk * 2048)
    print(f"soak: {n_chunks * chunk} frames, {n_chunks * chunk * 2048:.3g} elements, undecided fraction {frac:.4f}")
    assert 0.0 < frac < 0.25


def test_soak_digest_detects_a_single_byte(net_file):
    """the checksum the soak relies on: equal to a host recomputation, and different for a one-byte difference"""
    dnn = qd.QuantizedDnn.load_from_file(net_file("S"))
    ctx = dnn.get_new_lazy_context(300)
    try:
        ctx.calculate_until_output(synth.make_frames(300, 440, seed=1))
        h = ctx.hidden().reshape(-1).astype(np.uint64)
        with np.errstate(over="ignore"):
            w = (np.arange(h.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)) | np.uint64(1)
            want = int((h * w).sum(dtype=np.uint64))
            assert ctx.hidden_digest() == want
            h[12345] ^= np.uint64(1)
            assert int((h * w).sum(dtype=np.uint64)) != want
    finally:
        ctx.delete()
        dnn.delete()


@pytest.mark.parametrize("shape,stress,policy", [("S", False, "latency"), ("S", True, "latency"), ("L", False, "latency"), ("L", False, "throughput"),
                                                 ("ragged", False, "latency")])
def test_fused_kernel_equals_layer_by_layer(net_file, shape, stress, policy):
    """csrc/qlayer_fused.cu (all int8 layers + softmax in one persistent kernel, the default for batches that are one wave of
    tiles) against the layer-by-layer kernels (FDNN_FUSED=0) on the same context: last-hidden bytes, logits and softmax
    scores identical, for ragged batch sizes, several tiles per CTA in the output layer, and dense saturation"""
    dnn = qd.QuantizedDnn.load_from_file(net_file(shape, stress=stress))
    dnn.set_tile_policy(policy)
    dim = dnn.input_dimension()
    sizes = (1, 100, 128, 129, 512, 1000) if shape != "S" else (1, 100, 129, 512, 1000, 2500, 4096)
    try:
        for n in sizes:
            frames = synth.make_frames(n, dim, seed=60 + n)
            outs = []
            for flag in ("2", "0"):  # 2 = fused wherever it is applicable, also where the default plan prefers layer by layer
                os.environ["FDNN_FUSED"] = flag
                ctx = dnn.get_new_lazy_context(n)
                ctx.calculate_until_output(frames)
                outs.append((ctx.hidden().copy(), ctx.logits().copy(), dnn.calculate(frames).copy()))
                ctx.delete()
            assert np.array_equal(outs[0][0], outs[1][0]), f"{n} frames: last-hidden bytes differ"
            assert np.array_equal(outs[0][1].view(np.uint32), outs[1][1].view(np.uint32)), f"{n} frames: logits differ"
            assert np.array_equal(outs[0][2].view(np.uint32), outs[1][2].view(np.uint32)), f"{n} frames: softmax scores differ"
    finally:
        os.environ.pop("FDNN_FUSED", None)
        dnn.delete()


def test_softmax_overflow_and_switched_off_class(tmp_path):
    """SoftMax::apply has no max subtraction (dnn.cc:534-544): a logit above 88.7 overflows expf to +inf and the row becomes
    zeros and one NaN; a −inf bias gives exp = 0 and a valid distribution.  Same here."""
    from fast_dnn_b200 import formats
    layers, shift, scale = synth.make_network("tiny")
    w, b = layers[-1]
    b = b.copy()
    b[3] = -np.inf
    layers[-1] = (w, b)
    path = str(tmp_path / "minus_inf.bin")
    formats.write_dnn_bin(path, layers, shift, scale)
    frames = synth.make_frames(9, 12, seed=2)
    dnn, port = qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path)
    got, want = dnn.calculate(frames), port.calculate(frames)
    dnn.delete()
    assert np.all(got[:, 3] == 0) and np.all(want[:, 3] == 0) and np.isfinite(got).all()
    softmax_close(got, want)
    b[3] = 200.0
    layers[-1] = (w, b)
    path = str(tmp_path / "overflow.bin")
    formats.write_dnn_bin(path, layers, shift, scale)
    dnn, port = qd.QuantizedDnn.load_from_file(path), oracle_py.Port(path)
    got, want = dnn.calculate(frames), port.calculate(frames)
    dnn.delete()
    assert np.isnan(want[:, 3]).all() and np.array_equal(np.isnan(got), np.isnan(want))
    assert np.all(got[~np.isnan(got)] == 0) and np.all(want[~np.isnan(want)] == 0)


def test_variable_length_calls_share_pooled_workspaces(net_file):
    """utterances of many different lengths through calculate(): bucketed workspaces, launch sequences captured on the second
    sighting of a length — every call equals the rows of one long call"""
    dnn = qd.QuantizedDnn.load_from_file(net_file("S"))
    try:
        frames = synth.make_frames(1500, 440, seed=31)
        full = dnn.calculate(frames)
        rng = np.random.default_rng(5)
        for n in [1, 2, 127, 128, 129, 200, 200, 200, 333, 1024, 1025, 1500] + [int(v) for v in rng.integers(1, 1500, 20)]:
            lo = int(rng.integers(0, 1500 - n + 1))
            assert np.array_equal(dnn.calculate(frames[lo:lo + n]), full[lo:lo + n]), n
    finally:
        dnn.delete()


def test_context_may_outlive_its_model(net_file):
    """the reference's `delete context` never touches the dnn (jni_dnn.cc:119-133): Java code that deletes the QuantizedDnn
    before its LazyContexts must keep working"""
    dnn = qd.QuantizedDnn.load_from_file(net_file("tiny"))
    frames = synth.make_frames(8, 12, seed=3)
    ctx = dnn.get_new_lazy_context(8)
    ctx.calculate_until_output(frames)
    want = ctx.hidden().copy()
    last, H = dnn.layer_count() - 2, dnn.hidden_dimension()
    dnn.delete()
    got = np.empty((8, H), dtype=np.uint8)  # raw C ABI: the Python mirror asks the (deleted) model for the shapes
    qd._check(qd.lib().fdnn_ctx_hidden(ctx._h, last, 8, got.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(got, want)  # the context keeps its replica alive
    row = np.empty(20, dtype=np.float32)
    mask = np.ones(20, dtype=np.int8)
    qd._check(qd.lib().fdnn_ctx_lazy(ctx._h, 0, mask.ctypes.data_as(C.c_void_p), row.ctypes.data_as(C.c_void_p)))
    assert abs(float(row.sum()) - 1.0) < 1e-5
    ctx.delete()


def test_jni_calculate_long_input_through_the_sink(net_file):
    """Java_suskun_nn_QuantizedDnn_calculate on a pageable float[] longer than one streaming chunk: the scores arrive through
    the staged path (page-locked ring → SetFloatArrayRegion per 128-frame piece) in order, equal to the direct call"""
    jvm = jni_fake.FakeJvm()
    fn = jni_fake.bind(C.CDLL(qd.LIB_PATH))
    path = net_file("S")
    h = fn["initialize"](jvm.env, None, jvm.new_string(path), 3.0)
    assert h != 0 and not jvm.thrown
    n = 4096 + 4096 + 777
    frames = synth.make_frames(n, 440, seed=41)
    j_out = fn["calculate"](jvm.env, None, h, jvm.new_array(frames.reshape(-1).copy()), n, 440, 10)
    assert j_out and not jvm.thrown
    got = jvm.objects[j_out].reshape(n, 2000)
    direct = qd.QuantizedDnn.load_from_file(path)
    pinned_in, pinned_out = qd.PinnedArray((n, 440), np.float32), qd.PinnedArray((n, 2000), np.float32)
    pinned_in.array[:] = frames
    want = direct.calculate(pinned_in.array, out=pinned_out.array)  # page-locked caller memory: the direct path
    assert np.array_equal(got, want)
    assert np.array_equal(direct.calculate(frames), want)           # pageable caller memory through the C ABI
    direct.delete()
    fn["delete"](jvm.env, None, h)


# ---- multi-GPU parity (SURVEY.md §4 v: sharded results equal the 1-GPU results bit for bit) -------------------------------

def _free_port():
    """a port below the kernel's ephemeral range (an ephemeral one can be handed to somebody's outgoing connection between this
    probe and the rendezvous binding it — seen once on a GPU box: EADDRINUSE)"""
    import random
    for _ in range(200):
        port = random.randint(15000, 29999)
        with socket.socket() as s:
            try:
                s.bind(("127.0.0.1", port))
                return port
            except OSError:
                continue
    raise RuntimeError("no free port found")


RANK_SCRIPT = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r})
import fast_dnn_b200
from fast_dnn_b200 import quantized_dnn as qd, sharding, synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
n_gpus = torch.cuda.device_count()
local = rank % n_gpus
torch.cuda.set_device(local)
nccl = n_gpus >= world                      # one GPU per rank: NCCL; ranks sharing a GPU: gloo carries the broadcast
dist.init_process_group("nccl" if nccl else "gloo")
dev = torch.device("cuda", local) if nccl else torch.device("cpu")
blob = sharding.broadcast_blob(qd.pack({path!r}) if rank == 0 else None, src=0, device=dev)
dnn = (qd.QuantizedDnn.load_from_blob(blob.data_ptr(), device=local, size=blob.numel()) if nccl
       else qd.QuantizedDnn.load_from_blob(blob.numpy(), device=local))
n = {n}
lo, hi = sharding.shard_range(n, rank, world)
parts = [dnn.calculate(synth.make_frames(b - a, 440, seed=7, start=a)) for a, b in sharding.chunk_ranges(lo, hi, 1000)]
np.save(os.path.join({out!r}, f"rank{{rank}}.npy"), np.concatenate(parts))
counts = sharding.gather_counts(hi - lo, device=dev)
assert sum(counts) == n
dnn.delete()
dist.destroy_process_group()
"""


def test_two_ranks_sharded_output_equals_single_gpu(tmp_path, net_file):
    """torchrun, 2 ranks (one GPU each when the box has two, otherwise both on GPU 0 with the blob broadcast over gloo): rank 0
    packs, one broadcast of the blob, every rank computes its contiguous shard of a 5003-frame stream; concatenated shards ==
    the single-GPU result, byte for byte"""
    path, n = net_file("S"), 5003
    script = tmp_path / "rank.py"
    script.write_text(RANK_SCRIPT.format(root=ROOT, path=path, n=n, out=str(tmp_path)))
    for attempt in range(3):  # a rendezvous port can still be lost to a race: try another one
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", str(_free_port()), str(script)], timeout=600, capture_output=True, text=True)
        if r.returncode == 0 or "EADDRINUSE" not in r.stderr:
            break
    assert r.returncode == 0, r.stderr[-3000:]
    shards = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(2)])
    dnn = qd.QuantizedDnn.load_from_file(path)
    want = dnn.calculate(synth.make_frames(n, 440, seed=7))
    dnn.delete()
    assert shards.shape == want.shape and np.array_equal(shards.view(np.uint32), want.view(np.uint32))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs on the box")
def test_device_group_behind_one_handle(net_file):
    """fdnn_load_devices: one handle over every GPU of the box — one ncclBroadcast at load, calculate() shards its frames,
    lazy contexts go round-robin; all bytes equal the single-GPU results.  Also through the JNI entry point with FDNN_DEVICES."""
    path = net_file("S")
    devs = list(range(_device_count()))
    single = qd.QuantizedDnn.load_from_file(path, device=0)
    before = qd.lib().fdnn_nccl_broadcast_count()
    group = qd.QuantizedDnn.load_on_devices(path, devs)
    assert qd.lib().fdnn_nccl_broadcast_count() == before + 1, "exactly one broadcast of the weight blob"
    assert group.devices() == devs
    try:
        for n in (1, 100, 129, 512, 3000, 9001):
            frames = synth.make_frames(n, 440, seed=50 + n)
            assert np.array_equal(group.calculate(frames), single.calculate(frames)), n
        # the file-to-file front end on the group handle: every chunk of the feature file is sharded over the devices
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            frames = synth.make_frames(7000, 440, seed=77)
            feats, dump = os.path.join(d, "f.bin"), os.path.join(d, "o.bin")
            qd.write_feature_bin(feats, frames)
            assert group.calculate_file(feats, dump, chunk_frames=1500) == 7000
            assert np.array_equal(qd.read_output_dump(dump), single.calculate(frames))
        frames = synth.make_frames(64, 440, seed=9)
        masks = synth.make_masks(64, 2000, seed=4)
        want_ctx = single.get_new_lazy_context(64)
        want_ctx.calculate_until_output(frames)
        want = want_ctx.calculate_for_output_nodes_batch(masks)
        want_ctx.delete()
        for _ in range(len(devs) + 1):  # round-robin over the devices and back to the first
            ctx = group.get_new_lazy_context(64)
            ctx.calculate_until_output(frames)
            assert np.array_equal(ctx.calculate_for_output_nodes_batch(masks), want)
            ctx.delete()
    finally:
        group.delete()
        single.delete()
    code = (
        "import sys, ctypes as C, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); import fast_dnn_b200\n"
        "from fast_dnn_b200 import quantized_dnn as qd, synth; import jni_fake\n"
        "jvm = jni_fake.FakeJvm(); fn = jni_fake.bind(C.CDLL(qd.LIB_PATH))\n"
        "h = fn['initialize'](jvm.env, None, jvm.new_string(%r), 3.0); assert h and not jvm.thrown, jvm.thrown\n"
        "assert qd.lib().fdnn_device_count(C.c_void_p(h)) == %d\n"
        "x = synth.make_frames(2000, 440, seed=8)\n"
        "out = fn['calculate'](jvm.env, None, h, jvm.new_array(x.reshape(-1).copy()), 2000, 440, 10)\n"
        "np.save(sys.argv[1], jvm.objects[out].reshape(2000, 2000)); fn['delete'](jvm.env, None, h)\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), path, len(devs))
    out = "/tmp/fdnn_group_jni.npy"
    subprocess.run([sys.executable, "-c", code, out], check=True, env=dict(os.environ, FDNN_DEVICES="all"), timeout=300)
    single = qd.QuantizedDnn.load_from_file(path, device=0)
    assert np.array_equal(np.load(out), single.calculate(synth.make_frames(2000, 440, seed=8)))
    single.delete()
