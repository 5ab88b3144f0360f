"""Formats and offline tooling either side of the hot path: the C++
aligner/writer against the numpy mirror of FeedForwardNetwork.align, feature-file IO against the
shipped fixtures' layout, the output dump, and the quantization-quality criterion of FuncTest.diff."""
import os

import numpy as np
import pytest

from conftest import REFERENCE_ROOT
from fast_dnn_b200 import formats, synth
from fast_dnn_b200 import quantized_dnn as qd
import oracle_py


def test_cpp_aligner_equals_java_semantics(tmp_path):
    layers, shift, scale = synth.make_network((10, 40, 3, 7), seed=5)  # input 10 → 12, hidden 40 → 48, output stays 7
    raw, aligned_cpp, aligned_np = (str(tmp_path / n) for n in ("raw.bin", "cpp.bin", "np.bin"))
    formats.write_dnn_bin(raw, layers, shift, scale)
    with pytest.raises(qd.FdnnError):
        qd.pack(raw)  # hidden width 40 is not a multiple of 16: the loader refuses it, like the reference would misbehave
    qd.align_dnn_bin(raw, aligned_cpp, 4, 16)
    formats.write_dnn_bin(aligned_np, *formats.align_network(layers, shift, scale, 4, 16))
    assert open(aligned_cpp, "rb").read() == open(aligned_np, "rb").read()
    got, sh, sc = formats.read_dnn_bin(aligned_cpp)
    assert [w.shape for w, _ in got] == [(48, 12), (48, 48), (48, 48), (7, 48)]
    assert np.all(got[0][0][40:] == 0) and np.all(got[1][0][:, 40:] == 0) and np.all(got[0][1][40:] == 0) and sh.shape == (12,)
    # padding does not change any real output: the oracle on the aligned file vs a float64 forward of the unaligned net
    port = oracle_py.Port(aligned_cpp)
    frames = synth.make_frames(16, 12, seed=3)
    frames[:, 10:] = 0
    q = port.calculate(frames)
    assert q.shape == (16, 7) and np.allclose(q.sum(axis=1), 1.0, atol=1e-5)
    qd.pack(aligned_cpp)  # and our loader now takes it
    assert qd.lib().fdnn_align_dnn_bin(b"/nonexistent/x", aligned_cpp.encode(), 4, 16) == qd.FDNN_EIO


def test_feature_bin_round_trip_and_dump(tmp_path):
    x = synth.make_frames(37, 24, seed=1)
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    qd.write_feature_bin(a, x)
    formats.write_feature_bin(b, x)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert np.array_equal(qd.read_feature_bin(a), x) and np.array_equal(formats.read_feature_bin(a), x)
    dump = str(tmp_path / "out.bin")
    qd.write_output_dump(dump, x)
    assert np.array_equal(formats.read_output_dump(dump), x)  # native-endian, unlike the inputs (float_dnn.cc:128-164)
    raw = np.fromfile(dump, dtype="<u4", count=2)
    assert tuple(raw) == (37, 24)


def test_text_dump_format():
    """`ostream << float` (float_dnn.cc:149): %g with six significant digits, one blank between values, a line per row"""
    rows = np.array([[1.0, 0.5, 0.000125, 1e-5, 1.234567e-7], [123456.7, 0.1, 2.5e10, 0.0, -0.0]], dtype=np.float32)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "out.txt")
        qd.write_output_dump(path, rows, binary=False)
        text = open(path).read()
    assert text == "1 0.5 0.000125 1e-05 1.23457e-07\n123457 0.1 2.5e+10 0 -0\n"


@pytest.mark.skipif(not oracle_py.have_ref(), reason="needs the compiled reference (oracle/_ref)")
def test_dump_writers_equal_the_reference_cli(tmp_path, net_file):
    """The reference's own command-line driver (dnn.cc:20-83, `main` compiled as ref_main) writes a BIN and a TXT dump of
    a small feature file; our writers (the ones the file-to-file front end csrc/stream_file.cc uses) given the same
    scores produce the same bytes; the feature file it read was written by our writer."""
    frames = synth.make_frames(37, 432, seed=21)
    feats, ref_bin, ref_txt, our_bin, our_txt = (str(tmp_path / n) for n in ("f.bin", "ref.bin", "ref.txt", "our.bin", "our.txt"))
    qd.write_feature_bin(feats, frames)
    model = net_file("P")
    oracle_py.Ref.cli(model, feats, ref_bin, binary=True)
    oracle_py.Ref.cli(model, feats, ref_txt, binary=False)
    scores = qd.read_output_dump(ref_bin)
    assert scores.shape == (37, 2000) and np.allclose(scores.sum(axis=1), 1.0, atol=1e-4)
    # the CLI computes with batch 8 (dnn.cc:66); the result does not depend on the batch size
    ref = oracle_py.Ref(model)
    assert np.array_equal(scores, ref.calculate(frames, batch=10))
    qd.write_output_dump(our_bin, scores, binary=True)
    qd.write_output_dump(our_txt, scores, binary=False)
    assert open(our_bin, "rb").read() == open(ref_bin, "rb").read()
    assert open(our_txt, "rb").read() == open(ref_txt, "rb").read()


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="shipped fixtures live in /root/reference")
def test_reads_shipped_feature_files():
    x = qd.read_feature_bin(os.path.join(REFERENCE_ROOT, "data", "8khz.aligned.bin"))
    assert x.shape == (389, 432) and np.all(x[:, 429:] == 0)
    y = qd.read_feature_bin(os.path.join(REFERENCE_ROOT, "data", "16khz.bin"))
    assert y.shape == (100, 432)  # header says 100; the 101st row in the file is ignored (float_dnn.cc:88-102)
    assert np.array_equal(x, formats.read_feature_bin(os.path.join(REFERENCE_ROOT, "data", "8khz.aligned.bin")))


def naive_forward(layers, shift, scale, frames):
    """FeedForwardNetwork.calculate (FeedForwardNetwork.java:133-148, 360-414) in float64"""
    a = (frames.astype(np.float64) + shift) * scale
    for i, (w, b) in enumerate(layers):
        z = a @ w.astype(np.float64).T + b
        if i < len(layers) - 1:
            a = 1.0 / (1.0 + np.exp(-z))
        else:
            e = np.exp(z)
            a = e / e.sum(axis=1, keepdims=True)
    return a


@pytest.mark.parametrize("shape", ["S", "P"])
def test_quantization_quality_criterion(shape):
    """FuncTest.diff (FuncTest.java:59-74): per output node, Σ over frames |quantized − naive| ≤ 0.1"""
    layers, shift, scale = synth.make_network(shape)
    frames = synth.make_frames(100, layers[0][0].shape[1], seed=7)
    q = oracle_py.Port(synth.network_file(shape)).calculate(frames)
    f = naive_forward(layers, shift, scale, frames)
    assert np.abs(q - f).sum(axis=0).max() < 0.1
    assert np.array_equal(q.argmax(axis=1), f.argmax(axis=1))


@pytest.mark.gpu
def test_gpu_path_meets_the_quality_criterion_too():
    layers, shift, scale = synth.make_network("S")
    frames = synth.make_frames(100, 440, seed=7)
    dnn = qd.QuantizedDnn.load_from_file(synth.network_file("S"))
    q = dnn.calculate(frames)
    dnn.delete()
    assert np.abs(q - naive_forward(layers, shift, scale, frames)).sum(axis=0).max() < 0.1


def _kaldi_text(layers, with_splice=True):
    """nnet1 text the way Kaldi's nnet-copy --binary=false prints it, and the matching feature-transform text"""
    out = ["<Nnet> "]
    for j, (w, b) in enumerate(layers):
        out.append(f"<AffineTransform> {w.shape[0]} {w.shape[1]} ")
        out.append("<LearnRateCoef> 1 <BiasLearnRateCoef> 1 <MaxNorm> 0  [")
        for r, row in enumerate(w):
            out.append("  " + " ".join(repr(float(v)) for v in row) + (" ]" if r == w.shape[0] - 1 else " "))
        out.append(" [ " + " ".join(repr(float(v)) for v in b) + " ]")
        out.append(f"<Sigmoid> {w.shape[0]} {w.shape[0]} " if j + 1 < len(layers) else f"<Softmax> {w.shape[0]} {w.shape[0]} ")
    out.append("</Nnet> ")
    return "\n".join(out) + "\n"


def _transform_text(shift, scale, with_splice):
    out = ["<Nnet> "]
    if with_splice:
        out += [f"<Splice> {len(shift)} {len(shift) // 11} ", "[ -5 -4 -3 -2 -1 0 1 2 3 4 5 ]"]
    out += [f"<AddShift> {len(shift)} {len(shift)} ", "<LearnRateCoef> 0  [ " + " ".join(repr(float(v)) for v in shift) + " ]"]
    out += [f"<Rescale> {len(scale)} {len(scale)} ", "<LearnRateCoef> 0  [ " + " ".join(repr(float(v)) for v in scale[: len(scale) // 2])]
    out += ["  " + " ".join(repr(float(v)) for v in scale[len(scale) // 2:]) + " ]", "</Nnet> "]
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("with_splice", [True, False])
def test_kaldi_nnet1_text_importer(tmp_path, with_splice):
    """SURVEY.md §8f row 3 (FeedForwardNetwork.loadFromTextFile, FeedForwardNetwork.java:86-119,159-207): text model +
    feature transform → the same dnn.bin bytes as the binary writer produces from the same numbers"""
    layers, shift, scale = synth.make_network((22, 40, 3, 7), seed=9)
    nnet, trans, got, want = (str(tmp_path / n) for n in ("final.nnet.txt", "final.feature_transform", "imported.bin", "direct.bin"))
    open(nnet, "w").write(_kaldi_text(layers))
    open(trans, "w").write(_transform_text(shift, scale, with_splice))
    qd.import_kaldi_nnet1(nnet, trans, got)
    formats.write_dnn_bin(want, layers, shift, scale)
    assert open(got, "rb").read() == open(want, "rb").read()
    # the imported network goes through the aligner and the packer like any other
    aligned = str(tmp_path / "aligned.bin")
    qd.align_dnn_bin(got, aligned, 4, 16)
    assert qd.pack(aligned).size > 0
    # error behaviour: wrong transform width (IllegalStateException in the reference), missing file, no layers
    open(trans, "w").write(_transform_text(shift[:-1], scale[:-1], with_splice))
    assert qd.lib().fdnn_import_kaldi_nnet1(nnet.encode(), trans.encode(), got.encode()) == qd.FDNN_EFORMAT
    assert qd.lib().fdnn_import_kaldi_nnet1(b"/nonexistent/x", trans.encode(), got.encode()) == qd.FDNN_EIO
    open(nnet, "w").write("<Nnet>\n</Nnet>\n")
    assert qd.lib().fdnn_import_kaldi_nnet1(nnet.encode(), trans.encode(), got.encode()) == qd.FDNN_EFORMAT


needs_ref = pytest.mark.skipif(not (os.path.isdir(REFERENCE_ROOT)), reason="needs the compiled reference (/root/reference)")


def _assert_same_network(got, want):
    (gl, gsh, gsc), (wl, wsh, wsc) = got, want
    assert [w.shape for w, _ in gl] == [w.shape for w, _ in wl]
    for (gw, gb), (ww, wb) in zip(gl, wl):
        assert np.array_equal(gw.view(np.uint32), ww.view(np.uint32)) and np.array_equal(gb.view(np.uint32), wb.view(np.uint32))
    assert np.array_equal(gsh.view(np.uint32), wsh.view(np.uint32)) and np.array_equal(gsc.view(np.uint32), wsc.view(np.uint32))


@needs_ref
def test_aligned_file_read_back_by_the_reference_loader(tmp_path):
    """SURVEY.md §8f row 1, pinned on the reference side: what fdnn_align_dnn_bin writes is read back by the reference's own
    FloatDnn (float_dnn.cc:18-69) with the dimensions, every weight and bias, shift and scale FeedForwardNetwork.align
    (FeedForwardNetwork.java:50-58,264-281) prescribes, and the compiled reference then computes on it what the port computes."""
    layers, shift, scale = synth.make_network((10, 40, 3, 7), seed=5)
    raw, aligned = str(tmp_path / "raw.bin"), str(tmp_path / "aligned.bin")
    formats.write_dnn_bin(raw, layers, shift, scale)
    qd.align_dnn_bin(raw, aligned, 4, 16)
    got = oracle_py.Ref.load_float_network(aligned)
    want = formats.align_network(layers, shift, scale, 4, 16)
    _assert_same_network(got, want)
    assert [w.shape for w, _ in got[0]] == [(48, 12), (48, 48), (48, 48), (7, 48)]
    # padded rows/columns and biases are exact zeros; the real block is the original
    assert np.array_equal(got[0][0][0][:40, :10], layers[0][0]) and not got[0][0][0][40:].any() and not got[0][1][0][:, 40:].any()
    ref, port = oracle_py.Ref(aligned), oracle_py.Port(aligned)
    frames = synth.make_frames(23, 12, seed=3)
    frames[:, 10:] = 0
    assert np.array_equal(ref.calculate(frames, batch=10).view(np.uint32), port.calculate(frames).view(np.uint32))
    assert np.array_equal(ref.hidden_trace(frames), port.hidden_trace(frames))
    # an already aligned network passes through unchanged (idempotence)
    again = str(tmp_path / "again.bin")
    qd.align_dnn_bin(aligned, again, 4, 16)
    assert open(again, "rb").read() == open(aligned, "rb").read()


@needs_ref
@pytest.mark.parametrize("with_splice", [True, False])
def test_kaldi_import_read_back_by_the_reference_loader(tmp_path, with_splice):
    """SURVEY.md §8f row 3, pinned on the reference side: Kaldi nnet1 text → fdnn_import_kaldi_nnet1 → fdnn_align_dnn_bin →
    the reference's FloatDnn returns the numbers that were in the text (float32(repr) is exact), and the compiled reference's
    forward on the imported file equals the port's."""
    layers, shift, scale = synth.make_network((22, 40, 3, 7), seed=9)
    nnet, trans, imported, aligned = (str(tmp_path / n) for n in ("final.nnet.txt", "final.feature_transform", "imported.bin", "aligned.bin"))
    open(nnet, "w").write(_kaldi_text(layers))
    open(trans, "w").write(_transform_text(shift, scale, with_splice))
    qd.import_kaldi_nnet1(nnet, trans, imported)
    # the unaligned import, read by the reference loader: layer-0 inputs padded to ×4 (22 → 24) by the loader itself
    got = oracle_py.Ref.load_float_network(imported)
    want = formats.align_network(layers, shift, scale, 4, 1)
    _assert_same_network(got, want)
    qd.align_dnn_bin(imported, aligned, 4, 16)
    _assert_same_network(oracle_py.Ref.load_float_network(aligned), formats.align_network(layers, shift, scale, 4, 16))
    ref, port = oracle_py.Ref(aligned), oracle_py.Port(aligned)
    frames = synth.make_frames(17, 24, seed=4)
    frames[:, 22:] = 0
    assert np.array_equal(ref.calculate(frames, batch=10).view(np.uint32), port.calculate(frames).view(np.uint32))
