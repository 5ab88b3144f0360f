"""The file-to-file front end (csrc/stream_file.cc) on the GPU: a big-endian feature file in, the
reference's binary / text dump out — against the reference's own command-line driver (dnn.cc:20-83, run from
oracle/_ref) and against calculate() on the same frames, with ragged chunks, an unpadded feature width, and broken files."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import softmax_close
from fast_dnn_b200 import formats, synth
from fast_dnn_b200 import quantized_dnn as qd
import oracle_py

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chunk", [0, 384, 5000])
def test_file_to_file_equals_calculate(tmp_path, net_file, chunk):
    """BIN dump == the scores calculate() returns for the same frames, whatever the chunking (frames are independent and every
    kernel variant gives the same bits); the TXT dump is the text form of the same numbers."""
    dnn = qd.QuantizedDnn.load_from_file(net_file("S"))
    try:
        frames = synth.make_frames(3000, 440, seed=17)
        feats, out_bin, out_txt, want_txt = (str(tmp_path / n) for n in ("f.bin", "o.bin", "o.txt", "want.txt"))
        qd.write_feature_bin(feats, frames)
        want = dnn.calculate(frames)
        assert dnn.calculate_file(feats, out_bin, binary=True, chunk_frames=chunk) == 3000
        got = qd.read_output_dump(out_bin)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        if chunk == 384:
            assert dnn.calculate_file(feats, out_txt, binary=False, chunk_frames=chunk) == 3000
            qd.write_output_dump(want_txt, want, binary=False)
            assert open(out_txt, "rb").read() == open(want_txt, "rb").read()
    finally:
        dnn.delete()


@pytest.mark.skipif(not oracle_py.have_ref(), reason="needs the compiled reference (oracle/_ref)")
def test_file_to_file_against_the_reference_cli(tmp_path, net_file):
    """same model file, same feature file: our dump against the one the reference's command-line driver writes"""
    model = net_file("P")
    frames = synth.make_frames(700, 432, seed=23)
    feats, ref_bin, our_bin = (str(tmp_path / n) for n in ("f.bin", "ref.bin", "our.bin"))
    qd.write_feature_bin(feats, frames)
    oracle_py.Ref.cli(model, feats, ref_bin, binary=True)
    dnn = qd.QuantizedDnn.load_from_file(model)
    try:
        assert dnn.calculate_file(feats, our_bin, chunk_frames=256) == 700
    finally:
        dnn.delete()
    assert open(our_bin, "rb").read(8) == open(ref_bin, "rb").read(8)  # native-endian (700, 2000)
    got, want = qd.read_output_dump(our_bin), qd.read_output_dump(ref_bin)
    softmax_close(got, want)
    assert np.array_equal(got.argmax(axis=1), want.argmax(axis=1))


def test_unpadded_feature_width_and_broken_files(tmp_path):
    """a 430-input network is padded to 432 at load (float_dnn.cc:32-33); a feature file may carry 430 or 432 columns"""
    layers, shift, scale = synth.make_network((430, 256, 3, 500), seed=9)
    model = str(tmp_path / "n430.dnn.bin")
    formats.write_dnn_bin(model, layers, shift, scale)
    dnn = qd.QuantizedDnn.load_from_file(model)
    try:
        assert dnn.input_dimension() == 432
        x = synth.make_frames(300, 430, seed=4)
        padded = np.zeros((300, 432), dtype=np.float32)
        padded[:, :430] = x
        want = dnn.calculate(padded)
        f430, f432, out = (str(tmp_path / n) for n in ("f430.bin", "f432.bin", "o.bin"))
        qd.write_feature_bin(f430, x)
        qd.write_feature_bin(f432, padded)
        for f in (f430, f432):
            assert dnn.calculate_file(f, out, chunk_frames=128) == 300
            assert np.array_equal(qd.read_output_dump(out).view(np.uint32), want.view(np.uint32))
        # wrong width, truncated body, missing file, empty matrix
        L, done = qd.lib(), C.c_longlong()
        bad = str(tmp_path / "bad.bin")
        qd.write_feature_bin(bad, synth.make_frames(8, 428, seed=1))
        assert L.fdnn_calculate_file(dnn._h, bad.encode(), out.encode(), qd.FDNN_DUMP_BIN, 0, C.byref(done)) == qd.FDNN_EINVAL
        data = open(f432, "rb").read()
        open(bad, "wb").write(data[: len(data) // 2])
        assert L.fdnn_calculate_file(dnn._h, bad.encode(), out.encode(), qd.FDNN_DUMP_BIN, 0, C.byref(done)) == qd.FDNN_EIO
        assert L.fdnn_calculate_file(dnn._h, b"/nonexistent/f.bin", out.encode(), qd.FDNN_DUMP_BIN, 0, C.byref(done)) == qd.FDNN_EIO
        assert L.fdnn_calculate_file(dnn._h, f432.encode(), b"/nonexistent-dir/o.bin", qd.FDNN_DUMP_BIN, 0, C.byref(done)) == qd.FDNN_EIO
        qd.write_feature_bin(bad, np.zeros((0, 432), dtype=np.float32))
        assert dnn.calculate_file(bad, out) == 0 and qd.read_output_dump(out).shape == (0, 500)
    finally:
        dnn.delete()
