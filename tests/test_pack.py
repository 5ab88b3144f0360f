"""Host half of model loading (csrc/model_host.cc) through the C ABI's host-only entry points:
dnn.bin parsing, int8 quantization with the reference's quirks, LUT, saturation risk lists,
blob validation and error codes.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from fast_dnn_b200 import blob as B
from fast_dnn_b200 import formats, synth
from fast_dnn_b200 import quantized_dnn as qd
import oracle_py


@pytest.mark.parametrize("shape,stress", [("tiny", False), ("tiny", True), ("ragged", False), ("S", False), ("S", True), ("P", False)])
def test_blob_matches_oracle(shape, stress, net_file):
    path = net_file(shape, stress=stress)
    b = B.Blob(qd.pack(path))
    port = oracle_py.Port(path)
    assert (b.in_dim, b.hidden, b.out_dim, len(b.qlayers)) == (port.input_dim, port.hidden_dim, port.output_dim, port.qlayer_count)
    for got, want in zip(b.input_layer(), port.input_layer()):
        assert np.array_equal(got, want)
    # doubled LUT: slot v + 1282 with v = trunc(2c) holds QuantizedSigmoid::get for every c that truncates to v
    lut2 = b.lut2()
    ks = (np.abs(np.arange(-1282, 1283)) + 1) // 2 * np.sign(np.arange(-1282, 1283))
    lut = oracle_py.Port.sigmoid_lut()
    want = np.where(ks <= -640, 0, np.where(ks >= 640, 255, lut[np.clip(ks + 640, 0, 1279)]))
    assert np.array_equal(lut2, want.astype(np.uint8))
    for i in range(port.qlayer_count):
        w, bias, mult = b.qlayer(i)
        w2, b2, m2 = port.qlayer(i)
        assert np.array_equal(w, w2) and np.array_equal(bias, b2) and mult == m2
        q = b.qlayers[i]
        assert q["coeff"] == np.float32(mult) * np.float32(255.0)
        assert q["rcp_coeff"] == np.float32(1.0) / q["coeff"]
        # risk list = exactly the pairs whose same-sign magnitudes reach 129
        wi = w.astype(np.int32)
        pos = np.maximum(wi[:, 0::2], 0) + np.maximum(wi[:, 1::2], 0)
        neg = np.minimum(wi[:, 0::2], 0) + np.minimum(wi[:, 1::2], 0)
        nodes, pairs = np.nonzero((pos >= 129) | (neg <= -129))
        want = set(zip(pairs.tolist(), nodes.tolist()))
        assert int(q["n_fix"]) == len(want)
        kb = int(q["k_blocks"])
        assert kb == -(-w.shape[1] // 128)
        for variant, G in enumerate(B.FIX_GROUPS):
            # one copy per tile width G, ordered by (node // G, K block of the pair, node, pair)
            ptr, pair, w0, w1, node = b.fix_list(i, variant)
            assert len(pair) == len(want) and set(zip(pair.tolist(), node.tolist())) == want
            assert np.array_equal(w0, w[node, 2 * pair]) and np.array_equal(w1, w[node, 2 * pair + 1])
            ng = -(-w.shape[0] // G)
            key = (node.astype(np.int64) // G) * kb + (2 * pair.astype(np.int64)) // 128
            assert np.all(np.diff(key) >= 0)
            assert ptr[0] == 0 and ptr[-1] == len(pair) and len(ptr) == ng * kb + 1
            assert np.array_equal(np.searchsorted(key, np.arange(ng * kb + 1)), ptr)


def test_stress_network_hits_quantizer_quirks(net_file):
    """weights above +cutoff are not clipped and wrap modulo 256 (dnn.cc:493-499)"""
    layers, _, _ = synth.make_network("S", stress=True)
    w = layers[1][0]
    b = B.Blob(qd.pack(net_file("S", stress=True)))
    q, _, mult = b.qlayer(0)
    over = w > 3.0
    assert over.sum() > 0
    raw = np.round(w[over].astype(np.float32) * np.float32(mult)).astype(np.int64)
    assert raw.max() > 127  # wrap happens
    assert np.array_equal(q[over], (raw & 0xFF).astype(np.uint8).view(np.int8))
    under = w < -3.0
    assert np.all(q[under] == np.int8(np.round(np.float32(-3.0) * np.float32(mult))))


def test_fast_division_is_verified_for_every_layer(net_file):
    b = B.Blob(qd.pack(net_file("S")))
    assert all(int(q["fast_div"]) == 1 for q in b.qlayers)
    # spot-check the claim independently in float32/float64 arithmetic
    rng = np.random.default_rng(0)
    for q in b.qlayers:
        c, r = np.float32(q["coeff"]), np.float32(q["rcp_coeff"])
        s = rng.integers(-256 * 32768, 256 * 32768, size=200000).astype(np.float32)
        qq = (s * r).astype(np.float32)
        e = (s.astype(np.float64) - qq.astype(np.float64) * np.float64(c)).astype(np.float32)  # exact in double → fma
        q2 = (qq.astype(np.float64) + e.astype(np.float64) * np.float64(r)).astype(np.float32)
        assert np.array_equal(q2, s / c)


def test_unaligned_input_is_padded_to_multiple_of_four(tmp_path):
    layers, shift, scale = synth.make_network((10, 32, 3, 7), seed=3)
    path = str(tmp_path / "unaligned.dnn.bin")
    formats.write_dnn_bin(path, layers, shift, scale)
    b = B.Blob(qd.pack(path))
    assert b.in_dim == 12 and int(b.header["in_dim_file"]) == 10
    w0, _, sh, sc = b.input_layer()
    assert np.array_equal(w0[:, :10], layers[0][0]) and np.all(w0[:, 10:] == 0) and np.all(sh[10:] == 0) and np.all(sc[10:] == 0)
    port = oracle_py.Port(path)
    assert port.input_dim == 12


def _pack_rc(path, cutoff=3.0):
    blob, size = C.c_void_p(), C.c_size_t()
    rc = qd.lib().fdnn_pack(os.fsencode(path), cutoff, C.byref(blob), C.byref(size))
    if rc == 0:
        qd.lib().fdnn_blob_free(blob)
    return rc, qd.lib().fdnn_last_error().decode()


def test_error_codes(tmp_path, net_file):
    assert _pack_rc(str(tmp_path / "missing.bin"))[0] == qd.FDNN_EIO  # the reference dereferences a null FILE* here
    assert _pack_rc(net_file("tiny"), cutoff=0.0)[0] == qd.FDNN_EINVAL
    assert _pack_rc(net_file("tiny"), cutoff=-1.0)[0] == qd.FDNN_EINVAL
    good = open(net_file("tiny"), "rb").read()
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(good[: len(good) // 2])
    assert _pack_rc(str(trunc))[0] == qd.FDNN_EIO
    trunc.write_bytes(good[:-4])
    assert _pack_rc(str(trunc))[0] == qd.FDNN_EIO
    # two layers only: the reference needs two int8 layers (dnn.cc:199)
    layers, shift, scale = synth.make_network((12, 32, 1, 20))
    two = str(tmp_path / "two.bin")
    formats.write_dnn_bin(two, layers, shift, scale)
    assert _pack_rc(two)[0] == qd.FDNN_EFORMAT
    # hidden width not a multiple of 16 (dnn.cc:331 loads 16 bytes at a time)
    layers, shift, scale = synth.make_network((12, 24, 3, 20))
    bad = str(tmp_path / "h24.bin")
    formats.write_dnn_bin(bad, layers, shift, scale)
    rc, msg = _pack_rc(bad)
    assert rc == qd.FDNN_EFORMAT and "16" in msg
    # unequal hidden widths
    a = synth.make_network((12, 32, 2, 20))[0]
    c = synth.make_network((32, 48, 2, 20))[0]
    mixed = [a[0], c[0], (np.zeros((20, 48), np.float32), np.zeros(20, np.float32))]
    bad2 = str(tmp_path / "mixed.bin")
    formats.write_dnn_bin(bad2, mixed, shift, scale)
    assert _pack_rc(bad2)[0] == qd.FDNN_EFORMAT
    garbage = tmp_path / "garbage.bin"
    garbage.write_bytes(b"\xff" * 64)
    assert _pack_rc(str(garbage))[0] in (qd.FDNN_EFORMAT, qd.FDNN_EIO)


def test_blob_validation_rejects_corruption(net_file):
    blob = qd.pack(net_file("tiny"))
    h = C.c_void_p()
    lib = qd.lib()
    # without a GPU the device check comes first; validation is exercised through a corrupted header either way
    bad = blob.copy()
    bad[0] ^= 0xFF
    rc = lib.fdnn_load_blob(bad.ctypes.data_as(C.c_void_p), bad.nbytes, -1, C.byref(h))
    assert rc in (qd.FDNN_EFORMAT, qd.FDNN_ENOGPU)
    rc = lib.fdnn_load_blob(blob.ctypes.data_as(C.c_void_p), blob.nbytes - 256, -1, C.byref(h))
    assert rc in (qd.FDNN_EFORMAT, qd.FDNN_ENOGPU)


def test_doubled_lut_trick_equals_reference_rounding(net_file):
    """numpy float32 emulation of device qsig_slot() (csrc/device_common.cuh): one clamp, one exact
    doubling, one truncation, one lookup == QuantizedSigmoid::get (round half away, two clamps)"""
    lut2 = B.Blob(qd.pack(net_file("tiny"))).lut2()
    rng = np.random.default_rng(1)
    xs = np.concatenate([
        rng.normal(0, 3, 20000).astype(np.float32),
        (np.arange(-1300, 1301, dtype=np.float32) / np.float32(200.0)),           # every half-step of k
        np.nextafter(np.arange(-1300, 1301, dtype=np.float32) / np.float32(200.0), np.float32(np.inf)),
        np.nextafter(np.arange(-1300, 1301, dtype=np.float32) / np.float32(200.0), np.float32(-np.inf)),
        np.float32([0.004999999888241291, 0.005, 0.0049999995, -0.005, 6.4, -6.4, 6.395, 6.3949995, 1e9, -1e9, 2.2e7, 2.1474836e7,
                    np.inf, -np.inf, np.nan, 0.0, -0.0, 1e-40]),
    ])
    with np.errstate(invalid="ignore", over="ignore"):
        t = (xs * np.float32(100.0)).astype(np.float32)
        c = np.minimum(np.maximum(t, np.float32(-641.0)), np.float32(641.0))
        c = np.where(np.isnan(t), np.float32(-641.0), c)
        c = np.where(~(t < np.float32(2147483648.0)), np.float32(-641.0), c)
        v = np.trunc((c + c).astype(np.float32)).astype(np.int64)
    got = lut2[v + 1282]
    want = np.array([oracle_py.Port.qsigmoid(float(x)) for x in xs], dtype=np.uint8)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]


def test_parser_survives_mutated_files(tmp_path, net_file):
    """seeded mutations of a valid dnn.bin — header bytes, random bytes, truncations, extreme big-endian words (layer counts and
    dimensions of 2³¹) — must come back as an error code or a packed model, never a crash or an allocation the file cannot back
    (the reference's loader trusts every field, float_dnn.cc:18-69)"""
    good = open(net_file("tiny"), "rb").read()
    rng = np.random.default_rng(5)
    case = tmp_path / "case.bin"
    seen = set()
    for it in range(300):
        b = bytearray(good)
        kind = it % 4
        if kind == 0:
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, 20))] = int(rng.integers(0, 256))
        elif kind == 1:
            for _ in range(int(rng.integers(1, 8))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif kind == 2:
            b = b[: int(rng.integers(0, len(b)))]
        else:
            pos = int(rng.integers(0, len(b) // 4)) * 4
            b[pos:pos + 4] = int(rng.choice([0x7FFFFFFF, 0xFFFFFFFF, 0x80000000, 0, 1 << 24])).to_bytes(4, "big")
        case.write_bytes(bytes(b))
        rc, _ = _pack_rc(str(case))
        assert rc in (0, qd.FDNN_EIO, qd.FDNN_EFORMAT, qd.FDNN_EINVAL, qd.FDNN_ENOMEM), rc
        seen.add(rc)
    assert 0 in seen and len(seen) >= 2
