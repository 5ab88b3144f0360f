"""The drop-in boundary: libfast-dnn.so loads, exports every symbol include/fdnn.h declares plus
the eleven JNI symbols of the reference (suskun_nn_QuantizedDnn.h:15-96), and — on a machine
without a B200 — refuses to compute instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, _cuda_available
from fast_dnn_b200 import quantized_dnn as qd
from fast_dnn_b200 import synth


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fdnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdnn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    syms = header_symbols()
    assert len(syms) >= 35
    lib = C.CDLL(qd.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/fdnn.h but not exported"
    assert set(syms) == set(qd.SIGNATURES), set(syms) ^ set(qd.SIGNATURES)


def test_jni_symbols_exported_with_reference_names():
    out = subprocess.run(["nm", "-D", "--defined-only", qd.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    for s in qd.JNI_SYMBOLS:
        assert s in exported
    assert len(qd.JNI_SYMBOLS) == 11
    ref_header = "/root/reference/src/cpp/suskun_nn_QuantizedDnn.h"
    if os.path.exists(ref_header):
        declared = set(re.findall(r"(Java_suskun_nn_QuantizedDnn_\w+)", open(ref_header).read()))
        assert declared == set(qd.JNI_SYMBOLS)


def test_product_does_not_link_or_import_the_oracle():
    out = subprocess.run(["nm", "-D", qd.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "fdo_" not in out and "ref_calculate" not in out
    pkg = os.path.join(ROOT, "fast-dnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cc", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_py" not in text and "fdnn_oracle" not in text and "libfastdnn_ref" not in text, f


def test_version_and_lut_are_host_only():
    assert b"sm_100a" in qd.lib().fdnn_version()
    lut = qd.sigmoid_lut()
    assert lut[640] == 128 and lut[0] == 0 and lut[-1] >= 254


@pytest.mark.skipif(_cuda_available(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_no_compute(net_file):
    with pytest.raises(qd.NoGpuError):
        qd.QuantizedDnn.load_from_file(net_file("tiny"))
    blob = qd.pack(net_file("tiny"))  # host half still works
    with pytest.raises(qd.NoGpuError):
        qd.QuantizedDnn.load_from_blob(blob)
    with pytest.raises(qd.FdnnError):
        qd.PinnedArray((4, 4), np.float32)
    with pytest.raises(qd.NoGpuError):  # a device group is no way around it either
        qd.QuantizedDnn.load_on_devices(net_file("tiny"), [0, 1])
    os.environ["FDNN_DEVICES"] = "all"
    try:
        with pytest.raises(qd.NoGpuError):
            qd.QuantizedDnn.load_from_file(net_file("tiny"))
    finally:
        os.environ.pop("FDNN_DEVICES", None)
    assert qd.lib().fdnn_device_count(None) == qd.FDNN_EINVAL and qd.lib().fdnn_nccl_broadcast_count() == 0


def test_bad_arguments_are_refused_not_crashed_on(net_file):
    L = qd.lib()
    h = C.c_void_p()
    assert L.fdnn_load_devices(net_file("tiny").encode(), 3.0, None, 2, C.byref(h)) == qd.FDNN_EINVAL
    assert L.fdnn_load_devices(net_file("tiny").encode(), 3.0, (C.c_int * 1)(0), 0, C.byref(h)) == qd.FDNN_EINVAL
    assert L.fdnn_calculate_sink(None, None, 4, 12, None, None) == qd.FDNN_EINVAL
    assert L.fdnn_ctx_hidden_digest(None, 0, 0, None) == qd.FDNN_EINVAL
    assert L.fdnn_ctx_profile_pass(None, None, 1, None, 1, None, None) == qd.FDNN_EINVAL
    done = C.c_longlong(7)
    assert L.fdnn_calculate_file(None, b"/tmp/in.bin", b"/tmp/out.bin", qd.FDNN_DUMP_BIN, 0, C.byref(done)) == qd.FDNN_EINVAL
    assert done.value == 0  # nothing was written
    assert L.fdnn_output_dump_write_txt(None, None, 1, 1) == qd.FDNN_EINVAL
    assert L.fdnn_output_dump_write_txt(b"/nonexistent-dir/x.txt", (C.c_float * 1)(1.0), 1, 1) == qd.FDNN_EIO


def test_allocation_failures_do_not_cross_the_abi():
    """an impossible request comes back as an error code (function-try-blocks around the allocating entry points), not as
    a C++ exception unwinding into the caller"""
    L = qd.lib()
    one = (C.c_float * 4)(1.0, 2.0, 3.0, 4.0)
    rc = L.fdnn_feature_bin_write(b"/tmp/fdnn_never_written.bin", one, 2**31 - 1, 2**31 - 1)
    assert rc in (qd.FDNN_ENOMEM, qd.FDNN_ECUDA) and L.fdnn_last_error()
    assert not os.path.exists("/tmp/fdnn_never_written.bin")


def test_cutoff_must_be_positive(net_file):
    with pytest.raises(ValueError):
        qd.QuantizedDnn.load_from_file(net_file("tiny"), 0.0)


def test_bench_reference_arm_contract():
    """--impl reference prints one JSON line with the keys the driver reads (tiny run)"""
    import json
    env = dict(os.environ, FDNN_BENCH_SHAPE_FOR_TEST="S")
    out = subprocess.run(["python", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, check=True, env=env).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
