"""The JNI surface of libfast-dnn.so, driven through a fabricated JNIEnv (tests/jni_fake.py)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, REFERENCE_ROOT, _cuda_available
from fast_dnn_b200 import quantized_dnn as qd
from fast_dnn_b200 import synth
import jni_fake


def test_slot_indices_match_the_jni_function_table(tmp_path):
    """csrc/jni_min.h hard-codes function-table slots; check them against the reference's vendored
    jni.h (include/linux/jni.h) when that tree exists, and against tests/jni_fake.py always."""
    text = open(os.path.join(ROOT, "fast-dnn_b200", "csrc", "jni_min.h")).read()
    ours = {m.group(1): int(m.group(2)) for m in re.finditer(r"kJni(\w+)\s*=\s*(\d+)", text)}
    for name, slot in jni_fake.SLOTS.items():
        assert ours[name] == slot
    assert ours["TableSize"] == jni_fake.TABLE_SIZE
    jni_h = os.path.join(REFERENCE_ROOT, "include", "linux")
    if not os.path.isdir(jni_h):
        pytest.skip("vendored jni.h not present on this machine")
    src = tmp_path / "idx.cc"
    fields = [n for n in jni_fake.SLOTS]
    src.write_text("#include <jni.h>\n#include <cstddef>\n#include <cstdio>\nint main(){\n" +
                   "".join(f'printf("{f} %zu\\n", offsetof(JNINativeInterface_, {f})/sizeof(void*));\n' for f in fields) +
                   'printf("TableSize %zu\\n", sizeof(JNINativeInterface_)/sizeof(void*));\n'
                   'printf("sizes %zu %zu %zu\\n", sizeof(jlong), sizeof(jint), sizeof(jbyte));\n}\n')
    exe = tmp_path / "idx"
    subprocess.run(["g++", f"-I{jni_h}", str(src), "-o", str(exe)], check=True)
    out = dict(line.split(" ", 1) for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    for f in fields:
        assert int(out[f]) == ours[f], f
    assert int(out["TableSize"]) == ours["TableSize"]
    assert out["sizes"] == "8 4 1"


@pytest.mark.skipif(_cuda_available(), reason="checks the no-GPU behaviour")
def test_initialize_without_gpu_throws_illegal_state(net_file):
    jvm = jni_fake.FakeJvm()
    fn = jni_fake.bind(C.CDLL(qd.LIB_PATH))
    h = fn["initialize"](jvm.env, None, jvm.new_string(net_file("tiny")), 3.0)
    assert h == 0
    assert jvm.thrown and jvm.thrown[0][0] == "java/lang/IllegalStateException" and "CUDA" in jvm.thrown[0][1]


@pytest.mark.gpu
def test_full_java_call_sequence(net_file):
    """what QuantizedDnn.java does: loadFromFile → calculate; getNewLazyContext → calculateUntilOutput →
    calculateForOutputNodes per frame → delete"""
    jvm = jni_fake.FakeJvm()
    fn = jni_fake.bind(C.CDLL(qd.LIB_PATH))
    path = net_file("S")
    h = fn["initialize"](jvm.env, None, jvm.new_string(path), 3.0)
    assert h != 0 and not jvm.thrown
    assert fn["inputDimension"](jvm.env, None, h) == 440 and fn["outputDimension"](jvm.env, None, h) == 2000
    assert fn["layerCount"](jvm.env, None, h) == 5
    assert [fn["layerDimension"](jvm.env, None, h, i) for i in range(6)] == [512, 512, 512, 2000, -1, -1]

    n = 40
    frames = synth.make_frames(n, 440, seed=23)
    j_in = jvm.new_array(frames.reshape(-1).copy())
    j_out = fn["calculate"](jvm.env, None, h, j_in, n, 440, 10)
    assert j_out and not jvm.thrown
    got = jvm.objects[j_out].reshape(n, 2000)
    direct = qd.QuantizedDnn.load_from_file(path)
    want = direct.calculate(frames)
    assert np.array_equal(got, want)
    assert np.array_equal(jvm.objects[j_in], frames.reshape(-1)), "Java input array must not be modified"
    assert set(jvm.release_modes) == {jni_fake.JNI_ABORT}

    ctx = fn["getContext"](jvm.env, None, h, n, 8)
    assert ctx != 0
    fn["calculateUntilOutput"](jvm.env, None, ctx, j_in)
    masks = synth.make_masks(n, 2000, seed=5)
    lazy_ctx = direct.get_new_lazy_context(n)
    lazy_ctx.calculate_until_output(frames)
    for i in range(0, n, 7):
        j_mask = jvm.new_array(masks[i])
        j_row = fn["calculateLazy"](jvm.env, None, ctx, i, j_mask)
        lazy_ctx.current_vector_index = i
        assert np.array_equal(jvm.objects[j_row], lazy_ctx.calculate_for_output_nodes(masks[i]))
    # a mask of the wrong length is refused instead of being read past its end (jni_dnn.cc:97-117 trusts it)
    assert not fn["calculateLazy"](jvm.env, None, ctx, 0, jvm.new_array(np.ones(10, np.int8)))
    assert jvm.thrown and "mask length" in jvm.thrown[-1][1]
    # arrays shorter than what the call announces are refused as well (jni_dnn.cc:44-47, 89-91 read count × dim floats regardless)
    jvm.thrown.clear()
    j_short = jvm.new_array(frames.reshape(-1)[: 440 * (n - 1)].copy())
    assert not fn["calculate"](jvm.env, None, h, j_short, n, 440, 10)
    assert jvm.thrown and "shorter" in jvm.thrown[-1][1]
    jvm.thrown.clear()
    fn["calculateUntilOutput"](jvm.env, None, ctx, j_short)
    assert jvm.thrown and "shorter" in jvm.thrown[-1][1]
    # wrong input dimension → exception, no crash
    jvm.thrown.clear()
    assert not fn["calculate"](jvm.env, None, h, j_in, n, 436, 10)
    assert jvm.thrown and jvm.thrown[0][0] == "java/lang/IllegalStateException"
    lazy_ctx.delete()
    direct.delete()
    fn["deleteLazyContext"](jvm.env, None, ctx)
    fn["delete"](jvm.env, None, h)
