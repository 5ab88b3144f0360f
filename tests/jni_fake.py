"""A fabricated JNIEnv for driving the Java_suskun_nn_QuantizedDnn_* entry points without a JVM
(there is no JDK in the image).  JNIEnv is a pointer to a pointer to a table of function pointers
(reference: include/linux/jni.h, struct JNINativeInterface_); this module fills the slots the shim
uses with ctypes callbacks that keep "Java objects" in a Python registry."""
import ctypes as C

import numpy as np

SLOTS = {
    "FindClass": 6, "ThrowNew": 14, "GetStringUTFChars": 169, "ReleaseStringUTFChars": 170, "GetArrayLength": 171,
    "NewFloatArray": 181, "GetByteArrayElements": 184, "GetFloatArrayElements": 189, "ReleaseByteArrayElements": 192,
    "ReleaseFloatArrayElements": 197, "SetFloatArrayRegion": 213,
}
TABLE_SIZE = 233
JNI_ABORT = 2


class FakeJvm:
    def __init__(self):
        self.objects = {}          # handle → python object (bytes for strings, numpy arrays for arrays)
        self.next_handle = 0x1000
        self.thrown = []           # (class name, message)
        self.release_modes = []
        self.classes = {}
        self._keep = []
        table = (C.c_void_p * TABLE_SIZE)()

        def reg(name, restype, argtypes, fn):
            proto = C.CFUNCTYPE(restype, *argtypes)
            cb = proto(fn)
            self._keep.append(cb)
            table[SLOTS[name]] = C.cast(cb, C.c_void_p).value

        P, I = C.c_void_p, C.c_int
        reg("FindClass", P, [P, C.c_char_p], self._find_class)
        reg("ThrowNew", I, [P, P, C.c_char_p], self._throw_new)
        reg("GetStringUTFChars", P, [P, P, P], lambda env, s, iscopy: C.cast(self.objects[s], C.c_void_p).value)
        reg("ReleaseStringUTFChars", None, [P, P, P], lambda env, s, chars: None)
        reg("GetArrayLength", I, [P, P], lambda env, a: int(self.objects[a].size))
        reg("NewFloatArray", P, [P, I], lambda env, n: self.new_array(np.zeros(n, dtype=np.float32)))
        reg("GetByteArrayElements", P, [P, P, P], lambda env, a, iscopy: self.objects[a].ctypes.data)
        reg("GetFloatArrayElements", P, [P, P, P], lambda env, a, iscopy: self.objects[a].ctypes.data)
        reg("ReleaseByteArrayElements", None, [P, P, P, I], lambda env, a, p, mode: self.release_modes.append(mode))
        reg("ReleaseFloatArrayElements", None, [P, P, P, I], lambda env, a, p, mode: self.release_modes.append(mode))
        reg("SetFloatArrayRegion", None, [P, P, I, I, P], self._set_region)
        self.table = table
        self.table_ptr = C.c_void_p(C.addressof(table))   # JNIEnv  = const struct JNINativeInterface_ *
        self.env = C.pointer(self.table_ptr)               # JNIEnv* = what native methods receive

    def _handle(self):
        self.next_handle += 0x10
        return self.next_handle

    def _find_class(self, env, name):
        h = self._handle()
        self.classes[h] = name.decode()
        return h

    def _throw_new(self, env, cls, msg):
        self.thrown.append((self.classes.get(cls, "?"), msg.decode("utf-8", "replace")))
        return 0

    def _set_region(self, env, arr, start, n, buf):
        src = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_float)), shape=(n,)) if n else np.zeros(0, np.float32)
        self.objects[arr][start:start + n] = src

    def new_string(self, text: str):
        h = self._handle()
        self.objects[h] = C.create_string_buffer(text.encode())
        return h

    def new_array(self, a: np.ndarray):
        h = self._handle()
        self.objects[h] = np.ascontiguousarray(a)
        return h


def bind(lib):
    """argtypes/restypes of the eleven native methods (suskun_nn_QuantizedDnn.h:15-96)."""
    P, I, L, F = C.c_void_p, C.c_int, C.c_int64, C.c_float
    pre = "Java_suskun_nn_QuantizedDnn_"
    sigs = {
        "initialize": (L, [P, P, P, F]), "inputDimension": (I, [P, P, L]), "outputDimension": (I, [P, P, L]),
        "calculate": (P, [P, P, L, P, I, I, I]), "getContext": (L, [P, P, L, I, I]), "calculateUntilOutput": (None, [P, P, L, P]),
        "calculateLazy": (P, [P, P, L, I, P]), "deleteLazyContext": (None, [P, P, L]), "delete": (None, [P, P, L]),
        "layerDimension": (I, [P, P, L, I]), "layerCount": (I, [P, P, L]),
    }
    out = {}
    for name, (res, args) in sigs.items():
        fn = getattr(lib, pre + name)
        fn.restype, fn.argtypes = res, args
        out[name] = fn
    return out
