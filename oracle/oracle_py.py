"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU oracles.

  Port  → oracle/libfdnn_oracle.so        (plain-C restatement, oracle/fdnn_oracle.c)
  Ref   → oracle/_ref/libfastdnn_ref.so    (the unmodified reference compiled by oracle/Makefile)

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libfdnn_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libfastdnn_ref.so")
REFERENCE_SRC = "/root/reference/src/cpp"

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")


def build(port: bool = True, ref: bool = True) -> None:
    """Compile the checker libraries (building the checker is not using it)."""
    if port:
        subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref and os.path.isdir(REFERENCE_SRC):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class Port:
    """Plain-C restatement."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(PORT_SO):
                build(port=True, ref=False)
            L = C.CDLL(PORT_SO)
            L.fdo_load.restype = C.c_void_p
            L.fdo_load.argtypes = [C.c_char_p, C.c_float]
            L.fdo_free.argtypes = [C.c_void_p]
            for name in ("fdo_input_dim", "fdo_output_dim", "fdo_hidden_dim", "fdo_qlayer_count"):
                getattr(L, name).argtypes = [C.c_void_p]
                getattr(L, name).restype = C.c_int
            for name in ("fdo_qlayer_nodes", "fdo_qlayer_inputs"):
                getattr(L, name).argtypes = [C.c_void_p, C.c_int]
                getattr(L, name).restype = C.c_int
            L.fdo_qlayer_multiplier.argtypes = [C.c_void_p, C.c_int]
            L.fdo_qlayer_multiplier.restype = C.c_float
            L.fdo_qlayer_weights.argtypes = [C.c_void_p, C.c_int]
            L.fdo_qlayer_weights.restype = C.POINTER(C.c_int8)
            L.fdo_qlayer_bias.argtypes = [C.c_void_p, C.c_int]
            L.fdo_qlayer_bias.restype = C.POINTER(C.c_float)
            for name in ("fdo_input_weights", "fdo_input_bias", "fdo_shift", "fdo_scale"):
                getattr(L, name).argtypes = [C.c_void_p]
                getattr(L, name).restype = C.POINTER(C.c_float)
            L.fdo_sigmoid_lut.argtypes = [_u8p]
            L.fdo_qsigmoid.argtypes = [C.c_float]
            L.fdo_qsigmoid.restype = C.c_uint8
            L.fdo_node_sum.argtypes = [C.c_int, _u8p, _i8p]
            L.fdo_node_sum.restype = C.c_int32
            L.fdo_node_sum_nosat.argtypes = [C.c_int, _u8p, _i8p]
            L.fdo_node_sum_nosat.restype = C.c_int32
            L.fdo_hidden_trace.argtypes = [C.c_void_p, _f32p, C.c_int, _u8p]
            L.fdo_until_output.argtypes = [C.c_void_p, _f32p, C.c_int, _u8p, C.c_int]
            L.fdo_output_linear.argtypes = [C.c_void_p, _u8p, C.c_int, _f32p, C.c_int]
            L.fdo_softmax.argtypes = [_f32p, C.c_int]
            L.fdo_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, _f32p, C.c_int]
            L.fdo_lazy.argtypes = [C.c_void_p, _u8p, _i8p, _f32p]
            L.fdo_time_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_void_p]
            L.fdo_time_calculate.restype = C.c_double
            cls._lib = L
        return cls._lib

    def __init__(self, path: str, cutoff: float = 3.0):
        self.L = self.lib()
        self.h = self.L.fdo_load(os.fsencode(path), cutoff)
        if not self.h:
            raise IOError(f"oracle port could not load {path}")
        self.input_dim = self.L.fdo_input_dim(self.h)
        self.output_dim = self.L.fdo_output_dim(self.h)
        self.hidden_dim = self.L.fdo_hidden_dim(self.h)
        self.qlayer_count = self.L.fdo_qlayer_count(self.h)

    def close(self):
        if self.h:
            self.L.fdo_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def qlayer(self, i):
        n, k = self.L.fdo_qlayer_nodes(self.h, i), self.L.fdo_qlayer_inputs(self.h, i)
        w = np.ctypeslib.as_array(self.L.fdo_qlayer_weights(self.h, i), shape=(n, k)).copy()
        b = np.ctypeslib.as_array(self.L.fdo_qlayer_bias(self.h, i), shape=(n,)).copy()
        return w, b, float(self.L.fdo_qlayer_multiplier(self.h, i))

    def input_layer(self):
        H, I = self.hidden_dim, self.input_dim
        w = np.ctypeslib.as_array(self.L.fdo_input_weights(self.h), shape=(H, I)).copy()
        b = np.ctypeslib.as_array(self.L.fdo_input_bias(self.h), shape=(H,)).copy()
        sh = np.ctypeslib.as_array(self.L.fdo_shift(self.h), shape=(I,)).copy()
        sc = np.ctypeslib.as_array(self.L.fdo_scale(self.h), shape=(I,)).copy()
        return w, b, sh, sc

    @classmethod
    def sigmoid_lut(cls):
        out = np.zeros(1280, dtype=np.uint8)
        cls.lib().fdo_sigmoid_lut(out)
        return out

    @classmethod
    def qsigmoid(cls, x: float) -> int:
        return int(cls.lib().fdo_qsigmoid(float(x)))

    @classmethod
    def node_sum(cls, a, w, saturate=True):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        w = np.ascontiguousarray(w, dtype=np.int8)
        fn = cls.lib().fdo_node_sum if saturate else cls.lib().fdo_node_sum_nosat
        return int(fn(len(a), a, w))

    def hidden_trace(self, frames):
        x = _f32(frames)
        out = np.zeros((self.qlayer_count, x.shape[0], self.hidden_dim), dtype=np.uint8)
        self.L.fdo_hidden_trace(self.h, x, x.shape[0], out)
        return out

    def until_output(self, frames, threads=8):
        x = _f32(frames)
        out = np.zeros((x.shape[0], self.hidden_dim), dtype=np.uint8)
        self.L.fdo_until_output(self.h, x, x.shape[0], out, threads)
        return out

    def output_linear(self, hidden, threads=8):
        hidden = np.ascontiguousarray(hidden, dtype=np.uint8)
        out = np.zeros((hidden.shape[0], self.output_dim), dtype=np.float32)
        self.L.fdo_output_linear(self.h, hidden, hidden.shape[0], out, threads)
        return out

    @classmethod
    def softmax(cls, row):
        row = _f32(row).copy()
        cls.lib().fdo_softmax(row, row.size)
        return row

    def calculate(self, frames, threads=8):
        x = _f32(frames)
        out = np.zeros((x.shape[0], self.output_dim), dtype=np.float32)
        if x.shape[0]:
            self.L.fdo_calculate(self.h, x, x.shape[0], out, threads)
        return out

    def lazy(self, hidden_row, mask):
        out = np.zeros(self.output_dim, dtype=np.float32)
        self.L.fdo_lazy(self.h, np.ascontiguousarray(hidden_row, dtype=np.uint8),
                        np.ascontiguousarray(mask, dtype=np.int8), out)
        return out

    def time_calculate(self, frames, threads=1) -> float:
        x = _f32(frames)
        return float(self.L.fdo_time_calculate(self.h, x, x.shape[0], threads, None))


class Ref:
    """The compiled, unmodified reference (oracle/_ref)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                build(port=False, ref=True)
            L = C.CDLL(REF_SO)
            L.ref_load.restype = C.c_void_p
            L.ref_load.argtypes = [C.c_char_p, C.c_float]
            L.ref_free.argtypes = [C.c_void_p]
            for name in ("ref_input_dim", "ref_output_dim", "ref_qlayer_count", "ref_hidden_dim"):
                getattr(L, name).argtypes = [C.c_void_p]
                getattr(L, name).restype = C.c_int
            for name in ("ref_qlayer_nodes", "ref_qlayer_inputs"):
                getattr(L, name).argtypes = [C.c_void_p, C.c_int]
                getattr(L, name).restype = C.c_int
            L.ref_qlayer_multiplier.argtypes = [C.c_void_p, C.c_int]
            L.ref_qlayer_multiplier.restype = C.c_float
            L.ref_qlayer_weights.argtypes = [C.c_void_p, C.c_int, _i8p]
            L.ref_qlayer_bias.argtypes = [C.c_void_p, C.c_int, _f32p]
            L.ref_input_weights.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p]
            L.ref_sigmoid_lut.argtypes = [_u8p]
            L.ref_qsigmoid.argtypes = [C.c_float]
            L.ref_qsigmoid.restype = C.c_uint8
            L.ref_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
            L.ref_ctx_new.restype = C.c_void_p
            L.ref_ctx_new.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.ref_ctx_free.argtypes = [C.c_void_p]
            L.ref_ctx_until_output.argtypes = [C.c_void_p, _f32p]
            L.ref_ctx_hidden.argtypes = [C.c_void_p, _u8p]
            L.ref_ctx_output_linear.argtypes = [C.c_void_p, _f32p]
            L.ref_ctx_lazy.argtypes = [C.c_void_p, C.c_int, _i8p, _f32p]
            L.ref_hidden_trace.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, _u8p]
            L.ref_time_calculate.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
            L.ref_time_calculate.restype = C.c_double
            L.ref_time_lazy.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, _i8p, C.c_int, C.c_void_p]
            L.ref_time_lazy.restype = C.c_double
            L.ref_float_load.restype = C.c_void_p
            L.ref_float_load.argtypes = [C.c_char_p]
            L.ref_float_free.argtypes = [C.c_void_p]
            for name in ("ref_float_layer_count",):
                getattr(L, name).argtypes = [C.c_void_p]
                getattr(L, name).restype = C.c_int
            for name in ("ref_float_layer_inputs", "ref_float_layer_nodes"):
                getattr(L, name).argtypes = [C.c_void_p, C.c_int]
                getattr(L, name).restype = C.c_int
            L.ref_float_layer.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p]
            L.ref_float_shift_scale.argtypes = [C.c_void_p, _f32p, _f32p]
            L.ref_cli.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
            L.ref_cli.restype = C.c_int
            cls._lib = L
        return cls._lib

    @classmethod
    def cli(cls, model_path: str, input_path: str, out_path: str, binary: bool = True) -> None:
        """The reference's own command-line driver (dnn.cc:20-83): feature file in, score dump out (BIN or TXT).
        Runs in a child process that loads nothing but the reference library: the driver prints numbers through
        std::cout, which crashes inside a process where numpy/torch have brought their own C++ runtime pieces along."""
        for p in (model_path, input_path):
            if not os.path.exists(p):  # the reference dereferences a null FILE* (float_dnn.cc:171-174)
                raise IOError(p)
        cls.lib()  # builds it if need be
        code = ("import ctypes, sys; L = ctypes.CDLL(sys.argv[1]); L.ref_cli.argtypes = [ctypes.c_char_p] * 4; "
                "sys.exit(L.ref_cli(*[a.encode() for a in sys.argv[2:6]]))")
        r = subprocess.run([sys.executable, "-c", code, REF_SO, model_path, input_path, out_path, "BIN" if binary else "TXT"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
        if r.returncode != 0:
            raise RuntimeError(f"reference CLI returned {r.returncode}: {r.stdout.decode(errors='replace')[-400:]}")

    @classmethod
    def load_float_network(cls, path: str):
        """dnn.bin through the reference's own fp32 loader (FloatDnn, float_dnn.cc:18-69) →
        ([(W [nodes][inputs], bias)], shift, scale); layer-0 inputs padded to ×4 as the loader does."""
        L = cls.lib()
        if not os.path.exists(path):
            raise IOError(path)
        h = L.ref_float_load(os.fsencode(path))
        try:
            layers = []
            for i in range(L.ref_float_layer_count(h)):
                n, k = L.ref_float_layer_nodes(h, i), L.ref_float_layer_inputs(h, i)
                w, b = np.zeros((n, k), dtype=np.float32), np.zeros(n, dtype=np.float32)
                L.ref_float_layer(h, i, w, b)
                layers.append((w, b))
            I = layers[0][0].shape[1]
            shift, scale = np.zeros(I, dtype=np.float32), np.zeros(I, dtype=np.float32)
            L.ref_float_shift_scale(h, shift, scale)
            return layers, shift, scale
        finally:
            L.ref_float_free(h)

    def __init__(self, path: str, cutoff: float = 3.0):
        self.L = self.lib()
        if not os.path.exists(path):
            raise IOError(path)  # the reference dereferences a null FILE* (float_dnn.cc:171-174)
        self.h = self.L.ref_load(os.fsencode(path), cutoff)
        self.input_dim = self.L.ref_input_dim(self.h)
        self.output_dim = self.L.ref_output_dim(self.h)
        self.hidden_dim = self.L.ref_hidden_dim(self.h)
        self.qlayer_count = self.L.ref_qlayer_count(self.h)

    def close(self):
        if self.h:
            self.L.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def qlayer(self, i):
        n, k = self.L.ref_qlayer_nodes(self.h, i), self.L.ref_qlayer_inputs(self.h, i)
        w = np.zeros((n, k), dtype=np.int8)
        b = np.zeros(n, dtype=np.float32)
        self.L.ref_qlayer_weights(self.h, i, w)
        self.L.ref_qlayer_bias(self.h, i, b)
        return w, b, float(self.L.ref_qlayer_multiplier(self.h, i))

    def input_layer(self):
        H, I = self.hidden_dim, self.input_dim
        w = np.zeros((H, I), dtype=np.float32)
        b = np.zeros(H, dtype=np.float32)
        sh = np.zeros(I, dtype=np.float32)
        sc = np.zeros(I, dtype=np.float32)
        self.L.ref_input_weights(self.h, w, b, sh, sc)
        return w, b, sh, sc

    @classmethod
    def sigmoid_lut(cls):
        out = np.zeros(1280, dtype=np.uint8)
        cls.lib().ref_sigmoid_lut(out)
        return out

    @classmethod
    def qsigmoid(cls, x: float) -> int:
        return int(cls.lib().ref_qsigmoid(float(x)))

    def calculate(self, frames, batch=10):
        x = _f32(frames)
        out = np.zeros((x.shape[0], self.output_dim), dtype=np.float32)
        if x.shape[0]:
            self.L.ref_calculate(self.h, x, x.shape[0], x.shape[1], batch, out)
        return out

    def hidden_trace(self, frames, batch=10):
        x = _f32(frames)
        out = np.zeros((self.qlayer_count, x.shape[0], self.hidden_dim), dtype=np.uint8)
        self.L.ref_hidden_trace(self.h, x, x.shape[0], batch, out)
        return out

    def lazy_context(self, n, batch=8):
        return RefContext(self, n, batch)

    def time_calculate(self, frames, batch=10, threads=1, out=None) -> float:
        x = _f32(frames)
        ptr = out.ctypes.data_as(C.c_void_p) if out is not None else None
        return float(self.L.ref_time_calculate(self.h, x, x.shape[0], x.shape[1], batch, threads, ptr))


    def time_lazy(self, frames, masks, batch=8, threads=1) -> float:
        """wall seconds of the lazy protocol (until_output + one LazyOutputActivations per frame) over `threads` host threads"""
        x = _f32(frames)
        m = np.ascontiguousarray(masks, dtype=np.int8)
        assert m.shape == (x.shape[0], self.output_dim)
        return float(self.L.ref_time_lazy(self.h, x, x.shape[0], x.shape[1], batch, m, threads, None))


class RefContext:
    def __init__(self, ref: Ref, n: int, batch: int):
        self.ref, self.n = ref, n
        self.c = ref.L.ref_ctx_new(ref.h, n, batch)

    def until_output(self, frames):
        x = _f32(frames)
        assert x.shape == (self.n, self.ref.input_dim)
        self.ref.L.ref_ctx_until_output(self.c, x)

    def hidden(self):
        out = np.zeros((self.n, self.ref.hidden_dim), dtype=np.uint8)
        self.ref.L.ref_ctx_hidden(self.c, out)
        return out

    def output_linear(self):
        out = np.zeros((self.n, self.ref.output_dim), dtype=np.float32)
        self.ref.L.ref_ctx_output_linear(self.c, out)
        return out

    def lazy(self, idx, mask):
        out = np.zeros(self.ref.output_dim, dtype=np.float32)
        self.ref.L.ref_ctx_lazy(self.c, idx, np.ascontiguousarray(mask, dtype=np.int8), out)
        return out

    def close(self):
        if self.c:
            self.ref.L.ref_ctx_free(self.c)
            self.c = None
