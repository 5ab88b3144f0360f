"""Host-side mirror of the reference's ``suskun.nn.QuantizedDnn`` over the C ABI of libfast-dnn.so.

The reference's public surface is the Java class
/root/reference/src/java/suskun/nn/QuantizedDnn.java (loadFromFile :54-70, calculate :149-167,
getNewLazyContext :100-107, LazyContext :72-98).  No JVM exists in this image, so the same
operations are exposed here with the same names, argument meaning and error behaviour, as thin
ctypes calls into ``include/fdnn.h`` — the very entry points the JNI symbols in
``csrc/jni_shim.cc`` wrap.  All arithmetic runs in the CUDA library; nothing here computes, and
there is no CPU fallback: without the built library or without a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfast-dnn.so")

FDNN_OK, FDNN_EINVAL, FDNN_EIO, FDNN_EFORMAT, FDNN_ENOGPU, FDNN_ECUDA, FDNN_ENOMEM = 0, -1, -2, -3, -4, -5, -6
FDNN_DUMP_BIN, FDNN_DUMP_TXT = 1, 0


class FdnnError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fdnn error {code}: {message}")
        self.code = code


class NoGpuError(FdnnError):
    pass


_lib = None

# name → (restype, argtypes); every symbol declared in include/fdnn.h
_P, _I, _F, _SZ, _LL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
SIGNATURES = {
    "fdnn_last_error": (C.c_char_p, []),
    "fdnn_version": (C.c_char_p, []),
    "fdnn_load": (_I, [C.c_char_p, _F, _I, C.POINTER(_P)]),
    "fdnn_load_devices": (_I, [C.c_char_p, _F, C.POINTER(_I), _I, C.POINTER(_P)]),
    "fdnn_device_count": (_I, [_P]),
    "fdnn_device_at": (_I, [_P, _I]),
    "fdnn_nccl_broadcast_count": (_LL, []),
    "fdnn_shard_plan": (_I, [_I, _I, C.POINTER(_I), C.POINTER(_I)]),
    "fdnn_pack": (_I, [C.c_char_p, _F, C.POINTER(_P), C.POINTER(_SZ)]),
    "fdnn_blob_free": (None, [_P]),
    "fdnn_load_blob": (_I, [_P, _SZ, _I, C.POINTER(_P)]),
    "fdnn_align_dnn_bin": (_I, [C.c_char_p, C.c_char_p, _I, _I]),
    "fdnn_import_kaldi_nnet1": (_I, [C.c_char_p, C.c_char_p, C.c_char_p]),
    "fdnn_feature_bin_read": (_I, [C.c_char_p, C.POINTER(_I), C.POINTER(_I), C.POINTER(_P)]),
    "fdnn_feature_bin_write": (_I, [C.c_char_p, _P, _I, _I]),
    "fdnn_output_dump_write": (_I, [C.c_char_p, _P, _I, _I]),
    "fdnn_output_dump_write_txt": (_I, [C.c_char_p, _P, _I, _I]),
    "fdnn_free": (_I, [_P]),
    "fdnn_input_dim": (_I, [_P]),
    "fdnn_output_dim": (_I, [_P]),
    "fdnn_layer_count": (_I, [_P]),
    "fdnn_set_tile_policy": (_I, [_P, _I]),
    "fdnn_layer_dim": (_I, [_P, _I]),
    "fdnn_hidden_dim": (_I, [_P]),
    "fdnn_device": (_I, [_P]),
    "fdnn_calculate": (_I, [_P, _P, _I, _I, _I, _P]),
    "fdnn_calculate_sink": (_I, [_P, _P, _I, _I, _P, _P]),
    "fdnn_calculate_file": (_I, [_P, C.c_char_p, C.c_char_p, _I, _I, C.POINTER(_LL)]),
    "fdnn_ctx_new": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "fdnn_ctx_free": (_I, [_P]),
    "fdnn_ctx_frames": (_I, [_P]),
    "fdnn_ctx_output_dim": (_I, [_P]),
    "fdnn_ctx_input_dim": (_I, [_P]),
    "fdnn_ctx_until_output": (_I, [_P, _P]),
    "fdnn_ctx_lazy": (_I, [_P, _I, _P, _P]),
    "fdnn_ctx_lazy_batch": (_I, [_P, _P, _P]),
    "fdnn_ctx_forward_device": (_I, [_P, _P, _I, _P, _P]),
    "fdnn_ctx_until_output_device": (_I, [_P, _P, _I, _P]),
    "fdnn_ctx_lazy_batch_device": (_I, [_P, _P, _I, _P, _P]),
    "fdnn_ctx_profile_stages": (_I, [_P, _P, _I, _P, _I, _P]),
    "fdnn_ctx_profile_pass": (_I, [_P, _P, _I, _P, _I, _P, C.POINTER(_I)]),
    "fdnn_ctx_timeline": (_I, [_P, _I, _P]),
    "fdnn_ctx_input_undecided": (_I, [_P, C.POINTER(C.c_uint)]),
    "fdnn_ctx_set_trace": (_I, [_P, _I]),
    "fdnn_ctx_hidden": (_I, [_P, _I, _I, _P]),
    "fdnn_ctx_hidden_digest": (_I, [_P, _I, _I, C.POINTER(C.c_ulonglong)]),
    "fdnn_ctx_logits": (_I, [_P, _I, _P]),
    "fdnn_model_qlayer": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_F), _P, _P]),
    "fdnn_model_fixup_count": (_I, [_P, _I]),
    "fdnn_model_fast_div": (_I, [_P, _I]),
    "fdnn_model_uses_tensor_cores": (_I, [_P, _I]),
    "fdnn_sigmoid_lut": (_I, [_P]),
    "fdnn_host_alloc": (_I, [C.POINTER(_P), _SZ]),
    "fdnn_host_free": (_I, [_P]),
    "fdnn_launch_count": (_LL, []),
}

JNI_SYMBOLS = [
    "Java_suskun_nn_QuantizedDnn_" + n
    for n in ("initialize", "inputDimension", "outputDimension", "calculate", "getContext", "calculateUntilOutput",
              "calculateLazy", "deleteLazyContext", "delete", "layerDimension", "layerCount")
]


def lib() -> C.CDLL:
    """Loads libfast-dnn.so (fails loudly when it has not been built: no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                              "(or __graft_entry__.build()); there is no non-CUDA implementation")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != FDNN_OK:
        msg = (lib().fdnn_last_error() or b"").decode("utf-8", "replace")
        raise (NoGpuError if rc == FDNN_ENOGPU else FdnnError)(rc, msg)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _frames(input_, dim: int) -> np.ndarray:
    x = np.ascontiguousarray(input_, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError("input must be a [frames][dimension] matrix")
    if x.shape[1] != dim:  # QuantizedDnn.java:157-161
        raise ValueError(f"Input vector size {x.shape[1]} must be equal with network input size {dim}")
    return x


def pack(path: str, cutoff: float = 3.0) -> np.ndarray:
    """Host-only: dnn.bin → the relocatable model blob (uint8 array).  Works without a GPU."""
    blob, size = C.c_void_p(), C.c_size_t()
    _check(lib().fdnn_pack(os.fsencode(path), cutoff, C.byref(blob), C.byref(size)))
    try:
        return np.ctypeslib.as_array(C.cast(blob, C.POINTER(C.c_uint8)), shape=(size.value,)).copy()
    finally:
        lib().fdnn_blob_free(blob)


def shard_plan(n_frames: int, n_devices: int):
    """[(first, count)] per device: how calculate() on a device group cuts a call (host only)."""
    first, count = (C.c_int * n_devices)(), (C.c_int * n_devices)()
    used = lib().fdnn_shard_plan(n_frames, n_devices, first, count)
    if used < 0:
        raise ValueError("bad shard request")
    return [(first[d], count[d]) for d in range(n_devices)], used


def align_dnn_bin(in_path, out_path, input_alignment: int = 4, hidden_alignment: int = 16) -> None:
    """FeedForwardNetwork.align(4, 16) + saveBinary on the C++ side (host only)."""
    _check(lib().fdnn_align_dnn_bin(os.fsencode(in_path), os.fsencode(out_path), input_alignment, hidden_alignment))


def import_kaldi_nnet1(nnet_txt_path, transform_txt_path, out_path) -> None:
    """FeedForwardNetwork.loadFromTextFile + saveBinary on the C++ side (host only): Kaldi nnet1 text model and
    feature-transform text → unaligned dnn.bin (follow with align_dnn_bin)."""
    _check(lib().fdnn_import_kaldi_nnet1(os.fsencode(nnet_txt_path), os.fsencode(transform_txt_path), os.fsencode(out_path)))


def read_feature_bin(path) -> np.ndarray:
    n, d, data = C.c_int(), C.c_int(), C.c_void_p()
    _check(lib().fdnn_feature_bin_read(os.fsencode(path), C.byref(n), C.byref(d), C.byref(data)))
    try:
        return np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_float)), shape=(n.value * d.value,)).reshape(n.value, d.value).copy()
    finally:
        lib().fdnn_blob_free(data)


def write_feature_bin(path, frames) -> None:
    x = np.ascontiguousarray(frames, dtype=np.float32)
    _check(lib().fdnn_feature_bin_write(os.fsencode(path), _ptr(x), x.shape[0], x.shape[1]))


def write_output_dump(path, rows, binary: bool = True) -> None:
    """BatchData::dumpToFile(path, binary) — float_dnn.cc:128-164 (host only)."""
    x = np.ascontiguousarray(rows, dtype=np.float32)
    fn = lib().fdnn_output_dump_write if binary else lib().fdnn_output_dump_write_txt
    _check(fn(os.fsencode(path), _ptr(x), x.shape[0], x.shape[1]))


def read_output_dump(path) -> np.ndarray:
    """The binary dump back as [frames][dim] fp32 (native-endian uint32 frames, uint32 dim, fp32 rows)."""
    with open(path, "rb") as f:
        n, d = np.frombuffer(f.read(8), dtype=np.uint32)
        return np.fromfile(f, dtype=np.float32, count=int(n) * int(d)).reshape(int(n), int(d))


class PinnedArray:
    """numpy view over page-locked host memory from fdnn_host_alloc (zero-staging transfers)."""

    def __init__(self, shape, dtype):
        self._p = C.c_void_p()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(lib().fdnn_host_alloc(C.byref(self._p), max(self.nbytes, 1)))
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._p:
            self.array = None
            lib().fdnn_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class QuantizedDnn:
    """Mirror of suskun.nn.QuantizedDnn (QuantizedDnn.java:23-188)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        self._input_dimension = lib().fdnn_input_dim(self._h)
        self._output_dimension = lib().fdnn_output_dim(self._h)

    # -- loading ---------------------------------------------------------------------------------
    @classmethod
    def load_from_file(cls, dnn_file, weight_cut_off_value: float = 3.0, device: int = -1) -> "QuantizedDnn":
        """loadFromFile(File[, float]) — QuantizedDnn.java:54-70; cut-off must be positive (:55-57)."""
        if weight_cut_off_value <= 0:
            raise ValueError(f"Weight cut off value must be positive. But it is {weight_cut_off_value}")
        h = C.c_void_p()
        _check(lib().fdnn_load(os.fsencode(os.path.abspath(os.fspath(dnn_file))), weight_cut_off_value, device, C.byref(h)))
        return cls(h.value)

    loadFromFile = load_from_file

    @classmethod
    def load_on_devices(cls, dnn_file, devices, weight_cut_off_value: float = 3.0) -> "QuantizedDnn":
        """One handle over several GPUs (no counterpart in the reference): one ncclBroadcast of the packed weights at load,
        calculate() shards its frames over the devices, lazy contexts are handed out round-robin."""
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().fdnn_load_devices(os.fsencode(os.path.abspath(os.fspath(dnn_file))), weight_cut_off_value, devs, len(devices), C.byref(h)))
        return cls(h.value)

    def devices(self):
        return [lib().fdnn_device_at(self._h, i) for i in range(lib().fdnn_device_count(self._h))]

    @classmethod
    def load_from_blob(cls, blob, device: int = -1, size: int | None = None) -> "QuantizedDnn":
        """blob: numpy uint8 array (host) or an integer device pointer with `size` (e.g. the NCCL
        receive buffer of the load-time weight broadcast)."""
        h = C.c_void_p()
        if isinstance(blob, np.ndarray):
            b = np.ascontiguousarray(blob, dtype=np.uint8)
            _check(lib().fdnn_load_blob(_ptr(b), b.nbytes, device, C.byref(h)))
        else:
            _check(lib().fdnn_load_blob(C.c_void_p(int(blob)), int(size), device, C.byref(h)))
        return cls(h.value)

    # -- introspection -----------------------------------------------------------------------------
    def input_dimension(self) -> int:
        return self._input_dimension

    def output_dimension(self) -> int:
        return self._output_dimension

    def layer_count(self) -> int:
        return lib().fdnn_layer_count(self._h)

    def layer_dimension(self, layer_index: int) -> int:
        return lib().fdnn_layer_dim(self._h, layer_index)

    def hidden_dimension(self) -> int:
        return lib().fdnn_hidden_dim(self._h)

    def device(self) -> int:
        return lib().fdnn_device(self._h)

    def set_tile_policy(self, policy: str) -> None:
        """'latency' (default: shortest pass for one caller) or 'throughput' (least SM time per frame, for several contexts in
        flight on this model); results are identical.  A context (get_new_lazy_context, and the workspaces calculate() pools)
        takes the policy in force when it is created and keeps it."""
        _check(lib().fdnn_set_tile_policy(self._h, {"latency": 0, "throughput": 1}[policy]))

    inputDimension, outputDimension, layerCount, layerDimension = input_dimension, output_dimension, layer_count, layer_dimension

    def qlayer(self, i: int):
        """(int8 weights [N][K], fp32 bias [N], multiplier) of int8 layer i as held on the device."""
        n, k, mult = C.c_int(), C.c_int(), C.c_float()
        _check(lib().fdnn_model_qlayer(self._h, i, C.byref(n), C.byref(k), C.byref(mult), None, None))
        w = np.zeros((n.value, k.value), dtype=np.int8)
        b = np.zeros(n.value, dtype=np.float32)
        _check(lib().fdnn_model_qlayer(self._h, i, None, None, None, _ptr(w), _ptr(b)))
        return w, b, float(mult.value)

    def fixup_count(self, i: int) -> int:
        return lib().fdnn_model_fixup_count(self._h, i)

    def fast_div(self, i: int) -> bool:
        return bool(lib().fdnn_model_fast_div(self._h, i))

    def uses_tensor_cores(self, i: int) -> bool:
        return bool(lib().fdnn_model_uses_tensor_cores(self._h, i))

    # -- forward -------------------------------------------------------------------------------------
    def calculate(self, input_, batch_size: int = 10, out: np.ndarray | None = None) -> np.ndarray:
        """calculate(float[][] input[, batchSize]) — QuantizedDnn.java:149-167 → softmax rows [n][O].
        `batch_size` is the reference's CPU cache-blocking hint; results do not depend on it."""
        if len(input_) == 0:  # :154-156
            return np.zeros((0, 0), dtype=np.float32)
        x = _frames(input_, self._input_dimension)
        if out is None:
            out = np.empty((x.shape[0], self._output_dimension), dtype=np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (x.shape[0], self._output_dimension)
        _check(lib().fdnn_calculate(self._h, _ptr(x), x.shape[0], x.shape[1], batch_size, _ptr(out)))
        return out

    def calculate_file(self, feature_bin_path, out_path, binary: bool = True, chunk_frames: int = 0) -> int:
        """The reference's command-line data path (dnn.cc:55-78: BatchData(file) → Calculate → dumpToFile) file to file,
        streamed in chunks so that neither file is ever held in memory; returns the number of rows written."""
        done = C.c_longlong()
        _check(lib().fdnn_calculate_file(self._h, os.fsencode(os.fspath(feature_bin_path)), os.fsencode(os.fspath(out_path)),
                                         FDNN_DUMP_BIN if binary else FDNN_DUMP_TXT, chunk_frames, C.byref(done)))
        return int(done.value)

    def get_new_lazy_context(self, input_vector_count: int, batch_size: int = 8) -> "LazyContext":
        """getNewLazyContext(int[, int]) — QuantizedDnn.java:100-107."""
        h = C.c_void_p()
        _check(lib().fdnn_ctx_new(self._h, input_vector_count, batch_size, C.byref(h)))
        return LazyContext(self, h.value, input_vector_count)

    getNewLazyContext = get_new_lazy_context

    def delete(self) -> None:
        """delete() — QuantizedDnn.java:137-139.  Like the reference, not called automatically."""
        if self._h:
            lib().fdnn_free(self._h)
            self._h = None


class LazyContext:
    """Mirror of QuantizedDnn.LazyContext (QuantizedDnn.java:72-98): one context = one thread."""

    def __init__(self, dnn: QuantizedDnn, handle: int, input_vector_count: int):
        self.dnn = dnn
        self._h = C.c_void_p(handle)
        self.input_vector_count = input_vector_count
        self.current_vector_index = 0

    def calculate_until_output(self, input_) -> None:
        """calculateUntilOutput(float[][]) — :84-86; input must hold exactly inputVectorCount frames."""
        x = _frames(input_, self.dnn.input_dimension())
        if x.shape[0] != self.input_vector_count:
            raise ValueError(f"context was created for {self.input_vector_count} frames, got {x.shape[0]}")
        _check(lib().fdnn_ctx_until_output(self._h, _ptr(x)))
        self.current_vector_index = 0

    def calculate_for_output_nodes(self, active_nodes_mask) -> np.ndarray:
        """calculateForOutputNodes(byte[]) — :88-93; advances to the next frame on every call."""
        mask = np.ascontiguousarray(active_nodes_mask, dtype=np.int8)
        if mask.shape != (self.dnn.output_dimension(),):
            raise ValueError("mask length must equal the network's output dimension")
        out = np.empty(self.dnn.output_dimension(), dtype=np.float32)
        _check(lib().fdnn_ctx_lazy(self._h, self.current_vector_index, _ptr(mask), _ptr(out)))
        self.current_vector_index += 1
        return out

    calculateUntilOutput, calculateForOutputNodes = calculate_until_output, calculate_for_output_nodes

    def calculate_for_output_nodes_batch(self, masks) -> np.ndarray:
        """All frames of the context at once: masks [n][O] → [n][O] (BASELINE config 3)."""
        m = np.ascontiguousarray(masks, dtype=np.int8)
        if m.shape != (self.input_vector_count, self.dnn.output_dimension()):
            raise ValueError("masks must be [inputVectorCount][outputDimension]")
        out = np.empty(m.shape, dtype=np.float32)
        _check(lib().fdnn_ctx_lazy_batch(self._h, _ptr(m), _ptr(out)))
        return out

    # -- inspection (stage-level parity tests) -------------------------------------------------------
    def set_trace(self, enable: bool = True) -> None:
        _check(lib().fdnn_ctx_set_trace(self._h, int(enable)))

    def hidden(self, layer: int | None = None, n_frames: int | None = None) -> np.ndarray:
        n = self.input_vector_count if n_frames is None else n_frames
        if layer is None:
            layer = self.dnn.layer_count() - 2
        out = np.empty((n, self.dnn.hidden_dimension()), dtype=np.uint8)
        _check(lib().fdnn_ctx_hidden(self._h, layer, n, _ptr(out)))
        return out

    def hidden_digest(self, layer: int | None = None, n_frames: int | None = None) -> int:
        """64-bit position-weighted checksum of hidden(layer) computed on the device (soak tests)"""
        n = self.input_vector_count if n_frames is None else n_frames
        if layer is None:
            layer = self.dnn.layer_count() - 2
        d = C.c_ulonglong()
        _check(lib().fdnn_ctx_hidden_digest(self._h, layer, n, C.byref(d)))
        return int(d.value)

    def logits(self, n_frames: int | None = None) -> np.ndarray:
        n = self.input_vector_count if n_frames is None else n_frames
        out = np.empty((n, self.dnn.output_dimension()), dtype=np.float32)
        _check(lib().fdnn_ctx_logits(self._h, n, _ptr(out)))
        return out

    # -- device-resident entry points (pointers are integers, e.g. torch.Tensor.data_ptr()) ----------
    def forward_device(self, d_in: int, n_frames: int, d_out: int, stream: int = 0) -> None:
        _check(lib().fdnn_ctx_forward_device(self._h, C.c_void_p(d_in), n_frames, C.c_void_p(d_out), C.c_void_p(stream)))

    def until_output_device(self, d_in: int, n_frames: int, stream: int = 0) -> None:
        _check(lib().fdnn_ctx_until_output_device(self._h, C.c_void_p(d_in), n_frames, C.c_void_p(stream)))

    def lazy_batch_device(self, d_masks: int, n_frames: int, d_out: int, stream: int = 0) -> None:
        _check(lib().fdnn_ctx_lazy_batch_device(self._h, C.c_void_p(d_masks), n_frames, C.c_void_p(d_out), C.c_void_p(stream)))

    def input_undecided(self):
        """elements of the last pass the certified tensor-core input layer handed to the exact path (None: plain exact kernel)"""
        n = C.c_uint()
        _check(lib().fdnn_ctx_input_undecided(self._h, C.byref(n)))
        return None if n.value == 0xFFFFFFFF else int(n.value)

    def timeline(self, enable: bool):
        """arm (True) / read back (False) per-CTA phase stamps of the tensor-core layer kernels"""
        if enable:
            _check(lib().fdnn_ctx_timeline(self._h, 1, None))
            return None
        out = np.zeros((self.dnn.layer_count() - 1, 1024, 8), dtype=np.uint64)
        _check(lib().fdnn_ctx_timeline(self._h, 0, _ptr(out)))
        return out

    def profile_stages(self, d_in: int, n_frames: int, d_out: int, iters: int = 20) -> np.ndarray:
        """per-kernel milliseconds of one forward pass (bench aid): [input, int8 layers…, softmax]"""
        ms = np.zeros(self.dnn.layer_count() + 1, dtype=np.float32)
        _check(lib().fdnn_ctx_profile_stages(self._h, C.c_void_p(d_in), n_frames, C.c_void_p(d_out), iters, _ptr(ms)))
        return ms

    def profile_pass(self, d_in: int, n_frames: int, d_out: int, iters: int = 20):
        """(input-layer ms, rest-of-pass ms, fused?) of the pass as it normally runs (bench aid)"""
        ms = np.zeros(2, dtype=np.float32)
        fused = C.c_int()
        _check(lib().fdnn_ctx_profile_pass(self._h, C.c_void_p(d_in), n_frames, C.c_void_p(d_out), iters, _ptr(ms), C.byref(fused)))
        return float(ms[0]), float(ms[1]), bool(fused.value)

    def delete(self) -> None:
        """delete() — :95-97."""
        if self._h:
            lib().fdnn_ctx_free(self._h)
            self._h = None


def sigmoid_lut() -> np.ndarray:
    out = np.zeros(1280, dtype=np.uint8)
    _check(lib().fdnn_sigmoid_lut(_ptr(out)))
    return out


def launch_count() -> int:
    return int(lib().fdnn_launch_count())
