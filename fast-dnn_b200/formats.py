"""On-disk formats either side of the hot path (numpy, host only).

* ``dnn.bin`` network file — written by the reference's Java ``FeedForwardNetwork.saveBinary``
  (/root/reference/src/java/suskun/nn/FeedForwardNetwork.java:226-235, Layer.saveToStream :331-340)
  and read by the C++ ``FloatDnn`` constructor (/root/reference/src/cpp/float_dnn.cc:18-69).
  All words are 4-byte big-endian:  int32 layerCount; per layer int32 in, int32 out,
  fp32 W[out][in], fp32 bias[out];  then fp32 shift[in0], fp32 scale[in0].
* feature ``.bin`` — ``BatchData.serializeDataMatrix`` (BatchData.java:107-139) /
  ``BatchData(fileName)`` (float_dnn.cc:85-105): big-endian int32 frames, int32 dim, fp32 rows.
* output dump — ``BatchData::dumpToFile(…, binary=true)`` (float_dnn.cc:128-164): NATIVE-endian
  uint32 n, uint32 d, fp32 rows.

The product's loader for ``dnn.bin`` is C++ (csrc/model_host.cc); this module is the writer and
the python-side reader used by tools, tests and the bench to make synthetic networks.
"""
from __future__ import annotations

import numpy as np


def pad_to(num: int, div: int) -> int:
    """paddedSize() of float_dnn.cc:76-83."""
    dif = div - num % div
    return num if dif == div else num + dif


def write_dnn_bin(path, layers, shift, scale) -> None:
    """layers: sequence of (W[out][in] float32, bias[out] float32)."""
    with open(path, "wb") as f:
        f.write(np.array([len(layers)], dtype=">i4").tobytes())
        for w, b in layers:
            w = np.asarray(w, dtype=np.float32)
            b = np.asarray(b, dtype=np.float32)
            assert w.ndim == 2 and b.shape == (w.shape[0],)
            f.write(np.array([w.shape[1], w.shape[0]], dtype=">i4").tobytes())
            w.astype(">f4").tofile(f)
            b.astype(">f4").tofile(f)
        np.asarray(shift, dtype=np.float32).astype(">f4").tofile(f)
        np.asarray(scale, dtype=np.float32).astype(">f4").tofile(f)


def read_dnn_bin(path):
    """→ (layers [(W,b)…], shift, scale) as float32 arrays, UNPADDED (as stored)."""
    raw = np.fromfile(path, dtype=np.uint8)
    off = 0

    def take_i4(k):
        nonlocal off
        v = raw[off:off + 4 * k].view(">i4").astype(np.int64)
        off += 4 * k
        return v

    def take_f4(k):
        nonlocal off
        v = raw[off:off + 4 * k].view(">f4").astype(np.float32)
        off += 4 * k
        return v

    (count,) = take_i4(1)
    layers = []
    for _ in range(int(count)):
        n_in, n_out = (int(v) for v in take_i4(2))
        w = take_f4(n_in * n_out).reshape(n_out, n_in)
        b = take_f4(n_out)
        layers.append((w, b))
    in0 = layers[0][0].shape[1]
    shift = take_f4(in0)
    scale = take_f4(in0)
    return layers, shift, scale


def write_feature_bin(path, frames) -> None:
    frames = np.asarray(frames, dtype=np.float32)
    assert frames.ndim == 2
    with open(path, "wb") as f:
        f.write(np.array(frames.shape, dtype=">i4").tobytes())
        frames.astype(">f4").tofile(f)


def read_feature_bin(path) -> np.ndarray:
    """Reads exactly the header's frame count (the shipped data/16khz.bin holds one extra row,
    BatchData.java:126-138; the C++ loader ignores it, float_dnn.cc:88-102)."""
    raw = np.fromfile(path, dtype=np.uint8)
    n, d = (int(v) for v in raw[:8].view(">i4"))
    return raw[8:8 + 4 * n * d].view(">f4").astype(np.float32).reshape(n, d)


def read_output_dump(path) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8)
    n, d = (int(v) for v in raw[:8].view("<u4"))
    return raw[8:8 + 4 * n * d].view("<f4").reshape(n, d).copy()


def write_output_dump(path, rows) -> None:
    rows = np.ascontiguousarray(rows, dtype="<f4")
    with open(path, "wb") as f:
        f.write(np.array(rows.shape, dtype="<u4").tobytes())
        rows.tofile(f)


def align_network(layers, shift, scale, input_alignment: int = 4, hidden_alignment: int = 16):
    """numpy mirror of FeedForwardNetwork.align (FeedForwardNetwork.java:50-58, Layer.align :264-281):
    zero-pad the input width, every hidden width and the matching input widths; output width kept."""
    out = []
    for j, (w, b) in enumerate(layers):
        w = np.asarray(w, dtype=np.float32)
        b = np.asarray(b, dtype=np.float32)
        in_pad = pad_to(w.shape[1], input_alignment if j == 0 else hidden_alignment)
        out_pad = w.shape[0] if j == len(layers) - 1 else pad_to(w.shape[0], hidden_alignment)
        w2 = np.zeros((out_pad, in_pad), dtype=np.float32)
        w2[:w.shape[0], :w.shape[1]] = w
        b2 = np.zeros(out_pad, dtype=np.float32)
        b2[:b.shape[0]] = b
        out.append((w2, b2))
    in0 = pad_to(len(shift), input_alignment)
    sh = np.zeros(in0, dtype=np.float32)
    sc = np.zeros(in0, dtype=np.float32)
    sh[:len(shift)] = shift
    sc[:len(scale)] = scale
    return out, sh, sc
