"""Decoder for the relocatable model blob produced by ``fdnn_pack`` (csrc/fdnn_internal.h:
BlobHeader / BlobQLayer / FixEntry).  Host-side inspection only: tools and tests use it to look
at exactly what is uploaded to (and NCCL-broadcast between) the GPUs."""
from __future__ import annotations

import numpy as np

MAGIC = 0x424E4446
VERSION = 6
FIX_GROUPS = (64, 128, 256)
FIX_KBLOCK = 128

HEADER = np.dtype([
    ("magic", "<u4"), ("version", "<u4"), ("total_size", "<u8"),
    ("in_dim", "<i4"), ("in_dim_file", "<i4"), ("hidden", "<i4"), ("out_dim", "<i4"),
    ("n_qlayers", "<i4"), ("cutoff", "<f4"),
    ("off_w0", "<u8"), ("off_bias0", "<u8"), ("off_shift", "<u8"), ("off_scale", "<u8"),
    ("off_lut", "<u8"), ("off_qlayers", "<u8"),
])
QLAYER = np.dtype([
    ("nodes", "<i4"), ("inputs", "<i4"), ("multiplier", "<f4"), ("coeff", "<f4"), ("rcp_coeff", "<f4"),
    ("n_fix", "<u4"), ("fast_div", "<u4"), ("k_blocks", "<u4"),
    ("off_w", "<u8"), ("off_bias", "<u8"), ("off_fix_ptr", "<u8", (3,)), ("off_fix_ent", "<u8", (3,)),
])
assert HEADER.itemsize == 88 and QLAYER.itemsize == 96


class Blob:
    def __init__(self, data: np.ndarray):
        self.data = np.ascontiguousarray(data, dtype=np.uint8)
        self.header = self.data[:HEADER.itemsize].view(HEADER)[0]
        if int(self.header["magic"]) != MAGIC or int(self.header["version"]) != VERSION:
            raise ValueError("not a fast-dnn model blob")
        o = int(self.header["off_qlayers"])
        self.qlayers = self.data[o:o + QLAYER.itemsize * int(self.header["n_qlayers"])].view(QLAYER)

    def _f32(self, off, count):
        return self.data[int(off):int(off) + 4 * count].view("<f4")

    @property
    def in_dim(self):
        return int(self.header["in_dim"])

    @property
    def hidden(self):
        return int(self.header["hidden"])

    @property
    def out_dim(self):
        return int(self.header["out_dim"])

    def input_layer(self):
        H, I = self.hidden, self.in_dim
        return (self._f32(self.header["off_w0"], H * I).reshape(H, I), self._f32(self.header["off_bias0"], H),
                self._f32(self.header["off_shift"], I), self._f32(self.header["off_scale"], I))

    def lut2(self):
        """doubled sigmoid table: index trunc(2*clamp(x*100, -641, 641)) + 1282"""
        o = int(self.header["off_lut"])
        return self.data[o:o + 2565]

    def qlayer(self, i):
        q = self.qlayers[i]
        n, k = int(q["nodes"]), int(q["inputs"])
        w = self.data[int(q["off_w"]):int(q["off_w"]) + n * k].view(np.int8).reshape(n, k)
        return w, self._f32(q["off_bias"], n), float(q["multiplier"])

    def fix_list(self, i, variant=0):
        """→ (ptr uint32 [groups*k_blocks+1], pair uint32 [n_fix], w0 int8, w1 int8, node uint32) of the
        risk list grouped by FIX_GROUPS[variant] nodes; ordered by (node // G, pair // 64, node, pair)."""
        q = self.qlayers[i]
        G = FIX_GROUPS[variant]
        nc, nf = -(-int(q["nodes"]) // G) * int(q["k_blocks"]), int(q["n_fix"])
        o_ptr, o_ent = int(q["off_fix_ptr"][variant]), int(q["off_fix_ent"][variant])
        ptr = self.data[o_ptr:o_ptr + 4 * (nc + 1)].view("<u4")
        ent = self.data[o_ent:o_ent + 8 * nf].view("<u4").reshape(nf, 2)
        pair = ent[:, 0] & 0xFFFF
        w0 = ((ent[:, 0] >> 16) & 0xFF).astype(np.uint8).view(np.int8)
        w1 = ((ent[:, 0] >> 24) & 0xFF).astype(np.uint8).view(np.int8)
        return ptr, pair, w0, w1, ent[:, 1]
