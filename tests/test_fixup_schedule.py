"""CPU restatement of how csrc/input_tc.cu (input_fixup_block_kernel) walks the undecided elements of a block of 32 frames:
transposed bitmap words -> node list -> batches of 16 weight rows -> chunks of 8 (node, frame) elements -> consumer warps.
The GPU parity tests prove the kernel; this states the index arithmetic on its own and checks the invariants it relies on:
every set bit is visited exactly once, a chunk never leaves its batch, the rows of one chunk are consecutive slots (the
bank-conflict argument), and the producer's stand-in arrivals plus one arrival per chunk always add up to the barrier count."""
import numpy as np
import pytest

SLOTS, FRAMES, CONSUMERS, MAX_CHUNKS = 16, 32, 12, 16 * 32 // 8


def transpose32(words):
    """the five butterfly steps input_tc_kernel's epilogue does over a warp: lane i holds row i (bit c = column c)"""
    v = [int(w) for w in words]
    for j, m in ((16, 0x0000FFFF), (8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)):
        nv = list(v)
        for lane in range(32):
            p = v[lane ^ j]
            if lane & j == 0:
                nv[lane] = ((v[lane] & m) | ((p & m) << j)) & 0xFFFFFFFF
            else:
                nv[lane] = (v[lane] & ~m & 0xFFFFFFFF) | ((p & ~m & 0xFFFFFFFF) >> j)
        v = nv
    return v


def schedule(node_words):
    """node_words[n] = bits of the block's frames that left node n undecided -> list of (warp, batch, chunk, [(node, frame, slot)])"""
    listed = [(n, int(w)) for n, w in enumerate(node_words) if w]
    out, chunks_before = [], 0
    for b in range(0, len(listed), SLOTS):
        entries = listed[b:b + SLOTS]
        incl = np.cumsum([bin(w).count("1") for _, w in entries])
        total = int(incl[-1])
        n_chunks = (total + 7) // 8
        for j in range(n_chunks):
            elems = []
            for quad in range(8):
                e = 8 * j + quad
                if e >= total:
                    continue
                idx = int(np.sum(incl <= e))  # the entry whose elements include number e
                node, word = entries[idx]
                rank = e - (int(incl[idx]) - bin(word).count("1"))
                rest = word
                for _ in range(rank):
                    rest &= rest - 1
                frame = (rest & -rest).bit_length() - 1
                elems.append((node, frame, idx))
            out.append(((chunks_before + j) % CONSUMERS, b // SLOTS, j, elems))
        chunks_before += n_chunks
    return out


@pytest.mark.parametrize("density", [0.0, 0.005, 0.026, 0.3, 1.0])
def test_every_undecided_element_is_visited_once(density):
    rng = np.random.default_rng(int(density * 1000) + 1)
    bits = rng.random((FRAMES, 2048)) < density  # [frame][node]
    bits[7, :] |= density > 0  # a frame that certifies nothing (NaN row)
    frame_words = [[int(sum(int(bits[f, 32 * w + c]) << c for c in range(32))) for w in range(64)] for f in range(FRAMES)]
    # the kernel's transposed bitmap: per 32-node group, a 32 x 32 bit transpose over the warp
    node_words = []
    for w in range(64):
        node_words += transpose32([frame_words[f][w] for f in range(FRAMES)])
    for n in (0, 31, 32, 1000, 2047):
        assert node_words[n] == sum(int(bits[f, n]) << f for f in range(FRAMES))
    seen = np.zeros_like(bits)
    per_batch = {}
    for warp, batch, chunk, elems in schedule(node_words):
        assert 0 <= warp < CONSUMERS and 0 < len(elems) <= 8
        slots = sorted({s for _, _, s in elems})
        assert slots == list(range(slots[0], slots[-1] + 1)) and slots[-1] < SLOTS  # consecutive rows of one buffer
        per_batch[batch] = per_batch.get(batch, 0) + 1
        for node, frame, _ in elems:
            assert not seen[frame, node]
            seen[frame, node] = True
    assert np.array_equal(seen, bits)
    # "free again" barrier: one arrival per chunk + the producer's stand-ins = its count, never more
    assert all(0 < c <= MAX_CHUNKS for c in per_batch.values())
    assert len(per_batch) <= 2048 // SLOTS  # one "landed" barrier per batch
