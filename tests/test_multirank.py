"""The N > 1 plumbing on CPU: world_size-2 gloo job that packs on rank 0, broadcasts the blob once,
and shards a frame stream with no data-path collective."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_dnn_b200 import quantized_dnn as qd
from fast_dnn_b200 import sharding, synth


def test_shard_ranges_partition_the_stream():
    for n in (0, 1, 7, 512, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert list(sharding.chunk_ranges(10, 31, 8)) == [(10, 18), (18, 26), (26, 31)]
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_device_group_shard_plan_partitions_a_call():
    """fdnn_shard_plan (what fdnn_calculate does on a device group, host-only): contiguous, ordered, whole 128-frame tiles, balanced
    to one tile, nobody used for less than a tile, ragged end on the last used device"""
    for n in (0, 1, 127, 128, 129, 512, 1000, 4096, 9001, 1_000_000):
        for devs in (1, 2, 3, 8):
            plan, used = qd.shard_plan(n, devs)
            assert 1 <= used <= devs and len(plan) == devs
            pos = 0
            for d, (first, count) in enumerate(plan):
                assert first == pos and count >= 0
                assert (count > 0) == (d < used) or n == 0
                if d + 1 < used:
                    assert count % 128 == 0 and count > 0
                pos += count
            assert pos == n
            tiles = [(c + 127) // 128 for _, c in plan[:used]]
            assert max(tiles) - min(tiles) <= 1
    assert qd.shard_plan(512, 8)[1] == 4 and qd.shard_plan(100, 8)[1] == 1
    with pytest.raises(ValueError):
        qd.shard_plan(-1, 2)


def _free_port():
    """a port below the kernel's ephemeral range (an ephemeral one can be handed to somebody's outgoing connection between this
    probe and the rendezvous binding it — seen once on a GPU box: EADDRINUSE)"""
    import random
    for _ in range(200):
        port = random.randint(15000, 29999)
        with socket.socket() as s:
            try:
                s.bind(("127.0.0.1", port))
                return port
            except OSError:
                continue
    raise RuntimeError("no free port found")


def _worker(rank, world, port, path, n_frames, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        blob = sharding.broadcast_blob(qd.pack(path) if rank == 0 else None, src=0)
        digest = hashlib.sha256(blob.numpy().tobytes()).hexdigest()
        begin, end = sharding.shard_range(n_frames, rank, world)
        counts = sharding.gather_counts(end - begin)
        # each rank regenerates exactly its shard of the seeded stream, chunk by chunk
        rows = [synth.make_frames(hi - lo, 12, seed=7, start=lo) for lo, hi in sharding.chunk_ranges(begin, end, 16)]
        shard = np.concatenate(rows) if rows else np.zeros((0, 12), np.float32)
        np.save(os.path.join(result_dir, f"rank{rank}.npy"), shard)
        with open(os.path.join(result_dir, f"rank{rank}.txt"), "w") as f:
            f.write(f"{digest} {begin} {end} {sum(counts)} {len(blob)}")
    finally:
        dist.destroy_process_group()


def test_two_ranks_broadcast_blob_and_shard_frames(tmp_path, net_file):
    path = net_file("tiny")
    n_frames, world = 101, 2
    mp.spawn(_worker, args=(world, _free_port(), path, n_frames, str(tmp_path)), nprocs=world, join=True)
    want = hashlib.sha256(qd.pack(path).tobytes()).hexdigest()
    spans, shards = [], []
    for r in range(world):
        digest, begin, end, total, size = open(tmp_path / f"rank{r}.txt").read().split()
        assert digest == want, "every rank must hold rank 0's packed model bytes"
        assert int(total) == n_frames
        spans.append((int(begin), int(end)))
        shards.append(np.load(tmp_path / f"rank{r}.npy"))
    assert spans == [(0, 51), (51, 101)]
    assert np.array_equal(np.concatenate(shards), synth.make_frames(n_frames, 12, seed=7))
