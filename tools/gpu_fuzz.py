"""GPU fuzz (needs a B200; numpy + ctypes only, no torch): random small aligned networks — including depths of only two int8
layers, which the reference itself cannot run (dnn.cc:199) — with adversarial weights and frames through the C ABI against the
plain-C restatement: last-hidden bytes and logits bit for bit, scores within the stated tolerance, NaN rows in the same places.
Prints one JSON line; stops after --seconds of wall time.    python tools/gpu_fuzz.py --seed 1 --seconds 20"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fast_dnn_b200  # noqa: E402,F401
from fast_dnn_b200 import formats, synth  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd  # noqa: E402
import oracle_py  # noqa: E402  (checker)


def random_network(rng):
    I = int(rng.choice([4, 8, 12, 20, 128]))
    H = int(rng.choice([16, 32, 48, 128, 256]))
    O = int(rng.integers(1, 300))
    dims = [I] + [H] * int(rng.integers(2, 6)) + [O]
    layers = []
    for j in range(len(dims) - 1):
        k, n = dims[j], dims[j + 1]
        mode = int(rng.integers(0, 8))
        sigma = 10.0 ** rng.uniform(-3, 1.0)
        w = rng.normal(0, sigma, (n, k)).astype(np.float32)
        if mode == 1:
            w[rng.random((n, k)) < 0.05] *= 50
        elif mode == 2:
            w[:] = 0
        elif mode == 3:
            w = np.round(w * 4) / 4
        elif mode == 4:
            w[rng.random((n, k)) < 0.1] = np.float32(3.0) * rng.choice([-1, 1])
        elif mode == 5:
            w = rng.integers(-300, 300, (n, k)) / np.float32(rng.choice([1, 2, 7, 63.5, 127]))
        elif mode == 6 and j > 0:
            w[0, 0] = np.float32(rng.choice([1e30, -1e30]))
        bias = rng.normal(0, 10.0 ** rng.uniform(-2, 1.2), n).astype(np.float32)
        layers.append((np.asarray(w, dtype=np.float32), bias))
    shift = rng.normal(0, 0.1, I).astype(np.float32)
    scale = rng.uniform(0.05, 0.1, I).astype(np.float32)
    return dims, layers, shift, scale


def run(seed: int = 1, seconds: float = 20.0, max_networks: int = 1 << 30) -> dict:
    rng = np.random.default_rng(seed)
    t0 = time.time()
    stats = {"networks": 0, "frames": 0, "hidden_mismatch": [], "logits_mismatch": [], "nan_pattern_mismatch": [], "score_tolerance": [], "errors": []}
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "net.bin")
        while time.time() - t0 < seconds and stats["networks"] + len(stats["errors"]) < max_networks:
            dims, layers, shift, scale = random_network(rng)
            cutoff = float(rng.choice([3.0, 3.0, 0.5, 1.0, 10.0]))
            formats.write_dnn_bin(path, layers, shift, scale)
            n = int(rng.choice([1, 7, 33, 130, 300]))
            if rng.random() < 0.3:
                frames = synth.make_hostile_frames(n, dims[0], seed=int(rng.integers(1, 1000)))
            else:
                frames = rng.normal(0, 15, (n, dims[0])).astype(np.float32)
            tag = {"dims": dims, "cutoff": cutoff, "n": n}
            try:
                port = oracle_py.Port(path, cutoff)
                dnn = qd.QuantizedDnn.load_from_file(path, cutoff, device=0)
                got = dnn.calculate(frames.copy())
                ctx = dnn.get_new_lazy_context(n)
                ctx.calculate_until_output(frames.copy())
                hid, logits = ctx.hidden(), ctx.logits()
                ctx.delete()
                dnn.delete()
            except Exception as e:  # noqa: BLE001
                stats["errors"].append(dict(tag, error=repr(e)[:200]))
                continue
            want_hid = port.until_output(frames.copy())
            want_logits = (port.output_linear(want_hid) + port.qlayer(port.qlayer_count - 1)[1]).astype(np.float32)
            want = port.calculate(frames.copy())
            port.close()
            stats["networks"] += 1
            stats["frames"] += n
            if not np.array_equal(hid, want_hid):
                stats["hidden_mismatch"].append(dict(tag, bytes=int((hid != want_hid).sum())))
                continue
            same = (logits.view(np.uint32) == want_logits.view(np.uint32)) | (np.isnan(logits) & np.isnan(want_logits))
            if not same.all():
                stats["logits_mismatch"].append(dict(tag, elements=int((~same).sum())))
            if not np.array_equal(np.isnan(got), np.isnan(want)):
                stats["nan_pattern_mismatch"].append(dict(tag, got=int(np.isnan(got).sum()), want=int(np.isnan(want).sum())))
                continue
            ok = ~np.isnan(want)
            bad = np.abs(got[ok] - want[ok]) > 1e-9 + 2e-5 * np.abs(want[ok])
            if bad.any():
                stats["score_tolerance"].append(dict(tag, elements=int(bad.sum()), worst=float(np.abs(got[ok] - want[ok])[bad].max())))
    stats["seconds"] = round(time.time() - t0, 1)
    for k in ("hidden_mismatch", "logits_mismatch", "nan_pattern_mismatch", "score_tolerance", "errors"):
        stats[k + "_count"] = len(stats[k])
        stats[k] = stats[k][:6]
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=20.0)
    args = ap.parse_args()
    print(json.dumps(run(args.seed, args.seconds)))


if __name__ == "__main__":
    main()
