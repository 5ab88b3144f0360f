#!/usr/bin/env python
"""calculate() on PAGEABLE host arrays: page-locked staging inside the library (default) against the driver's own pageable copies
(FDNN_STAGE=0), per caller-thread count.  python tools/pageable_e2e.py"""
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np
    import fast_dnn_b200  # noqa: F401
    from fast_dnn_b200 import quantized_dnn as qd, synth
    threads, n = int(sys.argv[2]), int(sys.argv[3])
    dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
    x = [synth.make_frames(n, 440, seed=3 + t) for t in range(threads)]
    y = [np.zeros((n, 8000), np.float32) for _ in range(threads)]  # touched once, reused: like a recycled Java heap region

    def work(t, iters):
        for _ in range(iters):
            dnn.calculate(x[t], out=y[t])

    for t in range(threads):
        work(t, 3)
    iters = max(4, 40 * 512 // n)
    ws = [threading.Thread(target=work, args=(t, iters)) for t in range(threads)]
    t0 = time.perf_counter()
    [w.start() for w in ws]
    [w.join() for w in ws]
    dt = time.perf_counter() - t0
    print(f"{threads * iters * n / dt / 1e3:.0f}")
    sys.exit(0)

print("k frames/s through calculate() on pageable arrays; columns: caller threads 1 2 4 6")
for n in (512, 4096):
    for name, env in (("library staging", {}), ("driver pageable copies (FDNN_STAGE=0)", {"FDNN_STAGE": "0"})):
        row = []
        for threads in (1, 2, 4, 6):
            out = subprocess.run([sys.executable, __file__, "--one", str(threads), str(n)], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
            row.append(out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else "fail")
        print(f"n={n:5d} {name:40s} " + "  ".join(f"{v:>6s}" for v in row), flush=True)
