#!/usr/bin/env python
"""us per 512-frame step with several contexts in flight, for the fused kernel at reduced grid sizes (FDNN_FUSED_TGRID) against
the layer-by-layer kernels (FDNN_FUSED=0), tile policy 'throughput'.  python tools/fused_sweep.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch
    import fast_dnn_b200  # noqa: F401
    from fast_dnn_b200 import quantized_dnn as qd, synth
    lanes = int(sys.argv[2])
    dev = torch.device("cuda", 0)
    dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
    dnn.set_tile_policy("throughput")
    pool = 16
    d_in = [torch.from_numpy(synth.make_frames(512, 440, seed=1000 + i)).to(dev) for i in range(pool)]
    d_out = [torch.empty(512, 8000, dtype=torch.float32, device=dev) for _ in range(pool)]
    ctxs = [dnn.get_new_lazy_context(512) for _ in range(lanes)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]

    def step(i):
        k = i % lanes
        ctxs[k].forward_device(d_in[i % pool].data_ptr(), 512, d_out[i % pool].data_ptr(), streams[k].cuda_stream)

    for i in range(3 * pool * lanes):
        step(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(lanes)]
        e0.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_event(e0)
        for i in range(400):
            step(i)
        for s, e in zip(streams, e1):
            e.record(s)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e) for e in e1) / 400 * 1e3)
    print(f"{best:.1f}")
    sys.exit(0)

print("us per step (512 frames), tile policy throughput; rows: configuration, columns: contexts in flight " + os.environ.get("SWEEP_LANES", "1,2,3,4,6"))
for name, env in [("layer-by-layer", {"FDNN_FUSED": "0"})] + [(f"fused grid {g}", {"FDNN_FUSED": "2", "FDNN_FUSED_TGRID": str(g)}) for g in [int(v) for v in os.environ.get("SWEEP_GRIDS", "48,64,74,148").split(",")]]:
    row = []
    for lanes in [int(v) for v in os.environ.get("SWEEP_LANES", "1,2,3,4,6").split(",")]:
        out = subprocess.run([sys.executable, __file__, "--one", str(lanes)], capture_output=True, text=True, env=dict(os.environ, **env), timeout=300)
        row.append(out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else "fail")
    print(f"{name:18s} " + "  ".join(f"{v:>7s}" for v in row), flush=True)
