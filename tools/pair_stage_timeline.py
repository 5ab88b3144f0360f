#!/usr/bin/env python
"""Where a pipeline stage of the CTA-pair layer kernel spends its round trip (profiling aid).

qlayer_pair.cu stamps, for the first 256 pipeline turns of its first pair, when the producer starts waiting for the slot (0), has
issued its TMA boxes (1), the MMA warp sees the stage full (2, leader), has issued MMAs + commit (3, leader), a scan set sees the
commit (4) and has released the stage (5).  SM clocks: only stamps of one CTA compare.

The stamps are compiled in only on request (the MMA-issuing thread pays for every instruction in its loop):
  touch fast-dnn_b200/csrc/qlayer_pair.cu && make -C fast-dnn_b200/csrc NVFLAGS_EXTRA=-DFDNN_STAGE_STAMPS
  python tools/pair_stage_timeline.py [frames] [layer]
  touch fast-dnn_b200/csrc/qlayer_pair.cu && make -C fast-dnn_b200/csrc        # back to the production kernel
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fast_dnn_b200 import quantized_dnn as qd, synth  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 2
STAGES = 6
dnn = qd.QuantizedDnn.load_from_file(synth.network_file("L"), device=0)
d_in = torch.from_numpy(synth.make_frames(m, 440, seed=7)).cuda()
d_out = torch.empty(m, 8000, dtype=torch.float32, device="cuda")
ctx = dnn.get_new_lazy_context(m)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.forward_device(d_in.data_ptr(), m, d_out.data_ptr(), s)
torch.cuda.synchronize()
ctx.timeline(True)
ctx.forward_device(d_in.data_ptr(), m, d_out.data_ptr(), s)
torch.cuda.synchronize()
tl = ctx.timeline(False).astype(np.int64).reshape(-1, 8192)[layer]
ev = tl[2048:2048 + 2 * 8 * 256].reshape(2, 8, 256)  # [cta][event][turn]


def med(a):
    a = a[np.isfinite(a)]
    return "      -" if a.size == 0 else f"{np.median(a):7.0f}"


def span(a, b):
    """b − a per turn where both were stamped"""
    ok = (a > 0) & (b > 0)
    return np.where(ok, (b - a).astype(np.float64), np.nan)


lo, hi = 32, 250  # steady state
for cta, name in ((0, "leader"), (1, "peer")):
    e = ev[cta]
    if not (e[1] > 0).any():
        print(f"{name}: no stamps (kernel of layer {layer} is not the CTA-pair kernel?)")
        continue
    nxt = np.full(256, 0, dtype=np.int64)
    nxt[:-STAGES] = e[0][STAGES:]  # the producer's wait for the SAME slot's next use
    per_stage = np.diff(e[1][lo:hi].astype(np.float64))
    per_stage = per_stage[per_stage > 0]
    print(f"{name}: cycles, median over pipeline turns {lo}..{hi} of layer {layer} ({m} frames)")
    print(f"  TMA issue → next TMA issue (the stage period)          {med(per_stage)}")
    print(f"  producer blocked on the slot (wait → issue)            {med(span(e[0], e[1])[lo:hi])}")
    if cta == 0:
        print(f"  TMA issue → MMA warp sees it full (both CTAs' bytes)   {med(span(e[1], e[2])[lo:hi])}")
        print(f"  full → 4 MMAs + commit issued                          {med(span(e[2], e[3])[lo:hi])}")
        print(f"  commit issued → scan set sees the commit (MMAs done)   {med(span(e[3], e[4])[lo:hi])}")
    else:
        print(f"  TMA issue → scan set sees the commit                   {med(span(e[1], e[4])[lo:hi])}")
    print(f"  scan: commit seen → stage released                      {med(span(e[4], e[5])[lo:hi])}")
    rt = np.full(256, np.nan)
    ok = (e[1][:-STAGES] > 0) & (e[1][STAGES:] > 0)
    rt[:-STAGES] = np.where(ok, (e[1][STAGES:] - e[1][:-STAGES]).astype(np.float64), np.nan)
    print(f"  slot round trip (issue → issue of the same slot)       {med(rt[lo:hi])}")
    rel_to_issue = np.full(256, np.nan)
    ok = (e[5][:-STAGES] > 0) & (e[1][STAGES:] > 0)
    rel_to_issue[:-STAGES] = np.where(ok, (e[1][STAGES:] - e[5][:-STAGES]).astype(np.float64), np.nan)
    print(f"  released → the slot's next TMA issue                   {med(rel_to_issue[lo:hi])}")
ctx.delete()
dnn.delete()
