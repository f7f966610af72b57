#!/usr/bin/env python3
"""Experiment: the resident 256-frame step issued as N independent sub-batches (own handles, own streams), so that the latency-bound
matcher tail of one sub-batch overlaps the extraction of the next.  usage: slots_probe.py [steps]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from gtsam_vslam_b200 import synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
args = argparse.Namespace(no_temporal=False)
cfg = synth.CONFIGS["euroc"]
F = 256
for n in (1, 2, 4):
    arms = [bench.Arm(args, cfg, F // n, 0, seed=1000 + i) for i in range(n)]
    master = torch.cuda.Stream()
    for _ in range(3):
        for a in arms:
            a.step_resident()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(master)
    for a in arms:
        a.stream.wait_event(e0)
    for _ in range(steps):
        for a in arms:
            a.step_resident()
    for a in arms:
        ev = torch.cuda.Event()
        ev.record(a.stream)
        master.wait_event(ev)
    e1.record(master)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"sub-batches {n}: {ms:.3f} ms per 256-frame step -> {F / ms * 1e3:.0f} stereo frames/s", flush=True)
    del arms
    torch.cuda.empty_cache()
