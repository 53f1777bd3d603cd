#!/usr/bin/env python
"""Device-resident timing of the other BASELINE.json configs (parity-test cases, not bench lines):
   python tools/mode_bench.py [phase|color|laplace_gray|laplace4k] [--lanes N] [--steps K]
Prints frames/s plus the per-kernel device-time table (profile_kernels)."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import lvm_b200 as L
    from lvm_b200 import capi
    from lvm_b200.synth import synth_frame
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["phase", "color", "color6", "laplace_gray", "laplace4k", "laplace"])
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    W, H, CH = 1920, 1080, 3
    p = capi.McParams()
    lib = capi.lib()
    if args.mode == "phase":
        lib.mc_params_from_ui(C.byref(p), capi.MODE_PHASE, 50, 50.0, 0.4, 3.0, 0, 6, 30.0)
    elif args.mode in ("color", "color6"):
        lib.mc_params_from_ui(C.byref(p), capi.MODE_COLOR, 100, 0.0, 0.8, 1.2, 0, 3 if args.mode == "color" else 6, 30.0)
    elif args.mode == "laplace_gray":
        CH = 1
        lib.mc_params_from_ui(C.byref(p), capi.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 0, 6, 30.0)
    elif args.mode == "laplace4k":
        W, H = 3840, 2160
        lib.mc_params_from_ui(C.byref(p), capi.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 0, 8, 30.0)
    else:
        lib.mc_params_from_ui(C.byref(p), capi.MODE_LAPLACE, 20, 50.0, 0.4, 3.0, 0, 6, 30.0)
    lanes, T = args.lanes, 4
    base = [synth_frame(t, W, H, CH) for t in range(T)]
    clip = np.stack([np.stack([np.roll(base[t], (11 * k, 37 * k), axis=(0, 1)) for k in range(lanes)]) for t in range(T)])
    clip_d = torch.from_numpy(clip).cuda()
    out_d = torch.empty_like(clip_d[0])
    row = W * CH
    proc = L.MagnificationProcessor(0, lanes=lanes)
    stream = torch.cuda.ExternalStream(proc.stream)
    warm = 70 if args.mode.startswith("color") else 4   # Color: fill the 64-frame window first
    for i in range(warm):
        proc.process_device(clip_d[i % T].data_ptr(), W, H, CH, row, p, out_d.data_ptr(), row)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        proc.process_device(clip_d[i % T].data_ptr(), W, H, CH, row, p, out_d.data_ptr(), row)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    proc.set_option("profile_kernels", 1)
    for i in range(10):
        proc.process_device(clip_d[i % T].data_ptr(), W, H, CH, row, p, out_d.data_ptr(), row)
    prof = proc.profile_read()
    table = sorted(((k, lvl, n, tms / n * 1e3) for (k, lvl), (n, tms) in prof.items()), key=lambda r: -r[3] * r[2])
    print(json.dumps({"mode": args.mode, "w": W, "h": H, "c": CH, "lanes": lanes, "fps": lanes * args.steps / (ms * 1e-3),
                      "ms_per_frame": ms / args.steps / lanes,
                      "kernels": [{"kernel": f"{k}[{lvl}]", "us": round(us, 1)} for k, lvl, n, us in table]}))


if __name__ == "__main__":
    main()
