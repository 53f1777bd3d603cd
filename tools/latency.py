#!/usr/bin/env python
"""Single-stream (lanes = 1) latency / throughput of the blocking mc_process call — the way the reference's
ProcessingChain would drive the core — for 1080p BGR, pageable and pinned host frames."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import lvm_b200 as L
    from lvm_b200.synth import synth_frame
    W, H = 1920, 1080
    out = {}
    for mode, ui in (("laplace", L.MagUiValues(L.MagnificationMode.Laplace, 20, 50.0, 0.4, 3.0, 0, 6, 30.0)),
                     ("phase", L.MagUiValues(L.MagnificationMode.Phase, 50, 50.0, 0.4, 3.0, 0, 6, 30.0)),
                     ("color", L.MagUiValues(L.MagnificationMode.Color, 100, 0.0, 0.8, 1.2, 0, 3, 30.0))):
        cfg = L.ProcessorConfig(magnification=L.toParams(ui))
        proc = L.MagnificationProcessor(0)
        frames = [synth_frame(t, W, H, 3) for t in range(4)]
        pin_in = [torch.from_numpy(f).pin_memory() for f in frames]
        pin_out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        for i in range(70 if mode == "color" else 6):
            proc.process_image(frames[i % 4], cfg)
        lat = []
        for i in range(40):
            t0 = time.perf_counter()
            proc.process_image(frames[i % 4], cfg)           # pageable numpy in, fresh numpy out
            lat.append(time.perf_counter() - t0)
        p = L.processor._to_mc(cfg)
        lat_p = []
        for i in range(40):
            t0 = time.perf_counter()
            proc.process_host(pin_in[i % 4].data_ptr(), W, H, 3, W * 3, p, pin_out.data_ptr(), W * 3)   # blocking mc_process
            lat_p.append(time.perf_counter() - t0)
        out[mode] = {"pageable_ms_median": float(np.median(lat) * 1e3), "pinned_ms_median": float(np.median(lat_p) * 1e3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
