#!/usr/bin/env python
"""Diagnostics of the pipelined host<->device path: raw PCIe copy rates (alone / both directions /
under a running chain kernel) and the B200_PIPE_TRACE timeline of b200_d2d_run_batches."""
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["B200_PIPE_TRACE"] = "1"
b200 = importlib.import_module("ceda-demonstrations_b200")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    N = n * n
    dev = torch.device("cuda", 0)
    h = [torch.empty(N, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    for t in h:
        t.fill_(0.5)
    d = [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(2)]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn, reps=3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def h2d():
        with torch.cuda.stream(s1):
            d[0].copy_(h[0], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h[1].copy_(d[1], non_blocking=True)

    def both():
        h2d()
        d2h()

    gb = 8 * N / 1e9
    for name, fn in (("h2d alone", h2d), ("d2h alone", d2h), ("h2d + d2h concurrently", both)):
        t = timed(fn)
        print("%-26s %.1f ms  (%.1f GB/s per direction)" % (name, 1e3 * t, gb / t))
    del d
    wl = ["--nx", str(n), "--ny", str(n), "--integrator", "rkc", "--fixedstep", "1e-4", "--tf", "1.0", "--nout", "1", "--output", "0"]
    prob = b200.Diffusion2D(wl, device=0)
    prob.step(3)
    torch.cuda.synchronize()
    t = timed(lambda: prob.step(1), reps=4)
    print("step resident               %.1f ms" % (1e3 * t))
    prob.get_state(h[0])
    h[1].copy_(h[0])
    t_cur = prob.stats()["t"]
    prob.run_batches([h[0], h[1]], [h[2], h[3]], t_cur, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prob.run_batches([h[i % 2] for i in range(nb)], [h[2 + i % 2] for i in range(nb)], t_cur, 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("run_batches: %d batches %.1f ms total, %.1f ms / batch" % (nb, 1e3 * dt, 1e3 * dt / nb))
    prob.close()


if __name__ == "__main__":
    main()
