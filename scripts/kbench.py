#!/usr/bin/env python
"""Kernel-level microbenchmark of the fused STS stage kernel through the C-ABI.

    python scripts/kbench.py [--n 16384] [--rows 8,16,32,64] [--iters 20]

Times b200_stencil_lincomb with the RKC/RKL stage pattern [L(x), v1, v2, x, v4] on
rotating buffers (so every launch streams from HBM, like the real stage loop), with
CUDA events on the launching stream.  Reports GB/s on the 40 B/cell algorithmic basis.
"""
import argparse
import ctypes
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b200 = importlib.import_module("ceda-demonstrations_b200")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--rows", default="32")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--pattern", default="stage", help="stage | rhs | final | chainK; comma list = sweep in one process")
    ap.add_argument("--variant", default="1", help="chain kernel: 0 = two cells / thread, 1 = four (comma list)")
    ap.add_argument("--arith", default="exact", help="exact | fma (comma list)")
    ap.add_argument("--uniform", default="0", help="0 = random coefficient tables, 1 = uniform coefficients flagged in the geometry (comma list)")
    args = ap.parse_args()
    combos = [(p_, int(v_), a_, int(u_)) for p_ in args.pattern.split(",") for v_ in args.variant.split(",")
              for a_ in args.arith.split(",") for u_ in args.uniform.split(",")]
    nx = args.n
    ny = args.ny or args.n
    ctx = b200.Context(0)
    lib = b200.kernel_lib()
    dev = torch.device("cuda", 0)
    N = nx * ny
    bufs = [torch.rand(N, dtype=torch.float64, device=dev) for _ in range(6)]
    cx = [torch.rand(nx, dtype=torch.float64, device=dev) + 1.0 for _ in range(2)]
    cy = [torch.rand(ny, dtype=torch.float64, device=dev) + 1.0 for _ in range(2)]
    g_rand = b200.StencilGeom(nx, ny, cx[0].data_ptr(), cx[1].data_ptr(), cy[0].data_ptr(), cy[1].data_ptr(), None, None, None, None)
    ux = [torch.full((nx,), 1.25, dtype=torch.float64, device=dev) for _ in range(2)]
    uy = [torch.full((ny,), 1.75, dtype=torch.float64, device=dev) for _ in range(2)]
    g_uni = b200.StencilGeom(nx, ny, ux[0].data_ptr(), ux[1].data_ptr(), uy[0].data_ptr(), uy[1].data_ptr(), None, None, None, None,
                             1, 1.25, 1.25, 1.75, 1.75)
    coeffs = [1e-7, -0.3, 0.2, 1.1, -2e-8]
    out = {}
    for pattern, variant, arith, uniform, rows in [(p_, v_, a_, u_, int(r)) for (p_, v_, a_, u_) in combos for r in args.rows.split(",")]:
        if not pattern.startswith("chain") and (variant, arith, uniform) != combos[0][1:]:
            continue  # variant / arith / uniform only concern the chain kernels
        args.pattern, args.variant, args.arith = pattern, variant, arith
        g = g_uni if uniform else g_rand
        lib.b200_set_chain_variant(variant)
        lib.b200_set_contract(1 if arith == "fma" else 0)
        lib.b200_set_rows_per_block(rows)
        if args.pattern.startswith("chain"):
            lib.b200_set_chain_rows(rows)

        def launch(k):
            x, v1, yn, fn, z = bufs[k % 3], bufs[(k + 1) % 3], bufs[3], bufs[4], bufs[(k + 2) % 3]
            if args.pattern == "stage":
                ctx.stencil_lincomb(g, x, coeffs, [2, 0, 0, 1, 0], [None, v1, yn, None, fn], z)
                return 40.0
            if args.pattern == "rhs":
                ctx.stencil_lincomb(g, x, [1.0], [2], [None], z)
                return 16.0
            if args.pattern == "final":
                ex = b200.StageExtras(bufs[5].data_ptr(), None, None, None, None, None, None)
                ctx.stencil_lincomb(g, x, [0.8, -0.8, 0.4e-4, 0.4e-4], [0, 1, 0, 2], [yn, None, fn, None], z, ex)
                return 40.0
            if args.pattern.startswith("chain"):
                kk = int(args.pattern[5:])
                cs = [[1e-7 * (l + 1), -0.3, 0.2, 1.1, -2e-8] for l in range(kk)]
                outs = [None] * kk
                outs[kk - 1] = z
                outs[kk - 2] = bufs[5]
                ctx.stencil_chain(g, x, v1, yn, fn, cs, outs)
                return 40.0 * kk
            raise SystemExit("unknown pattern")

        for k in range(5):
            launch(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(args.iters):
            bpc = launch(k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.iters
        gbs = bpc * N / (ms * 1e-3) / 1e9
        out[(pattern, variant, arith, uniform, rows)] = {"ms": ms, "GBs": gbs}
        kk = int(args.pattern[5:]) if args.pattern.startswith("chain") else 1
        print("n=%dx%d pattern=%s variant=%d arith=%s uniform=%d rows=%d: %.3f ms/launch  %.1f GB/s (%.0f B/cell basis)  %.3e cell-updates/s"
              % (nx, ny, args.pattern, args.variant, args.arith, uniform, rows, ms, gbs, bpc, kk * N / (ms * 1e-3)))
    return out


if __name__ == "__main__":
    main()
