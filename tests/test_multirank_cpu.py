"""N>1 host logic on CPU: world_size-2/4 `gloo` process groups exercise the 2-D block
decomposition (ceda-demonstrations_b200.block_decomposition, the Python mirror of
UserData::setup) and the halo-exchange pairing of start_exchange/end_exchange
(diffusion_2D.cpp:400-584: what a rank sends west is its west neighbour's east halo, ...),
with the oracle's laplacian as the per-rank compute.  The assembled multi-rank result must equal
the single-rank periodic result exactly; a WRMS norm reduced over ranks must equal the global one
up to summation order."""
import ctypes
import importlib
import math
import os
import socket
import sys

import numpy as np
import pytest
from conftest import ROOT, OrcGrid, P, make_grid

NX, NY = 37, 26


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b200 = importlib.import_module("ceda-demonstrations_b200")
    orc = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle_sts.so"))
    orc.orc_wsqrsum.restype = ctypes.c_double

    d = b200.block_decomposition(NX, NY, rank, world)
    nxl, nyl = d["nx_loc"], d["ny_loc"]
    g = make_grid(NX, NY, kx=0.8, ky=1.9, inhom=True, nx_loc=nxl, ny_loc=nyl, is_=d["is"], js=d["js"])
    # the same global field on every rank, cut to the local block
    U = np.random.default_rng(5).standard_normal((NY, NX))
    u = np.ascontiguousarray(U[d["js"] : d["js"] + nyl, d["is"] : d["is"] + nxl]).ravel()
    Ws, Es, Ss, Ns = np.zeros(nyl), np.zeros(nyl), np.zeros(nxl), np.zeros(nxl)
    orc.orc_pack(ctypes.byref(g), P(u), P(Ws), P(Es), P(Ss), P(Ns))

    def exchange(send, dst, src, n):
        """send `send` to rank dst, receive n doubles from rank src (self-sends are local copies)."""
        if dst == rank and src == rank:
            return send.copy()
        recv = torch.zeros(n, dtype=torch.float64)
        reqs = [dist.isend(torch.from_numpy(send.copy()), dst), dist.irecv(recv, src)]
        for r in reqs:
            r.wait()
        return recv.numpy().copy()

    # west edge -> west neighbour (arrives as its E halo); my E halo comes from my east neighbour
    Er = exchange(Ws, d["ipW"], d["ipE"], nyl)
    Wr = exchange(Es, d["ipE"], d["ipW"], nyl)
    Nr = exchange(Ss, d["ipS"], d["ipN"], nxl)
    Sr = exchange(Ns, d["ipN"], d["ipS"], nxl)
    f = np.zeros(nxl * nyl)
    orc.orc_laplacian(ctypes.byref(g), P(u), P(f), P(Wr), P(Er), P(Sr), P(Nr))

    # global WRMS-style reduction: local sum of squares, all-reduce (nvector_parallel.c:721-730)
    w = np.full(nxl * nyl, 0.5)
    loc = torch.tensor([orc.orc_wsqrsum(P(f), P(w), ctypes.c_int64(f.size))], dtype=torch.float64)
    dist.all_reduce(loc)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), f=f.reshape(nyl, nxl), is_=d["is"], js=d["js"], wsum=loc.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_halo_exchange_matches_single_rank(orc, tmp_path, world):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = make_grid(NX, NY, kx=0.8, ky=1.9, inhom=True)
    U = np.random.default_rng(5).standard_normal((NY, NX))
    F = np.zeros(NX * NY)
    orc.orc_laplacian(ctypes.byref(g), P(np.ascontiguousarray(U).ravel()), P(F), None, None, None, None)
    F = F.reshape(NY, NX)
    got = np.full((NY, NX), np.nan)
    wsum = None
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        f = z["f"]
        got[int(z["js"]) : int(z["js"]) + f.shape[0], int(z["is_"]) : int(z["is_"]) + f.shape[1]] = f
        wsum = float(z["wsum"][0])
    assert np.array_equal(got, F)
    w = np.full(NX * NY, 0.5)
    want = orc.orc_wsqrsum(P(np.ascontiguousarray(F).ravel()), P(w), ctypes.c_int64(NX * NY))
    assert wsum == pytest.approx(want, rel=1e-13)
