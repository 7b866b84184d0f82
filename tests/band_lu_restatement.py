"""Test infrastructure: numpy restatement of SUNDIALS' banded LU (SUN/src/sundials/sundials_band.c, bandGBTRF :168-280 and
bandGBTRS :282-323) on a dense array -- same pivot rule (strictly larger magnitude wins), same operation order, separate
multiply and add (numpy float64 has no contraction).  Used to check the per-block device kernels of csrc/react_kernels.cuh:
applied to the WHOLE (2 npts) x (2 npts) matrix with ml = mu = 2 it never leaves the 2 x 2 blocks, which is the claim the
kernels rest on."""
import numpy as np


def band_gbtrf(a, ml, smu):
    """In-place LU of the dense square array `a` treated as a band matrix (entries outside the band are ignored, fill-in
    up to smu super-diagonals).  Returns (pivots, info) like bandGBTRF."""
    n = a.shape[0]
    p = np.zeros(n, dtype=np.int64)
    for k in range(n - 1):
        last_row = min(n - 1, k + ml)
        l, mx = k, abs(a[k, k])
        for i in range(k + 1, last_row + 1):
            if abs(a[i, k]) > mx:
                l, mx = i, abs(a[i, k])
        p[k] = l
        if a[l, k] == 0.0:
            return p, k + 1
        swap = l != k
        if swap:
            a[l, k], a[k, k] = a[k, k], a[l, k]
        mult = -1.0 / a[k, k]
        for i in range(k + 1, last_row + 1):
            a[i, k] *= mult
        last_col = min(k + smu, n - 1)
        for j in range(k + 1, last_col + 1):
            a_kj = a[l, j]
            if swap:
                a[l, j] = a[k, j]
                a[k, j] = a_kj
            if a_kj != 0.0:
                for i in range(k + 1, last_row + 1):
                    a[i, j] += a_kj * a[i, k]
    p[n - 1] = n - 1
    if a[n - 1, n - 1] == 0.0:
        return p, n
    return p, 0


def band_gbtrs(a, ml, smu, p, b):
    n = a.shape[0]
    for k in range(n - 1):
        l = p[k]
        mult = b[l]
        if l != k:
            b[l] = b[k]
            b[k] = mult
        for i in range(k + 1, min(n - 1, k + ml) + 1):
            b[i] += mult * a[i, k]
    for k in range(n - 1, -1, -1):
        b[k] /= a[k, k]
        mult = -b[k]
        for i in range(max(0, k - smu), k):
            b[i] += mult * a[i, k]
    return b
