"""Synthetic SO-DSO-like scan sets.

Every scan has `npts` points in the camera frame (x right, y down, z forward), cropped to
|p| < 45 m like pts_preprocess.h:144, with well separated PCA eigenvalues and a few wall /
pole clusters so that the height-range bins are non-trivial.  Intensities are k/8 with
k in 0..2040 (mean of an 8-pixel uint8 pattern, OutputWrapperSODSO.cpp:47-50) so that float
sums over a scan are exact and therefore order-independent (SC.cpp:60-64).

Seeds: scan s uses numpy SeedSequence(20261017 + s).  Used by tests, bench.py and smoke().
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20261017
MAX_RHO = 45.0


def make_scan(index: int, npts: int = 4096):
    rng = np.random.default_rng(BASE_SEED + index)
    n_cluster = 8
    n_cl_pts = npts // 4
    n_bg = npts - n_cl_pts
    pts = np.empty((0, 3))
    # background: flat-ish ground slab
    while pts.shape[0] < n_bg:
        k = n_bg * 2
        x = rng.uniform(-25, 25, k)
        z = rng.uniform(-40, 40, k)
        y = np.clip(rng.normal(0.0, 1.5, k), -6, 6)
        p = np.stack([x, y, z], axis=1)
        p = p[np.linalg.norm(p, axis=1) < MAX_RHO]
        pts = np.concatenate([pts, p])[:n_bg]
    # walls / poles: gaussian blobs, sigma 1 m in plan, 4 m height spread
    cl = np.empty((0, 3))
    centres = np.stack([rng.uniform(-22, 22, n_cluster), np.zeros(n_cluster), rng.uniform(-35, 35, n_cluster)], axis=1)
    while cl.shape[0] < n_cl_pts:
        k = n_cl_pts * 2
        c = centres[rng.integers(0, n_cluster, k)]
        p = c + np.stack([rng.normal(0, 1.0, k), rng.uniform(-2.0, 2.0, k), rng.normal(0, 1.0, k)], axis=1)
        p = p[np.linalg.norm(p, axis=1) < MAX_RHO]
        cl = np.concatenate([cl, p])[:n_cl_pts]
    xyz = np.concatenate([pts, cl])
    # intensity correlates with position so that the binary channel carries place information
    base = 1020 + 600 * np.sin(xyz[:, 0] * 0.21 + index * 0.37) * np.cos(xyz[:, 2] * 0.13 - index * 0.11)
    k8 = np.clip(np.rint(base + rng.normal(0, 120, npts)), 0, 2040)
    inten = (k8 / 8.0).astype(np.float32)
    perm = rng.permutation(npts)
    return np.ascontiguousarray(xyz[perm]), np.ascontiguousarray(inten[perm])


def revisit(xyz, inten, index: int, jitter: float = 0.05, resample: float = 0.10):
    """Planted loop: rotate about the vertical (y, PCA-up) axis by a random yaw, jitter, resample 10%."""
    rng = np.random.default_rng(BASE_SEED + 7_000_000 + index)
    yaw = rng.uniform(0, 2 * np.pi)
    c, s = np.cos(yaw), np.sin(yaw)
    out = xyz.copy()
    out[:, 0] = c * xyz[:, 0] + s * xyz[:, 2]
    out[:, 2] = -s * xyz[:, 0] + c * xyz[:, 2]
    out += rng.normal(0, jitter, out.shape)
    it = inten.copy()
    n = xyz.shape[0]
    k = int(n * resample)
    sel = rng.choice(n, k, replace=False)
    x = rng.uniform(-25, 25, k)
    z = rng.uniform(-40, 40, k)
    y = np.clip(rng.normal(0.0, 1.5, k), -6, 6)
    p = np.stack([x, y, z], axis=1)
    nr = np.linalg.norm(p, axis=1)
    p[nr >= MAX_RHO] *= (MAX_RHO - 1.0) / nr[nr >= MAX_RHO, None]
    out[sel] = p
    it[sel] = (rng.integers(0, 2041, k) / 8.0).astype(np.float32)
    # keep inside the crop
    nr = np.linalg.norm(out, axis=1)
    bad = nr >= MAX_RHO
    out[bad] *= ((MAX_RHO - 1e-3) / nr[bad])[:, None]
    return np.ascontiguousarray(out), it


def make_scan_set(nscans: int, npts: int = 4096, planted_loops: bool = False, first: int = 0):
    """-> xyz (nscans*npts, 3) f64, inten f32, off int64[nscans+1].

    planted_loops: scans [nscans/2, nscans) revisit scans [0, nscans/2)."""
    xyz = np.empty((nscans * npts, 3))
    inten = np.empty(nscans * npts, dtype=np.float32)
    half = nscans // 2 if planted_loops else nscans
    for s in range(nscans):
        if s < half or not planted_loops:
            p, it = make_scan(first + s, npts)
        else:
            src = s - half
            p, it = revisit(xyz[src * npts:(src + 1) * npts], inten[src * npts:(src + 1) * npts], first + s)
        xyz[s * npts:(s + 1) * npts] = p
        inten[s * npts:(s + 1) * npts] = it
    off = np.arange(nscans + 1, dtype=np.int64) * npts
    return xyz, inten, off
