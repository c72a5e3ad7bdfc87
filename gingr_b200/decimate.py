"""Mesh decimation for the multi-resolution callers (SimpleRegistrator.decimateState, SimpleRegistrator.scala:84-106;
IndependentPointDistanceEvaluator.scala:39-42).

SUBSTITUTION (stated wherever results are reported): scalismo's `mesh.operations.decimate(n)` is a quadric
edge-collapse simplification inside scalismo 1.0-RC1, which is not in the reference tree.  This module is a deterministic
SHORTEST-EDGE HALF-EDGE COLLAPSE instead: the shortest remaining edge is collapsed onto one of its end points (vertices
never move, so the result is a subset of the input vertices), subject to the link condition (the surface stays a manifold,
closed meshes stay closed, boundaries stay boundaries), an orientation test of every changed triangle against the INPUT
surface's vertex normals, and a no-new-slivers rule.  Same role --
a coarser valid triangulation of about n vertices on which the model is re-referenced with the nearest-neighbour
interpolator -- but not scalismo's vertices.  Point clouds (no triangles) are thinned by uniform-grid clustering.
Host side, outside the hot path: a pure Python / numpy specification (about 65 microseconds per removed vertex) and a native
twin (csrc_host/decimate.cpp, compiled on first use, identical result, about 10x faster) that is used when a host C++
compiler is present."""
from __future__ import annotations

import heapq
from typing import Tuple

import numpy as np


# ---------------------------------------------------------------------------------------------
# point clouds: one representative per occupied grid cell
# ---------------------------------------------------------------------------------------------
def _cell_keys(p: np.ndarray, lo: np.ndarray, h: float) -> np.ndarray:
    c = np.floor((p - lo) / h).astype(np.int64)
    dims = c.max(axis=0) + 1
    return (c[:, 2] * dims[1] + c[:, 1]) * dims[0] + c[:, 0]


def decimate_points(points, n_target: int) -> np.ndarray:
    """Indices (ascending) of about n_target input points: the point nearest to the centroid of each occupied cell of a
    uniform grid whose cell size is searched for the requested count."""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    n = int(n_target)
    if n <= 0:
        raise ValueError("decimate: the requested number of points must be positive")
    if n >= p.shape[0]:
        return np.arange(p.shape[0])
    lo = p.min(axis=0)
    extent = float(np.max(p.max(axis=0) - lo))
    if extent == 0.0:
        return np.arange(1)
    h_lo, h_hi = extent * 1e-6, extent * 1.0000001
    best_h, best_err = h_hi, None
    for _ in range(48):                                           # occupied cells fall with the cell size: bisect on log h
        h = float(np.sqrt(h_lo * h_hi))
        cnt = np.unique(_cell_keys(p, lo, h)).size
        err = abs(cnt - n)
        if best_err is None or err < best_err:
            best_h, best_err = h, err
        if cnt > n:
            h_lo = h
        elif cnt < n:
            h_hi = h
        else:
            break
    keys = _cell_keys(p, lo, best_h)
    uniq, inv = np.unique(keys, return_inverse=True)
    k = uniq.size
    cnt = np.bincount(inv, minlength=k).astype(np.float64)
    cen = np.stack([np.bincount(inv, weights=p[:, d], minlength=k) / cnt for d in range(3)], axis=1)
    d2 = ((p - cen[inv]) ** 2).sum(axis=1)
    order = np.lexsort((np.arange(p.shape[0]), d2, inv))          # per cell: nearest to the centroid, lowest index on ties
    first = np.ones(order.size, dtype=bool)
    first[1:] = inv[order][1:] != inv[order][:-1]
    return np.sort(order[first])


# ---------------------------------------------------------------------------------------------
# triangle meshes: shortest-edge half-edge collapse
# ---------------------------------------------------------------------------------------------
def _input_vertex_normals(p: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Unit vertex normals of the input mesh (normalised sum of the incident triangle normals; zero where undefined)."""
    tt = np.asarray(t, dtype=np.int64)
    fn = np.cross(p[tt[:, 1]] - p[tt[:, 0]], p[tt[:, 2]] - p[tt[:, 0]])
    vn = np.zeros_like(p)
    for k in range(3):
        np.add.at(vn, tt[:, k], fn)
    ln = np.sqrt((vn * vn).sum(1))
    vn[ln > 0] /= ln[ln > 0][:, None]
    return vn


_NATIVE = None


def _native_lib():
    """libgingr_host.so (csrc_host/decimate.cpp), compiled on first use with the host C++ compiler; None when that is not
    possible -- the Python implementation below is the specification and produces the identical result, only slower."""
    global _NATIVE
    if _NATIVE is None:
        import ctypes
        import os
        import shutil
        import subprocess
        here = os.path.dirname(os.path.abspath(__file__))
        src = os.path.join(here, "csrc_host", "decimate.cpp")
        so = os.path.join(here, "lib", "libgingr_host.so")
        _NATIVE = False
        try:
            if os.environ.get("GINGR_HOST_NATIVE", "1") == "0":
                raise OSError("disabled")
            if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
                cxx = shutil.which("g++") or shutil.which("c++")
                if cxx is None:
                    raise OSError("no host C++ compiler")
                os.makedirs(os.path.dirname(so), exist_ok=True)
                # several processes (ranks, pytest-xdist) may get here at once: build to a private name, rename atomically
                tmp = f"{so}.{os.getpid()}.tmp"
                try:
                    subprocess.run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", src, "-o", tmp],
                                   check=True, capture_output=True)
                    os.replace(tmp, so)
                finally:
                    if os.path.exists(tmp):
                        os.remove(tmp)
            lib = ctypes.CDLL(so)
            lib.gingr_host_decimate.restype = ctypes.c_int32
            _NATIVE = lib
        except (OSError, subprocess.CalledProcessError) as e:
            if os.environ.get("GINGR_HOST_NATIVE", "1") != "0":
                import warnings
                warnings.warn(f"gingr_b200.decimate: native helper unavailable ({e}); using the Python specification "
                              "(identical result, slower)")
            _NATIVE = False
    return _NATIVE or None


def _collapse_native(lib, p: np.ndarray, t: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray]:
    import ctypes
    pts = np.ascontiguousarray(p, dtype=np.float64)
    tri = np.ascontiguousarray(t, dtype=np.int32)
    vn = np.ascontiguousarray(_input_vertex_normals(pts, tri))
    keep = np.zeros(pts.shape[0], dtype=np.uint8)
    tri_out = np.empty_like(tri)
    t_out = ctypes.c_int32()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
    lib.gingr_host_decimate(ctypes.c_int32(pts.shape[0]), pts.ctypes.data_as(dp), ctypes.c_int32(tri.shape[0]), tri.ctypes.data_as(ip),
                            vn.ctypes.data_as(dp), ctypes.c_int32(n), keep.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                            tri_out.ctypes.data_as(ip), ctypes.byref(t_out))
    return np.flatnonzero(keep).astype(np.int64), tri_out[:t_out.value].astype(np.int64).reshape(-1, 3)


def _collapse(p: np.ndarray, t: np.ndarray, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """-> (kept vertex indices ascending, triangles over the ORIGINAL vertex numbering)."""
    P = [tuple(row) for row in p.tolist()]
    tris = [list(row) for row in t.tolist()]
    vt = [set() for _ in range(len(P))]                           # incident triangle ids per vertex
    for ti, (a, b, c) in enumerate(tris):
        vt[a].add(ti); vt[b].add(ti); vt[c].add(ti)
    alive = sum(1 for s in vt if s)                               # vertices used by a triangle
    VN = [tuple(row) for row in _input_vertex_normals(p, t).tolist()]

    def neighbours(x):
        out = set()
        for ti in vt[x]:
            out.update(tris[ti])
        out.discard(x)
        return out

    def is_boundary(x, nbrs):
        return any(len(vt[x] & vt[y]) == 1 for y in nbrs)

    def sq(a, b):
        dx, dy, dz = a[0] - b[0], a[1] - b[1], a[2] - b[2]
        return dx * dx + dy * dy + dz * dz                         # explicit products (no pow): the native twin does the same

    def normal(a, b, c):
        ux, uy, uz = b[0] - a[0], b[1] - a[1], b[2] - a[2]
        wx, wy, wz = c[0] - a[0], c[1] - a[1], c[2] - a[2]
        return uy * wz - uz * wy, uz * wx - ux * wz, ux * wy - uy * wx

    def can_remove(u, v, nu, nv):
        """Half-edge collapse u -> v: link condition, boundary rules, no fold-over."""
        shared_t = vt[u] & vt[v]
        if len(shared_t) not in (1, 2):
            return False
        opposite = set()
        for ti in shared_t:
            opposite.update(tris[ti])
        opposite.discard(u); opposite.discard(v)
        if (nu & nv) != opposite:
            return False
        edge_on_boundary = len(shared_t) == 1
        if is_boundary(u, nu) and not edge_on_boundary:
            return False                                          # a boundary vertex only slides along its boundary
        if len(nu | nv) - 2 < 3 and not edge_on_boundary:
            return False                                          # would leave a two-triangle pillow
        for ti in vt[u]:
            if ti in shared_t:
                continue
            ids = [v if x == u else x for x in tris[ti]]
            a2, b2, c2 = P[ids[0]], P[ids[1]], P[ids[2]]
            n1 = normal(a2, b2, c2)
            l1 = n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2]
            if l1 == 0.0:
                return False
            # orientation against the INPUT surface (its vertex normals), not against the previous triangle: turning is
            # not allowed to accumulate over successive collapses
            for x in ids:
                m = VN[x]
                dot = n1[0] * m[0] + n1[1] * m[1] + n1[2] * m[2]
                if dot <= 0.0 or dot * dot < 0.25 * l1:                 # more than 60 degrees off the input surface
                    return False
            # no new slivers: quality 2 sqrt(3) |n| / (sum of squared edges) is 1 for an equilateral triangle
            e2 = (sq(a2, b2) + sq(b2, c2) + sq(c2, a2))
            q = 0.25 * e2
            if 12.0 * l1 < q * q:
                a, b, c = (P[x] for x in tris[ti])
                n0 = normal(a, b, c)
                l0 = n0[0] * n0[0] + n0[1] * n0[1] + n0[2] * n0[2]
                e0 = sq(a, b) + sq(b, c) + sq(c, a)
                if l1 * e0 * e0 < l0 * e2 * e2:                            # worse than the triangle it replaces
                    return False
        return True

    heap = []
    seen = set()
    for a, b, c in tris:
        for x, y in ((a, b), (b, c), (c, a)):
            e = (x, y) if x < y else (y, x)
            if e not in seen:
                seen.add(e)
                heap.append((sq(P[x], P[y]), e[0], e[1]))
    del seen
    heapq.heapify(heap)
    deferred = []                                                 # edges that could not collapse yet; retried after progress
    progress = False
    while alive > n and alive > 4:
        if not heap:
            if not progress or not deferred:
                break
            heap, deferred, progress = deferred, [], False
            heapq.heapify(heap)
            continue
        d, x, y = heapq.heappop(heap)
        if not vt[x] or not vt[y] or not (vt[x] & vt[y]):
            continue                                              # stale: a vertex or the edge is gone
        nx, ny = neighbours(x), neighbours(y)
        # remove the higher index first (deterministic), else the other direction
        for u, v, nu, nv in ((y, x, ny, nx), (x, y, nx, ny)):
            if can_remove(u, v, nu, nv):
                shared_t = vt[u] & vt[v]
                for ti in shared_t:
                    for w in tris[ti]:
                        vt[w].discard(ti)
                for ti in list(vt[u]):
                    tri = tris[ti]
                    tri[tri.index(u)] = v
                    vt[v].add(ti)
                vt[u] = set()
                alive -= 1
                alive -= sum(1 for w in nu if not vt[w])           # a vertex left without triangles (open meshes)
                for w in nu - nv:
                    if w != v and vt[w]:
                        heapq.heappush(heap, (sq(P[v], P[w]), min(v, w), max(v, w)))
                progress = True
                break
        else:
            deferred.append((d, x, y))
    kept = np.array([i for i, s in enumerate(vt) if s], dtype=np.int64)
    live = sorted({ti for s in vt for ti in s})
    return kept, np.array([tris[ti] for ti in live], dtype=np.int64).reshape(-1, 3)


def decimate(points, triangles, n_target: int) -> Tuple[np.ndarray, np.ndarray]:
    """-> (points' [n', 3] -- a subset of the input vertices, in input order --, triangles' [t', 3] int32) with n' = n_target
    whenever the collapse rules allow it (never more than the input).  Deterministic."""
    p = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
    n = int(n_target)
    if n <= 0:
        raise ValueError("decimate: the requested number of points must be positive")
    if triangles is None or np.asarray(triangles).size == 0:
        return p[decimate_points(p, n)].copy(), np.zeros((0, 3), np.int32)
    t = np.asarray(triangles, dtype=np.int32).reshape(-1, 3)
    if t.min() < 0 or t.max() >= p.shape[0]:
        raise ValueError("decimate: triangle index out of range")
    if n >= p.shape[0]:
        return p.copy(), t.copy()
    lib = _native_lib()
    kept, nt = _collapse_native(lib, p, t, n) if lib is not None else _collapse(p, t, n)
    new_id = np.full(p.shape[0], -1, dtype=np.int64)
    new_id[kept] = np.arange(kept.size)
    return p[kept].copy(), new_id[nt].astype(np.int32)
