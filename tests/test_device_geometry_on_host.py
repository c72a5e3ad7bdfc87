"""The per-candidate arithmetic of the K2 kernels (gingr_b200/csrc/closest_geom.cuh) compiled FOR THE HOST with g++
(the CUDA qualifiers degrade to ignored attributes; -ffp-contract=off plays the role of nvcc's -fmad=false) and compared
with the oracle's C restatement on the same inputs: the shipped device source, not a copy, evaluated on the CPU.  No GPU."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="needs g++ and the CUDA headers")

WRAPPER = r'''
#include <cmath>
using std::sqrt;
#include "closest_geom.cuh"
extern "C" {
void host_closest_on_triangle(int n, const double* q, const double* tri, double* out) {
  for (int i = 0; i < n; ++i)
    gingr::closest_on_triangle(q[3 * i], q[3 * i + 1], q[3 * i + 2], tri + 9 * i, tri + 9 * i + 3, tri + 9 * i + 6,
                               out[3 * i], out[3 * i + 1], out[3 * i + 2]);
}
void host_line_triangle_hit(int n, const double* o, const double* d, const double* tri, int* hit, double* dist, double* pt) {
  for (int i = 0; i < n; ++i) {
    double s = 0.0, x = 0.0, y = 0.0, z = 0.0;
    hit[i] = gingr::line_triangle_hit(o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], tri + 9 * i,
                                      tri + 9 * i + 3, tri + 9 * i + 6, s, x, y, z) ? 1 : 0;
    dist[i] = s; pt[3 * i] = x; pt[3 * i + 1] = y; pt[3 * i + 2] = z;
  }
}
}
'''


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostgeom")
    src = d / "wrap.cpp"
    src.write_text(WRAPPER)
    so = str(d / "libhostgeom.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-attributes", "-I" + CUDA_INC,
           "-I" + os.path.join(ROOT, "gingr_b200", "csrc"), str(src), "-o", so]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return ctypes.CDLL(so)


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def test_closest_on_triangle_source_equals_the_oracle_bit_for_bit(hostlib, oracle):
    rng = np.random.default_rng(0)
    n = 4000
    tri = rng.normal(size=(n, 3, 3)) * rng.choice([1.0, 1e-3, 80.0], size=(n, 1, 1))
    tri[::7, 2] = tri[::7, 0] + (tri[::7, 1] - tri[::7, 0]) * 0.4 + 1e-9 * rng.normal(size=(len(tri[::7]), 3))     # slivers
    q = rng.normal(size=(n, 3)) * rng.choice([0.5, 5.0, 200.0], size=(n, 1))
    q[::5] = tri[::5, 0] + 0.3 * (tri[::5, 1] - tri[::5, 0])                                                        # on an edge
    q[::11] = tri[::11, 2]                                                                                          # at a vertex
    out = np.empty((n, 3))
    hostlib.host_closest_on_triangle(n, _dp(np.ascontiguousarray(q)), _dp(np.ascontiguousarray(tri)), _dp(out))
    t012 = np.array([[0, 1, 2]], dtype=np.int32)
    for i in range(n):
        cp = oracle.closest_on_surface(q[i:i + 1], tri[i], t012)[0][0]
        assert np.array_equal(out[i], cp), (i, out[i], cp)


def test_line_triangle_hit_source_equals_the_oracle(hostlib, oracle):
    rng = np.random.default_rng(1)
    n = 3000
    tri = rng.normal(size=(n, 3, 3)) * 10.0
    o = rng.normal(size=(n, 3)) * 10.0
    inside = tri[:, 0] * 0.5 + tri[:, 1] * 0.3 + tri[:, 2] * 0.2
    d = inside - o + rng.normal(size=(n, 3)) * rng.choice([0.0, 3.0, 30.0], size=(n, 1))      # many hits, some misses
    hit = np.empty(n, dtype=np.int32)
    dist = np.empty(n)
    pt = np.empty((n, 3))
    hostlib.host_line_triangle_hit(n, _dp(np.ascontiguousarray(o)), _dp(np.ascontiguousarray(d)), _dp(np.ascontiguousarray(tri)),
                                   hit.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _dp(dist), _dp(pt))
    t012 = np.array([[0, 1, 2]], dtype=np.int32)
    n_hit = 0
    for i in range(n):
        od, oh = oracle.line_mesh_nearest(o[i:i + 1], d[i:i + 1], tri[i], t012)
        if np.isinf(od[0]):
            assert hit[i] == 0, i
        else:
            n_hit += 1
            assert hit[i] == 1 and dist[i] == od[0] and np.array_equal(pt[i], oh[0]), (i, dist[i], od[0])
    assert 0.3 * n < n_hit < n
