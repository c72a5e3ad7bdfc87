// decimate.cpp -- native twin of gingr_b200/decimate.py::_collapse (shortest-edge half-edge collapse under the link
// condition, orientation test against the input vertex normals, no-new-slivers rule).  Same rules, same IEEE double
// arithmetic (compile with -ffp-contract=off), same (length, lower id, higher id) pop order, so the result is identical to
// the Python implementation, which stays the specification and the fallback when no host compiler is present.
// Host code only: mesh decimation is caller-side preparation of the multi-resolution schedule, not part of the device path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <queue>
#include <tuple>
#include <vector>

namespace {

using Edge = std::tuple<double, int32_t, int32_t>;
struct V3 { double x, y, z; };

inline double sq(const V3& a, const V3& b) {
  const double dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}
inline V3 normal(const V3& a, const V3& b, const V3& c) {
  const double ux = b.x - a.x, uy = b.y - a.y, uz = b.z - a.z;
  const double wx = c.x - a.x, wy = c.y - a.y, wz = c.z - a.z;
  return {uy * wz - uz * wy, uz * wx - ux * wz, ux * wy - uy * wx};
}

struct Mesh {
  std::vector<V3> P, VN;
  std::vector<int32_t> tri;                 // 3 per triangle, mutated in place
  std::vector<std::vector<int32_t>> vt;     // incident triangle ids per vertex (unordered, unique)

  bool has_tri(int32_t v, int32_t t) const { return std::find(vt[v].begin(), vt[v].end(), t) != vt[v].end(); }
  void drop_tri(int32_t v, int32_t t) {
    auto it = std::find(vt[v].begin(), vt[v].end(), t);
    if (it != vt[v].end()) { *it = vt[v].back(); vt[v].pop_back(); }
  }
  std::vector<int32_t> neighbours(int32_t x) const {
    std::vector<int32_t> out;
    for (int32_t t : vt[x])
      for (int k = 0; k < 3; ++k) {
        const int32_t w = tri[3 * t + k];
        if (w != x) out.push_back(w);
      }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
  }
  std::vector<int32_t> shared(int32_t u, int32_t v) const {
    std::vector<int32_t> out;
    for (int32_t t : vt[u])
      if (has_tri(v, t)) out.push_back(t);
    return out;
  }
  bool is_boundary(int32_t x, const std::vector<int32_t>& nb) const {
    for (int32_t y : nb)
      if (shared(x, y).size() == 1) return true;
    return false;
  }
  bool can_remove(int32_t u, int32_t v, const std::vector<int32_t>& nu, const std::vector<int32_t>& nv) const {
    const std::vector<int32_t> st = shared(u, v);
    if (st.size() != 1 && st.size() != 2) return false;
    std::vector<int32_t> opposite;
    for (int32_t t : st)
      for (int k = 0; k < 3; ++k) {
        const int32_t w = tri[3 * t + k];
        if (w != u && w != v) opposite.push_back(w);
      }
    std::sort(opposite.begin(), opposite.end());
    opposite.erase(std::unique(opposite.begin(), opposite.end()), opposite.end());
    std::vector<int32_t> common;
    std::set_intersection(nu.begin(), nu.end(), nv.begin(), nv.end(), std::back_inserter(common));
    if (common != opposite) return false;
    const bool edge_on_boundary = st.size() == 1;
    if (is_boundary(u, nu) && !edge_on_boundary) return false;
    std::vector<int32_t> uni;
    std::set_union(nu.begin(), nu.end(), nv.begin(), nv.end(), std::back_inserter(uni));
    if ((int)uni.size() - 2 < 3 && !edge_on_boundary) return false;
    for (int32_t t : vt[u]) {
      if (std::find(st.begin(), st.end(), t) != st.end()) continue;
      int32_t ids[3];
      for (int k = 0; k < 3; ++k) ids[k] = tri[3 * t + k] == u ? v : tri[3 * t + k];
      const V3 &a2 = P[ids[0]], &b2 = P[ids[1]], &c2 = P[ids[2]];
      const V3 n1 = normal(a2, b2, c2);
      const double l1 = n1.x * n1.x + n1.y * n1.y + n1.z * n1.z;
      if (l1 == 0.0) return false;
      for (int k = 0; k < 3; ++k) {
        const V3& m = VN[ids[k]];
        const double dot = n1.x * m.x + n1.y * m.y + n1.z * m.z;
        if (dot <= 0.0 || dot * dot < 0.25 * l1) return false;
      }
      const double e2 = sq(a2, b2) + sq(b2, c2) + sq(c2, a2);
      const double q = 0.25 * e2;
      if (12.0 * l1 < q * q) {
        const V3 &a = P[tri[3 * t]], &b = P[tri[3 * t + 1]], &c = P[tri[3 * t + 2]];
        const V3 n0 = normal(a, b, c);
        const double l0 = n0.x * n0.x + n0.y * n0.y + n0.z * n0.z;
        const double e0 = sq(a, b) + sq(b, c) + sq(c, a);
        if (l1 * e0 * e0 < l0 * e2 * e2) return false;
      }
    }
    return true;
  }
};

}  // namespace

extern "C" {

// pts [n][3], tri [T][3], vn [n][3] unit vertex normals of the input (computed by the caller so that both implementations
// use the same numbers).  keep [n] receives 0 / 1; tri_out [T][3] the surviving triangles over the ORIGINAL numbering,
// *T_out their count.  Returns the number of kept vertices.
int32_t gingr_host_decimate(int32_t n, const double* pts, int32_t T, const int32_t* tri, const double* vn, int32_t target,
                            uint8_t* keep, int32_t* tri_out, int32_t* T_out) {
  Mesh m;
  m.P.resize(n);
  m.VN.resize(n);
  for (int32_t i = 0; i < n; ++i) {
    m.P[i] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
    m.VN[i] = {vn[3 * i], vn[3 * i + 1], vn[3 * i + 2]};
  }
  m.tri.assign(tri, tri + (size_t)3 * T);
  m.vt.assign(n, {});
  for (int32_t t = 0; t < T; ++t)
    for (int k = 0; k < 3; ++k) {
      const int32_t v = tri[3 * t + k];
      if (!m.has_tri(v, t)) m.vt[v].push_back(t);
    }
  int64_t alive = 0;
  for (int32_t i = 0; i < n; ++i) alive += !m.vt[i].empty();

  std::vector<Edge> init;
  init.reserve((size_t)3 * T);
  for (int32_t t = 0; t < T; ++t)
    for (int k = 0; k < 3; ++k) {
      int32_t x = tri[3 * t + k], y = tri[3 * t + (k + 1) % 3];
      if (x > y) std::swap(x, y);
      init.emplace_back(0.0, x, y);
    }
  std::sort(init.begin(), init.end());
  init.erase(std::unique(init.begin(), init.end()), init.end());
  for (Edge& e : init) {
    const int32_t x = std::get<1>(e), y = std::get<2>(e);
    std::get<0>(e) = (m.P[x].x - m.P[y].x) * (m.P[x].x - m.P[y].x) + (m.P[x].y - m.P[y].y) * (m.P[x].y - m.P[y].y) +
                     (m.P[x].z - m.P[y].z) * (m.P[x].z - m.P[y].z);
  }
  using Heap = std::priority_queue<Edge, std::vector<Edge>, std::greater<Edge>>;
  Heap heap(std::greater<Edge>(), std::move(init));
  std::vector<Edge> deferred;
  bool progress = false;
  while (alive > target && alive > 4) {
    if (heap.empty()) {
      if (!progress || deferred.empty()) break;
      heap = Heap(std::greater<Edge>(), std::move(deferred));
      deferred.clear();
      progress = false;
      continue;
    }
    const Edge e = heap.top();
    heap.pop();
    const int32_t x = std::get<1>(e), y = std::get<2>(e);
    if (m.vt[x].empty() || m.vt[y].empty() || m.shared(x, y).empty()) continue;
    const std::vector<int32_t> nx = m.neighbours(x), ny = m.neighbours(y);
    bool done = false;
    for (int dir = 0; dir < 2 && !done; ++dir) {
      const int32_t u = dir == 0 ? y : x, v = dir == 0 ? x : y;
      const std::vector<int32_t>& nu = dir == 0 ? ny : nx;
      const std::vector<int32_t>& nv = dir == 0 ? nx : ny;
      if (!m.can_remove(u, v, nu, nv)) continue;
      for (int32_t t : m.shared(u, v))
        for (int k = 0; k < 3; ++k) m.drop_tri(m.tri[3 * t + k], t);
      const std::vector<int32_t> ut = m.vt[u];
      for (int32_t t : ut) {
        for (int k = 0; k < 3; ++k)
          if (m.tri[3 * t + k] == u) { m.tri[3 * t + k] = v; break; }
        if (!m.has_tri(v, t)) m.vt[v].push_back(t);
      }
      m.vt[u].clear();
      alive -= 1;
      for (int32_t w : nu)
        if (m.vt[w].empty()) alive -= 1;
      for (int32_t w : nu) {
        if (w == v || m.vt[w].empty() || std::binary_search(nv.begin(), nv.end(), w)) continue;
        const double dd = (m.P[v].x - m.P[w].x) * (m.P[v].x - m.P[w].x) + (m.P[v].y - m.P[w].y) * (m.P[v].y - m.P[w].y) +
                          (m.P[v].z - m.P[w].z) * (m.P[v].z - m.P[w].z);
        heap.emplace(dd, std::min(v, w), std::max(v, w));
      }
      progress = true;
      done = true;
    }
    if (!done) deferred.push_back(e);
  }
  int32_t kept = 0;
  for (int32_t i = 0; i < n; ++i) {
    keep[i] = m.vt[i].empty() ? 0 : 1;
    kept += keep[i];
  }
  std::vector<uint8_t> live((size_t)T, 0);
  for (int32_t i = 0; i < n; ++i)
    for (int32_t t : m.vt[i]) live[t] = 1;
  int32_t nt = 0;
  for (int32_t t = 0; t < T; ++t)
    if (live[t]) {
      tri_out[3 * nt] = m.tri[3 * t];
      tri_out[3 * nt + 1] = m.tri[3 * t + 1];
      tri_out[3 * nt + 2] = m.tri[3 * t + 2];
      ++nt;
    }
  *T_out = nt;
  return kept;
}

}  // extern "C"
