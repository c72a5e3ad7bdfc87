// closest_geom.cuh -- the per-candidate arithmetic of the K2 closest-point kernels, shared by the brute-force scans
// (closest.cu) and the uniform-grid searches (grid.cu) so that both evaluate every candidate with the SAME individually
// rounded FP64 operations (both translation units are compiled with -fmad=false).  Exactness argument: a search
// structure only decides WHICH candidates are evaluated; as long as every candidate that can win is evaluated with
// this code and the winner is chosen by (value, lowest index), the result is bit-identical to the full scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gingr {

// ---------------------------------------------------------------------------------------------
// closest point on a triangle (Ericson 5.1.5).  Same operation order as the oracle.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void closest_on_triangle(double px, double py, double pz, const double* a, const double* b,
                                                    const double* c, double& ox, double& oy, double& oz) {
  const double abx = b[0] - a[0], aby = b[1] - a[1], abz = b[2] - a[2];
  const double acx = c[0] - a[0], acy = c[1] - a[1], acz = c[2] - a[2];
  const double apx = px - a[0], apy = py - a[1], apz = pz - a[2];
  const double d1 = abx * apx + aby * apy + abz * apz;
  const double d2 = acx * apx + acy * apy + acz * apz;
  if (d1 <= 0.0 && d2 <= 0.0) { ox = a[0]; oy = a[1]; oz = a[2]; return; }
  const double bpx = px - b[0], bpy = py - b[1], bpz = pz - b[2];
  const double d3 = abx * bpx + aby * bpy + abz * bpz;
  const double d4 = acx * bpx + acy * bpy + acz * bpz;
  if (d3 >= 0.0 && d4 <= d3) { ox = b[0]; oy = b[1]; oz = b[2]; return; }
  const double vc = d1 * d4 - d3 * d2;
  if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) {
    const double v = d1 / (d1 - d3);
    ox = a[0] + v * abx; oy = a[1] + v * aby; oz = a[2] + v * abz; return;
  }
  const double cpx = px - c[0], cpy = py - c[1], cpz = pz - c[2];
  const double d5 = abx * cpx + aby * cpy + abz * cpz;
  const double d6 = acx * cpx + acy * cpy + acz * cpz;
  if (d6 >= 0.0 && d5 <= d6) { ox = c[0]; oy = c[1]; oz = c[2]; return; }
  const double vb = d5 * d2 - d1 * d6;
  if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) {
    const double w = d2 / (d2 - d6);
    ox = a[0] + w * acx; oy = a[1] + w * acy; oz = a[2] + w * acz; return;
  }
  const double va = d3 * d6 - d5 * d4;
  if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
    const double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    ox = b[0] + w * (c[0] - b[0]); oy = b[1] + w * (c[1] - b[1]); oz = b[2] + w * (c[2] - b[2]); return;
  }
  const double denom = 1.0 / (va + vb + vc);
  const double v = vb * denom, w = vc * denom;
  ox = a[0] + abx * v + acx * w;
  oy = a[1] + aby * v + acy * w;
  oz = a[2] + abz * v + acz * w;
}

// Intersection of the infinite line o + s d with triangle (a, b, c) (scalismo getIntersectionPoints, SURVEY.md A6;
// Moeller-Trumbore without the s >= 0 test).  Returns false for no hit or a hit equal to o (`.filter(f => f != p)`,
// ClosestPointRegistrator.scala:66); otherwise the hit point and its distance from o.
__device__ __forceinline__ bool line_triangle_hit(double ox, double oy, double oz, double dx, double dy, double dz,
                                                  const double* a, const double* b, const double* c, double& dist,
                                                  double& ix, double& iy, double& iz) {
  const double e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
  const double e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
  const double hx = dy * e2z - dz * e2y, hy = dz * e2x - dx * e2z, hz = dx * e2y - dy * e2x;
  const double det = e1x * hx + e1y * hy + e1z * hz;
  if (det == 0.0) return false;
  const double inv = 1.0 / det;
  const double tx = ox - a[0], ty = oy - a[1], tz = oz - a[2];
  const double u = (tx * hx + ty * hy + tz * hz) * inv;
  if (u < 0.0 || u > 1.0) return false;
  const double qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
  const double vv = (dx * qx + dy * qy + dz * qz) * inv;
  if (vv < 0.0 || u + vv > 1.0) return false;
  const double s = (e2x * qx + e2y * qy + e2z * qz) * inv;
  ix = ox + s * dx; iy = oy + s * dy; iz = oz + s * dz;
  if (ix == ox && iy == oy && iz == oz) return false;
  const double ex = ox - ix, ey = oy - iy, ez = oz - iz;
  dist = sqrt(ex * ex + ey * ey + ez * ez);
  return true;
}

}  // namespace gingr
