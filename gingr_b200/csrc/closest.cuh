// closest.cuh -- host-side interface of the K2 closest-point kernels (closest.cu).
#pragma once
#include <vector>

#include "common.cuh"

namespace gingr {

struct ClosestWorkspace {
  int s_nn = 1, s_surf = 1, s_line = 1;  // candidate-range splits of the three scans
  DevBuf<double> part_d2;     // [smax][M]
  DevBuf<int32_t> part_idx;   // [smax][M]
  DevBuf<double> part_cp;     // [s_surf][M][3]
  DevBuf<double> d2;          // [M]
  DevBuf<int32_t> idx;        // [M]   nearest target vertex
  DevBuf<double> cp;          // [M][3] corresponding point
  DevBuf<uint8_t> w;          // [M]   0/1 weight
  DevBuf<double> n_tpl;       // [M][3] template vertex normals
  DevBuf<double> mean_dist;   // [1]
  int32_t ensure(gingr_ctx* ctx, int M, int N, int T_target, int T_template);
  void release();
};

void build_vertex_adjacency(int n, int T, const int32_t* tri, std::vector<int32_t>* off, std::vector<int32_t>* adj);
void compute_boundary_flags(int n, int T, const int32_t* tri, std::vector<uint8_t>* flags);
int32_t mesh_static_upload(gingr_ctx* ctx, int n, const double* verts_aos_host, int T, const int32_t* tri_host,
                           DevBuf<double>* normals, DevBuf<uint8_t>* boundary);
int32_t vertex_normals_enqueue(gingr_ctx* ctx, int n, const double* d_verts_aos, const int32_t* d_tri,
                               const int32_t* d_adj_off, const int32_t* d_adj, double* d_normals);
int32_t nn_vertex_enqueue(gingr_ctx* ctx, ClosestWorkspace& ws, int M, const double* d_q, int N,
                          const double* d_pts_soa, double* d_d2, int32_t* d_idx);
int32_t icp_closest_enqueue(gingr_ctx* ctx, ClosestWorkspace& ws, const gingr_target* tgt, int M, const double* d_tpl,
                            int T_tpl, const int32_t* d_tpl_tri, const int32_t* d_adj_off, const int32_t* d_adj,
                            int method);

}  // namespace gingr
