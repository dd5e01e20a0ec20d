// sgc_lb.cu — the lattice-Boltzmann application's per-cell loops on the device.
//
// The reference's only multi-box application (application/latticeBoltzmann) advances a
// 19-component D3Q19 level by   collision -> exchange(PeriodicX|PeriodicY, TrimCorner) ->
// bounce-back ghost fill -> shifted-copy stream (+ level swap) -> macroscopic moments
// (LBLevel.cpp:7-127).  Exchange, bounce-back and stream are copies: they run through the batched
// copy kernel (exchange plan, copy sets).  This file holds the two floating-point loops,
// LBPatch::collision / LBPhysics::collision (LBPatch.H:27-46, LBPhysics.H:5-14) and
// LBPatch::macroscopic / LBPhysics::macroscopic (LBPatch.H:9-23, LBPhysics.H:17-35), the
// generators of the two copy lists, and the device-resident step loop.
//
// Arithmetic follows the reference's source expression operation for operation with no
// contraction (the library is compiled --fmad=false), i.e. it is bit-identical to the reference
// built with -ffp-contract=off and within rounding (<= 1e-12 relative) of its default build,
// whose contractions are the compiler's choice.  Both kernels stream every value once:
// collision 19 reads + 4 reads + 19 writes, macroscopic 19 reads + 4 writes per cell.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "sgc_device.h"

namespace sgc {

namespace {

constexpr int kQ = 19;          // LBParameters::g_numVelDir
constexpr int kMacro = 4;       // density + velocity (LBParameters::g_numState)

// LBParameters.H:38-49,71-96,134-140
__constant__ double c_w[kQ] = {1. / 3.,  1. / 18., 1. / 18., 1. / 18., 1. / 18., 1. / 18., 1. / 18.,
                               1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.,
                               1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36., 1. / 36.};
__constant__ int c_e[kQ][3] = {{0, 0, 0},   {-1, 0, 0}, {1, 0, 0},  {0, -1, 0}, {0, 1, 0},   {0, 0, -1}, {0, 0, 1},
                               {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0},  {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1},
                               {1, 0, 1},   {0, -1, -1}, {0, 1, -1}, {0, -1, 1}, {0, 1, 1}};
const int h_e[kQ][3] = {{0, 0, 0},   {-1, 0, 0}, {1, 0, 0},  {0, -1, 0}, {0, 1, 0},   {0, 0, -1}, {0, 0, 1},
                        {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0},  {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1},
                        {1, 0, 1},   {0, -1, -1}, {0, 1, -1}, {0, -1, 1}, {0, 1, 1}};
const int h_opp[kQ] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
constexpr double kTau = 0.516, kCs2 = 1.0 / 3, kForce = 0.000001042;   // LBParameters.H:52-54

constexpr int kLbThreads = 256;

// element offset (component 0) of valid cell number t of a fab, cells numbered x fastest
__device__ __forceinline__ long long valid_cell(const FabGeom& g, int t) {
  const int i = t % g.vn[0];
  const int r = t / g.vn[0];
  const int j = r % g.vn[1], k = r / g.vn[1];
  return fab_index(g, i + g.vlo[0] - g.lo[0], j + g.vlo[1] - g.lo[1], k + g.vlo[2] - g.lo[2], 0);
}

// LBPhysics::collision for the 19 directions of one cell, in place
__device__ __forceinline__ void collide_cell(double* f /* [kQ] */, double rho, double u0, double u1, double u2) {
  const double usq = ((u0 * u0 + u1 * u1) + u2 * u2) / (2 * kCs2);
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    const double eu = (u0 * c_e[q][0] + u1 * c_e[q][1]) + u2 * c_e[q][2];
    const double feq = (c_w[q] * rho) * (((1 + eu / kCs2) + (eu * eu) / (2 * kCs2 * kCs2)) - usq);
    f[q] = (f[q] + (feq - f[q]) / kTau) + 3 * c_w[q] * c_e[q][0] * kForce;
  }
}

// LBPhysics::macroscopic of one cell
__device__ __forceinline__ void moments_cell(const double* f, double* m /* [4] */) {
  double r = 0, a = 0, b = 0, c = 0;
#pragma unroll
  for (int q = 0; q < kQ; ++q) {
    r += f[q];
    a += f[q] * c_e[q][0];
    b += f[q] * c_e[q][1];
    c += f[q] * c_e[q][2];
  }
  m[0] = r; m[1] = a / r; m[2] = b / r; m[3] = c / r;
}

__global__ void __launch_bounds__(kLbThreads)
k_lb_collide(const FabGeom* __restrict__ gf, const FabGeom* __restrict__ gm, double* __restrict__ fi,
             const double* __restrict__ macro) {
  const FabGeom F = gf[blockIdx.y], M = gm[blockIdx.y];
  const int cells = F.vn[0] * F.vn[1] * F.vn[2];
  for (int t = blockIdx.x * kLbThreads + threadIdx.x; t < cells; t += gridDim.x * kLbThreads) {
    const long long cf = valid_cell(F, t), cm = valid_cell(M, t);
    const double rho = macro[cm], u0 = macro[cm + M.cs], u1 = macro[cm + 2 * M.cs], u2 = macro[cm + 3 * M.cs];
    double f[kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) f[q] = fi[cf + q * F.cs];
    collide_cell(f, rho, u0, u1, u2);
#pragma unroll
    for (int q = 0; q < kQ; ++q) fi[cf + q * F.cs] = f[q];
  }
}

__global__ void __launch_bounds__(kLbThreads)
k_lb_macroscopic(const FabGeom* __restrict__ gf, const FabGeom* __restrict__ gm, const double* __restrict__ fi,
                 double* __restrict__ macro) {
  const FabGeom F = gf[blockIdx.y], M = gm[blockIdx.y];
  const int cells = F.vn[0] * F.vn[1] * F.vn[2];
  for (int t = blockIdx.x * kLbThreads + threadIdx.x; t < cells; t += gridDim.x * kLbThreads) {
    const long long cf = valid_cell(F, t), cm = valid_cell(M, t);
    double f[kQ], m[kMacro];
#pragma unroll
    for (int q = 0; q < kQ; ++q) f[q] = fi[cf + q * F.cs];
    moments_cell(f, m);
#pragma unroll
    for (int d = 0; d < kMacro; ++d) macro[cm + d * M.cs] = m[d];
  }
}

// Fused stream + macroscopic (+ the next step's collision): every valid cell PULLS f_q from the
// cell at -e_q of the exchanged, bounce-back-filled level `cur` (LBPatch::stream's shifted copy,
// read side), writes it to `nxt`, accumulates the moments in the reference's order and stores them;
// with COLLIDE the collision of the following step is applied before the store, so that the value
// crosses HBM once per step instead of three times.
template <bool COLLIDE>
__global__ void __launch_bounds__(kLbThreads)
k_lb_stream_moments(const FabGeom* __restrict__ gf, const FabGeom* __restrict__ gm, const double* __restrict__ cur,
                    double* __restrict__ nxt, double* __restrict__ macro) {
  const FabGeom F = gf[blockIdx.y], M = gm[blockIdx.y];
  const int cells = F.vn[0] * F.vn[1] * F.vn[2];
  const int ox = F.vlo[0] - F.lo[0], oy = F.vlo[1] - F.lo[1], oz = F.vlo[2] - F.lo[2];
  for (int t = blockIdx.x * kLbThreads + threadIdx.x; t < cells; t += gridDim.x * kLbThreads) {
    const int i = t % F.vn[0];
    const int r = t / F.vn[0];
    const int j = r % F.vn[1], k = r / F.vn[1];
    double f[kQ], m[kMacro];
#pragma unroll
    for (int q = 0; q < kQ; ++q)
      f[q] = cur[fab_index(F, i + ox - c_e[q][0], j + oy - c_e[q][1], k + oz - c_e[q][2], q)];
    moments_cell(f, m);
    const long long cm = fab_index(M, i + ox, j + oy, k + oz, 0);
#pragma unroll
    for (int d = 0; d < kMacro; ++d) macro[cm + d * M.cs] = m[d];
    if (COLLIDE) collide_cell(f, m[0], m[1], m[2], m[3]);
    const long long cf = fab_index(F, i + ox, j + oy, k + oz, 0);
#pragma unroll
    for (int q = 0; q < kQ; ++q) nxt[cf + q * F.cs] = f[q];
  }
}

int check_lb_levels(const sgc_level_s* fi, const sgc_level_s* macro, const char* who) {
  if (!fi || !macro) return fail(SGC_ERR_ARG, std::string(who) + ": null level");
  if (fi->esz != 8 || macro->esz != 8) return fail(SGC_ERR_UNSUPPORTED, std::string(who) + ": levels must hold 8-byte Real");
  if (fi->ncomp != kQ || macro->ncomp != kMacro)
    return fail(SGC_ERR_ARG, std::string(who) + ": expects a 19-component distribution level and a 4-component macroscopic level");
  if (fi->nbox != macro->nbox || memcmp(fi->boxes.data(), macro->boxes.data(), sizeof(IBox) * (size_t)fi->nbox) != 0)
    return fail(SGC_ERR_GEOMETRY, std::string(who) + ": levels are on different boxes");
  return 0;
}

dim3 lb_grid(const sgc_level_s* l) {
  long long mx = 1;
  for (const FabGeom& g : l->geom) mx = std::max(mx, (long long)g.vn[0] * g.vn[1] * g.vn[2]);
  return dim3((unsigned)std::min<long long>((mx + kLbThreads - 1) / kLbThreads, 4096), (unsigned)l->nbox, 1);
}

}  // namespace
}  // namespace sgc

using namespace sgc;

extern "C" {

int sgc_lb_collide(sgc_level_t fi, sgc_level_t macro) {
  SGC_REQUIRE_RT();
  int st = check_lb_levels(fi, macro, "sgc_lb_collide");
  if (st || fi->nbox == 0) return st;
  k_lb_collide<<<lb_grid(fi), kLbThreads, 0, rt().compute>>>(fi->d_geom, macro->d_geom, (double*)fi->d_arena,
                                                            (const double*)macro->d_arena);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

int sgc_lb_macroscopic(sgc_level_t macro, sgc_level_t fi) {
  SGC_REQUIRE_RT();
  int st = check_lb_levels(fi, macro, "sgc_lb_macroscopic");
  if (st || fi->nbox == 0) return st;
  k_lb_macroscopic<<<lb_grid(fi), kLbThreads, 0, rt().compute>>>(fi->d_geom, macro->d_geom, (const double*)fi->d_arena,
                                                                (double*)macro->d_arena);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

int sgc_lb_copy_ops(sgc_level_t fi, const int dlo[3], const int dhi[3], int which, sgc_copy_op* ops, int cap) {
  if (!fi || !dlo || !dhi || which < 0 || which > 1) return fail(SGC_ERR_ARG, "sgc_lb_copy_ops: bad argument");
  if (fi->ncomp != kQ || fi->ng < 1) return fail(SGC_ERR_ARG, "sgc_lb_copy_ops: needs a 19-component level with ghost cells");
  std::vector<sgc_copy_op> v;
  static const int top[5][2] = {{6, 5}, {13, 12}, {14, 11}, {17, 16}, {18, 15}};   // source, destination
  for (int b = 0; b < fi->nbox; ++b) {
    const IBox& box = fi->boxes[b];
    if (which == 0) {
      // LBPatch::stream: prev(box, q) = curr(box shifted by e[opposite(q)], q)
      for (int q = 0; q < kQ; ++q) {
        sgc_copy_op o;
        memset(&o, 0, sizeof o);
        o.dst_box = o.src_box = b;
        o.dst_comp = o.src_comp = q;
        o.ncomp = 1;
        o.comp_flags = ~0u;
        for (int d = 0; d < 3; ++d) {
          o.dst_lo[d] = box.lo[d]; o.dst_hi[d] = box.hi[d];
          o.src_lo[d] = box.lo[d] + h_e[h_opp[q]][d];
        }
        v.push_back(o);
      }
    } else {
      // LBLevel.cpp:65-112: top boxes reflect 6,13,14,17,18 into the ghost layer above, ELSE bottom
      // boxes reflect 5,12,11,16,15 into the layer below
      int side = 0;
      if (box.hi[2] == dhi[2]) side = 1; else if (box.lo[2] == dlo[2]) side = -1;
      if (!side) continue;
      const int z = side > 0 ? box.hi[2] : box.lo[2];
      for (int m = 0; m < 5; ++m) {
        const int sc = side > 0 ? top[m][0] : top[m][1], dc = side > 0 ? top[m][1] : top[m][0];
        sgc_copy_op o;
        memset(&o, 0, sizeof o);
        o.dst_box = o.src_box = b;
        o.dst_comp = dc; o.src_comp = sc;
        o.ncomp = 1;
        o.comp_flags = ~0u;
        for (int d = 0; d < 3; ++d) {
          const int lo = d == 2 ? z : box.lo[d], hi = d == 2 ? z : box.hi[d];
          o.src_lo[d] = lo;
          o.dst_lo[d] = lo + h_e[sc][d]; o.dst_hi[d] = hi + h_e[sc][d];
        }
        v.push_back(o);
      }
    }
  }
  if (ops) {
    if ((int)v.size() > cap) return fail(SGC_ERR_ARG, "sgc_lb_copy_ops: buffer too small");
    if (!v.empty()) memcpy(ops, v.data(), sizeof(sgc_copy_op) * v.size());
  }
  return (int)v.size();
}

int sgc_lb_run(sgc_level_t fi, sgc_level_t prev, sgc_level_t macro, sgc_plan_t exchange, sgc_copyset_t bounce,
               sgc_copyset_t stream, int nstep, int fuse, int* result_in_prev) {
  SGC_REQUIRE_RT();
  int st = check_lb_levels(fi, macro, "sgc_lb_run");
  if (!st) st = check_lb_levels(prev, macro, "sgc_lb_run");
  if (st) return st;
  if (!exchange || !bounce || nstep < 0 || (!fuse && !stream)) return fail(SGC_ERR_ARG, "sgc_lb_run: bad argument");
  sgc_level_s* cur = fi;
  sgc_level_s* nxt = prev;
  const dim3 grid = lb_grid(fi);
  for (int s = 0; s < nstep; ++s) {
    // the collision of step s: stand-alone for the first step, otherwise already applied by the
    // fused kernel of step s-1
    if (!fuse || s == 0) { st = sgc_lb_collide(cur, macro); if (st) return st; }
    st = sgc_exchange(cur, exchange);
    if (!st) st = sgc_copyset_run(bounce, cur, cur);
    if (st) return st;
    if (!fuse) {
      st = sgc_copyset_run(stream, nxt, cur);
      if (!st) st = sgc_lb_macroscopic(macro, nxt);
      if (st) return st;
    } else if (fi->nbox > 0) {
      if (s + 1 < nstep)
        k_lb_stream_moments<true><<<grid, kLbThreads, 0, rt().compute>>>(cur->d_geom, macro->d_geom, (const double*)cur->d_arena,
                                                                         (double*)nxt->d_arena, (double*)macro->d_arena);
      else
        k_lb_stream_moments<false><<<grid, kLbThreads, 0, rt().compute>>>(cur->d_geom, macro->d_geom, (const double*)cur->d_arena,
                                                                          (double*)nxt->d_arena, (double*)macro->d_arena);
      rt().launches++;
      SGC_CUDA(cudaGetLastError());
    }
    std::swap(cur, nxt);
  }
  if (result_in_prev) *result_in_prev = (cur == prev) ? 1 : 0;
  return 0;
}

}  // extern "C"
