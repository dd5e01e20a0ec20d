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
#ifndef SGC_LB_MINB
#define SGC_LB_MINB 3            // CTAs per SM the per-cell kernels are compiled for (80 registers: measured best of 1..4)
#endif

// element offset (component 0) of valid cell number t of a fab, cells numbered x fastest
__device__ __forceinline__ long long valid_cell(const FabGeom& g, int t) {
  const int i = t % g.vn[0];
  const int r = t / g.vn[0];
  const int j = r % g.vn[1], k = r / g.vn[1];
  return fab_index(g, i + g.vlo[0] - g.lo[0], j + g.vlo[1] - g.lo[1], k + g.vlo[2] - g.lo[2], 0);
}

// ---------------------------------------------------------------------------------------------
// Correctly rounded division by a compile-time constant in three FP64 instructions.
//
// The collision divides 58 times per cell by the four constants cs2, 2*cs2, 2*cs2*cs2 and tau; an
// IEEE double division costs ~12 instructions on the FP64 pipe, which (not HBM) is what bounds this
// kernel on B200.  With rc = RN(1/c):   q = RN(x*rc);  r = x - q*c (exact, one fma);
// x/c == RN(q + r*rc).  This is correctly rounded for every normal x unless x/c lies within ~2^-53
// ulp of a rounding boundary; for a given c those x are the solutions of a linear congruence, a
// handful of mantissas, and tests/test_divconst_proof.py enumerates and checks them with exact
// rational arithmetic for each of the four constants (none fails) — so the result equals the
// reference's `/` bit for bit.  -0, subnormals, huge values, inf and nan take the real division.
// ---------------------------------------------------------------------------------------------
template <int K> struct DivConst;
template <> struct DivConst<0> { static constexpr double c = kCs2; };
template <> struct DivConst<1> { static constexpr double c = 2 * kCs2; };
template <> struct DivConst<2> { static constexpr double c = 2 * kCs2 * kCs2; };
template <> struct DivConst<3> { static constexpr double c = kTau; };

// `bad` collects, without a branch, whether any operand fell outside the range the three-instruction
// form is proven for (|x| < 2^-959 or >= 2^961, inf, nan, -0; +0 is exact: q = r = +0): the caller then
// redoes the whole cell with real divisions (EXACT).  Keeping the 58 divisions of a cell free of
// branches lets the compiler interleave their dependency chains — the kernel is bound by FP64 latency,
// not by the operation count.
template <int K, bool EXACT>
__device__ __forceinline__ double divc(double x, unsigned& bad) {
  constexpr double c = DivConst<K>::c;
  if (EXACT) return x / c;
  constexpr double rc = 1.0 / c;
  const unsigned h = (unsigned)__double2hiint(x), l = (unsigned)__double2loint(x);
  bad |= (unsigned)((((h >> 20) & 0x7ffu) - 64u >= 1920u) & ((h | l) != 0u));
  const double q = __dmul_rn(x, rc);
  const double r = __fma_rn(-q, c, x);
  return __fma_rn(r, rc, q);
}

// LBPhysics::collision for the 19 directions of one cell, in place (LBPhysics.H:5-14):
//   eu  = (u0*e0 + u1*e1) + u2*e2
//   feq = (w*rho)*( ((1 + eu/cs2) + (eu*eu)/(2*cs2*cs2)) - ((u0*u0 + u1*u1) + u2*u2)/(2*cs2) )
//   f   = (f + (feq - f)/tau) + ((3*w)*e0)*force
// evaluated with the same operations on the same values, minus the ones whose result is known:
// products with the lattice components 0 and +-1 are exact, so eu is one of +-u_a, +-(u_a + u_b),
// +-(u_a - u_b) (negation commutes with rounding; the sign of a zero eu cannot reach f), and
// eu/cs2, eu*eu/(2 cs2^2) are shared by the two directions of a pair.
template <bool EXACT>
__device__ __forceinline__ void collide_cell(double* f /* [kQ] */, double rho, double u0, double u1, double u2, unsigned& bad) {
  const double usq = divc<1, EXACT>((u0 * u0 + u1 * u1) + u2 * u2, bad);
  const double wr0 = (1. / 3.) * rho, wr1 = (1. / 18.) * rho, wr2 = (1. / 36.) * rho;
  constexpr double F1 = 3 * (1. / 18.) * kForce, F2 = 3 * (1. / 36.) * kForce;   // ((3*w)*(+-1))*force
  // one direction: t = ((1 +- a) + b) - usq, then relax towards (w*rho)*t and add the body force
#define SGC_LB_RELAX(q, t, wr, force) \
  f[q] = (f[q] + divc<3, EXACT>((wr) * (t) - f[q], bad)) + (force)
  {
    const double t = (1.0 - usq);                       // eu == 0: (1 + 0) + 0
    SGC_LB_RELAX(0, t, wr0, 0.0);
  }
#define SGC_LB_PAIR(qpos, qneg, eu, wr, fpos, fneg)                      \
  {                                                                     \
    const double eu_ = (eu);                                            \
    const double a = divc<0, EXACT>(eu_, bad), b = divc<2, EXACT>(eu_ * eu_, bad); \
    const double tp = ((1.0 + a) + b) - usq, tn = ((1.0 - a) + b) - usq; \
    SGC_LB_RELAX(qpos, tp, wr, fpos);                                   \
    SGC_LB_RELAX(qneg, tn, wr, fneg);                                   \
  }
  SGC_LB_PAIR(2, 1, u0, wr1, F1, -F1)
  SGC_LB_PAIR(4, 3, u1, wr1, 0.0, 0.0)
  SGC_LB_PAIR(6, 5, u2, wr1, 0.0, 0.0)
  SGC_LB_PAIR(10, 7, u0 + u1, wr2, F2, -F2)
  SGC_LB_PAIR(8, 9, u0 - u1, wr2, F2, -F2)
  SGC_LB_PAIR(14, 11, u0 + u2, wr2, F2, -F2)
  SGC_LB_PAIR(12, 13, u0 - u2, wr2, F2, -F2)
  SGC_LB_PAIR(18, 15, u1 + u2, wr2, 0.0, 0.0)
  SGC_LB_PAIR(16, 17, u1 - u2, wr2, 0.0, 0.0)
#undef SGC_LB_PAIR
#undef SGC_LB_RELAX
}

// LBPhysics::macroscopic of one cell (LBPhysics.H:17-35): the four sums run over q = 0..18 in order,
// starting from 0; terms multiplied by a zero lattice component add +-0 to a sum that is never -0
// and are skipped, f*(+-1) is +-f.
__device__ __forceinline__ void moments_cell(const double* f, double* m /* [4] */) {
  double r = 0.0 + f[0];
#pragma unroll
  for (int q = 1; q < kQ; ++q) r += f[q];
  double a = 0.0 - f[1];
  a += f[2]; a -= f[7]; a += f[8]; a -= f[9]; a += f[10]; a -= f[11]; a += f[12]; a -= f[13]; a += f[14];
  double b = 0.0 - f[3];
  b += f[4]; b -= f[7]; b -= f[8]; b += f[9]; b += f[10]; b -= f[15]; b += f[16]; b -= f[17]; b += f[18];
  double c = 0.0 - f[5];
  c += f[6]; c -= f[11]; c -= f[12]; c += f[13]; c += f[14]; c -= f[15]; c -= f[16]; c += f[17]; c += f[18];
  m[0] = r; m[1] = a / r; m[2] = b / r; m[3] = c / r;
}

// device check of divc against the real division (tests)
template <int K>
__global__ void k_divc_check(const double* __restrict__ x, double* __restrict__ fast, double* __restrict__ exact, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned bad = 0;
  const double v = divc<K, false>(x[i], bad);
  fast[i] = bad ? divc<K, true>(x[i], bad) : v;     // what a cell with this operand ends up with
  exact[i] = x[i] / DivConst<K>::c;
}

__global__ void __launch_bounds__(kLbThreads, SGC_LB_MINB)
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
    unsigned bad = 0;
    collide_cell<false>(f, rho, u0, u1, u2, bad);
    if (bad) {                                   // an operand outside the proven range: real divisions
#pragma unroll
      for (int q = 0; q < kQ; ++q) f[q] = fi[cf + q * F.cs];
      collide_cell<true>(f, rho, u0, u1, u2, bad);
    }
#pragma unroll
    for (int q = 0; q < kQ; ++q) fi[cf + q * F.cs] = f[q];
  }
}

__global__ void __launch_bounds__(kLbThreads, SGC_LB_MINB)
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
__global__ void __launch_bounds__(kLbThreads, SGC_LB_MINB)
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
    // source of direction q = the cell at -e_q.  With one ghost column per side it lies in the main block at
    // a fixed offset from this cell, except across the box's x faces (first / last column and e_x = +-1),
    // where it is in the x-ghost strip; the lattice vectors are compile-time constants after unrolling, so
    // the 9 directions with e_x = 0 cost one add each (the generic fab_index costs ~10 instructions).
    const long long c0 = F.off + ((long long)(k + oz) * F.n[1] + (j + oy)) * F.mx + (i + ox - F.ngx);
    const bool fast = F.ngx == 1 && ox == 1;
    const bool xlo = i == 0, xhi = i == F.vn[0] - 1;
    auto source = [&](int q) -> long long {
      constexpr int E[kQ][3] = {{0, 0, 0},   {-1, 0, 0}, {1, 0, 0},  {0, -1, 0}, {0, 1, 0},   {0, 0, -1}, {0, 0, 1},
                                {-1, -1, 0}, {1, -1, 0}, {-1, 1, 0}, {1, 1, 0},  {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1},
                                {1, 0, 1},   {0, -1, -1}, {0, 1, -1}, {0, -1, 1}, {0, 1, 1}};     // = c_e
      const int e0 = E[q][0], e1 = E[q][1], e2 = E[q][2];
      if (!fast) return fab_index(F, i + ox - e0, j + oy - e1, k + oz - e2, q);
      long long src = c0 + q * F.cs - (e0 + (long long)e1 * F.mx + (long long)e2 * F.mx * F.n[1]);
      if (e0 != 0 && (e0 > 0 ? xlo : xhi))
        src = F.xoff + q * F.xcs + ((long long)(e0 > 0 ? 0 : F.n[2]) + (k + oz - e2)) * F.n[1] + (j + oy - e1);
      return src;
    };
#pragma unroll
    for (int q = 0; q < kQ; ++q) f[q] = cur[source(q)];
    moments_cell(f, m);
    const long long cm = fab_index(M, i + ox, j + oy, k + oz, 0);
#pragma unroll
    for (int d = 0; d < kMacro; ++d) macro[cm + d * M.cs] = m[d];
    if (COLLIDE) {
      unsigned bad = 0;
      collide_cell<false>(f, m[0], m[1], m[2], m[3], bad);
      if (bad) {                                 // an operand outside the proven range: real divisions
#pragma unroll
        for (int q = 0; q < kQ; ++q) f[q] = cur[source(q)];
        collide_cell<true>(f, m[0], m[1], m[2], m[3], bad);
      }
    }
    const long long cf = fast ? c0 : fab_index(F, i + ox, j + oy, k + oz, 0);
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

int sgc_lb_selftest_divconst(int which, const double* x, double* fast, double* exact, long long n, double* constant) {
  SGC_REQUIRE_RT();
  if (which < 0 || which > 3 || !x || !fast || !exact || n < 1) return fail(SGC_ERR_ARG, "sgc_lb_selftest_divconst: bad argument");
  double *dx = nullptr, *df = nullptr, *de = nullptr;
  SGC_CUDA(cudaMalloc(&dx, sizeof(double) * n));
  SGC_CUDA(cudaMalloc(&df, sizeof(double) * n));
  SGC_CUDA(cudaMalloc(&de, sizeof(double) * n));
  SGC_CUDA(cudaMemcpy(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice));
  const unsigned blocks = (unsigned)((n + 255) / 256);
  switch (which) {
    case 0: k_divc_check<0><<<blocks, 256, 0, rt().compute>>>(dx, df, de, n); if (constant) *constant = DivConst<0>::c; break;
    case 1: k_divc_check<1><<<blocks, 256, 0, rt().compute>>>(dx, df, de, n); if (constant) *constant = DivConst<1>::c; break;
    case 2: k_divc_check<2><<<blocks, 256, 0, rt().compute>>>(dx, df, de, n); if (constant) *constant = DivConst<2>::c; break;
    default: k_divc_check<3><<<blocks, 256, 0, rt().compute>>>(dx, df, de, n); if (constant) *constant = DivConst<3>::c; break;
  }
  cudaError_t e = cudaStreamSynchronize(rt().compute);
  if (e == cudaSuccess) e = cudaMemcpy(fast, df, sizeof(double) * n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(exact, de, sizeof(double) * n, cudaMemcpyDeviceToHost);
  cudaFree(dx); cudaFree(df); cudaFree(de);
  if (e != cudaSuccess) return fail(SGC_ERR_CUDA, cudaGetErrorString(e));
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
  auto step = [&](int s) -> int {
    // the collision of step s: stand-alone for the first step, otherwise already applied by the
    // fused kernel of step s-1
    int st = 0;
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
    return 0;
  };
  int s = 0;
  // The shipped configuration is 64 x 32 x 32 cells: a step is three ~10 us launches, i.e. launch latency.
  // Steps 1 .. nstep-2 are identical up to the ping-pong, so two of them are captured and replayed as a graph.
  if (nstep >= 12 && step_graph_ok(exchange->nremote)) {
    for (; s < 3; ++s) { st = step(s); if (st) return st; }       // step 0 (with its own collision) + one pair
    StepGraph g;
    st = g.begin();
    for (int k = 0; k < 2 && !st; ++k) st = step(s + k);           // (s + k + 1 < nstep: the <true> kernel)
    const int st2 = g.end();
    if (st || st2) return st ? st : st2;
    const int times = (nstep - 1 - s) / 2;                         // leaves at least the last step
    st = g.replay(times);
    if (st) return st;
    s += 2 * times;
  }
  for (; s < nstep; ++s) { st = step(s); if (st) return st; }
  if (result_in_prev) *result_in_prev = (cur == prev) ? 1 : 0;
  return 0;
}

}  // extern "C"
