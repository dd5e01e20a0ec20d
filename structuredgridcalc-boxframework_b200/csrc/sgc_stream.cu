// sgc_stream.cu — the bandwidth-tuned 7-point sweep for uniform levels (BASELINE configs 2-5).
//
// Data model exploited: all fabs of a level are contiguous in one arena, each in the reference's
// Fortran order with its ghost shell INSIDE the fab (LevelData.H:266-271).  A z-plane of a fab
// (n0*n1 doubles, halo rows and columns included) is therefore one contiguous run of bytes, and
// the whole level is one uniform sequence of planes: plane q = (box q / n2, fab plane q % n2).
//
// Kernel shape (persistent, one CTA per SM):
//   * the plane sequence of the launch is cut into `gridDim.x` contiguous ranges of stages
//     (a stage = PS planes); CTA c streams its range through a ring of NS stages in shared memory;
//   * loads are 1-D bulk async copies (cp.async.bulk, the TMA engine: SASS UBLKCP) of one whole
//     stage each, completion signalled on an mbarrier per ring slot; NS-3 stages (~100 KB) are in
//     flight per SM while three are being read, which covers HBM latency at full bandwidth
//     without tying up registers;
//   * every thread owns CPT columns (x,y) of the box and marches along z: the centre and z-1
//     values roll through registers, z+1 / x±1 / y±1 come from shared memory (conflict-free:
//     lanes of a warp read consecutive doubles); results go straight to global memory, a warp
//     storing 32 consecutive doubles of one row;
//   * ghost planes are loaded (they carry the z halo) but produce no output, so crossing from one
//     fab to the next needs no special case; a range that starts or ends inside a fab re-reads one
//     stage of its neighbour range (2 x 148 planes per launch, < 1 % extra traffic).
//
// HBM traffic per cell-update: (n0*n1*n2)/(vnx*vny*vnz)*8 B read (66^3/64^3 -> 8.77 B) + 8 B
// written; the algorithmic figure used for the roofline is 16 B.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "sgc_device.h"

namespace sgc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SGC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SGC_DONE_%=;\n"
      "bra SGC_WAIT_%=;\n"
      "SGC_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy through the async proxy (TMA), completing `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ double sd2(double p, double twoc, double m) {   // (p - 2c) + m
  return __dadd_rn(__dsub_rn(p, twoc), m);
}

template <int VNX, int VNY, int NG, int PS, int NS, int NT>
__global__ void __launch_bounds__(NT, 1)
k_heat_stream(const double* __restrict__ in, double* __restrict__ out, int nstages, int n2, double f, int fma) {
  constexpr int N0 = VNX + 2 * NG, N1 = VNY + 2 * NG;
  constexpr int PLANE = N0 * N1;                 // elements per plane
  constexpr int STAGE = PS * PLANE;              // elements per stage
  constexpr uint32_t STAGE_BYTES = STAGE * 8u;
  constexpr int RING_PLANES = NS * PS;
  constexpr int COLS = VNX * VNY;
  constexpr int CPT = (COLS + NT - 1) / NT;      // columns per thread
  static_assert(STAGE_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES);

  const int tid = threadIdx.x;
  const int s_begin = (int)((long long)blockIdx.x * nstages / gridDim.x);
  const int s_end = (int)((long long)(blockIdx.x + 1) * nstages / gridDim.x);
  if (s_begin >= s_end) return;
  const int l_begin = s_begin > 0 ? s_begin - 1 : 0;
  const int l_end = s_end < nstages ? s_end + 1 : nstages;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(full + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int next_load = l_begin;       // meaningful in thread 0 only
  auto issue = [&](int t) {
    const int slot = t % NS;
    mbar_expect_tx(full + slot, STAGE_BYTES);
    bulk_g2s(ring + (size_t)slot * STAGE, in + (size_t)t * STAGE, STAGE_BYTES, full + slot);
  };
  if (tid == 0) {
    const int stop = l_begin + NS < l_end ? l_begin + NS : l_end;
    for (; next_load < stop; ++next_load) issue(next_load);
  }

  int soff[CPT];
#pragma unroll
  for (int m = 0; m < CPT; ++m) {
    const int idx = tid + m * NT;
    const int i = idx % VNX, j = idx / VNX;
    soff[m] = (idx < COLS) ? (j + NG) * N0 + (i + NG) : -1;
  }
  double cc[CPT], zm[CPT];
  bool primed = false;
  uint32_t phase = 0;
  int waited = l_begin - 1;

  for (int s = s_begin; s < s_end; ++s) {
    const int need = (s + 1 < l_end) ? s + 1 : l_end - 1;
    while (waited < need) {
      ++waited;
      const int slot = waited % NS;
      mbar_wait(full + slot, (phase >> slot) & 1u);
      phase ^= 1u << slot;
    }
#pragma unroll
    for (int p = 0; p < PS; ++p) {
      const int q = s * PS + p;                 // plane index within this launch
      const int kz = q % n2;
      if (kz < NG || kz >= n2 - NG) { primed = false; continue; }
      const double* __restrict__ pc = ring + (size_t)(q % RING_PLANES) * PLANE;
      const double* __restrict__ pm = ring + (size_t)((q - 1) % RING_PLANES) * PLANE;
      const double* __restrict__ pp = ring + (size_t)((q + 1) % RING_PLANES) * PLANE;
      double* __restrict__ po = out + (size_t)q * PLANE;
      if (!primed) {
#pragma unroll
        for (int m = 0; m < CPT; ++m)
          if (soff[m] >= 0) { zm[m] = pm[soff[m]]; cc[m] = pc[soff[m]]; }
        primed = true;
      }
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        if (soff[m] < 0) continue;
        const int o = soff[m];
        const double c = cc[m];
        const double zp = pp[o];
        const double twoc = __dmul_rn(2.0, c);
        const double tx = sd2(pc[o + 1], twoc, pc[o - 1]);
        const double ty = sd2(pc[o + N0], twoc, pc[o - N0]);
        const double tz = sd2(zp, twoc, zm[m]);
        const double S = __dadd_rn(__dadd_rn(tx, ty), tz);
        po[o] = fma ? __fma_rn(f, S, c) : __dadd_rn(c, __dmul_rn(f, S));
        zm[m] = c;
        cc[m] = zp;
      }
    }
    __syncthreads();            // every warp is done with stage s-1: its slot may be refilled
    if (tid == 0) {
      while (next_load < l_end && next_load - NS <= s - 1) { issue(next_load); ++next_load; }
    }
  }
}

template <int VNX, int VNY, int NG, int PS, int NS, int NT>
static int launch_cfg(const double* in, double* out, int nstages, int n2, double f, int fma, cudaStream_t s) {
  constexpr size_t smem = (size_t)NS * PS * (VNX + 2 * NG) * (VNY + 2 * NG) * 8 + NS * 8;
  static bool configured = false;
  auto kern = k_heat_stream<VNX, VNY, NG, PS, NS, NT>;
  if (!configured) {
    SGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = rt().sm_count;
  if (grid > nstages) grid = nstages;
  kern<<<grid, NT, smem, s>>>(in, out, nstages, n2, f, fma);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// returns 1 when the geometry is not covered (caller falls back to the table-driven kernel)
int launch_heat_stream(sgc_level_s* unew, sgc_level_s* u, double f, int fma, int box_begin, int box_end,
                       cudaStream_t s) {
  if (!u->uniform || !unew->uniform || u->ncomp != 1 || unew->ncomp != 1 || u->ng != unew->ng || u->ng != 1) return 1;
  const FabGeom& g = u->geom[0];
  if (memcmp(g.n, unew->geom[0].n, sizeof g.n) != 0) return 1;
  const int vnx = g.vn[0], vny = g.vn[1], n2 = g.n[2];
  const long long fab = g.cs;
  const double* in = (const double*)u->d_arena + (long long)box_begin * fab;
  double* out = (double*)unew->d_arena + (long long)box_begin * fab;
  const long long planes = (long long)(box_end - box_begin) * n2;
  if (planes > 0x7fffffffLL) return 1;
  if (getenv("SGC_NO_STREAM")) return 1;
#define SGC_STREAM_CASE(VX, VY, PS_, NS_, NT_)                                                       \
  if (vnx == VX && vny == VY && n2 % PS_ == 0)                                                       \
    return launch_cfg<VX, VY, 1, PS_, NS_, NT_>(in, out, (int)(planes / PS_), n2, f, fma, s)
  SGC_STREAM_CASE(64, 64, 1, 6, 512);
  SGC_STREAM_CASE(32, 32, 2, 10, 512);
  SGC_STREAM_CASE(16, 16, 6, 12, 256);
  SGC_STREAM_CASE(8, 8, 10, 24, 64);
#undef SGC_STREAM_CASE
  return 1;
}

}  // namespace sgc
