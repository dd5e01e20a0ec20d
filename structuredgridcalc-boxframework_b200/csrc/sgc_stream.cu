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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// NCW consumer warps + 1 producer warp.  Consumer warps never synchronise with each other: a slot
// of the ring is handed back to the producer through an `empty` mbarrier that counts one arrival
// per consumer warp, so warps drift apart by up to NS-2 stages and stalls of one warp (store
// back-pressure, a late plane) do not stop the others.
// MODE: 0 = the sweep.  1..3 are roofline diagnostics of the same pipeline (env SGC_STREAM_MODE,
// never used by the product path): 1 = plain copy of the centre value, 2 = loads only, 3 = stores only.
template <int VNX, int VNY, int NG, int PS, int NS, int NCW, int MODE = 0>
__global__ void __launch_bounds__((NCW + 1) * 32, 1)
k_heat_stream(const double* __restrict__ in, double* __restrict__ out, int nstages, int n2, double f, int fma) {
  constexpr int NT = NCW * 32;                   // consumer threads
  constexpr int N0 = VNX + 2 * NG, N1 = VNY + 2 * NG;
  constexpr int PLANE = N0 * N1;                 // elements per plane
  constexpr int STAGE = PS * PLANE;              // elements per stage
  constexpr uint32_t STAGE_BYTES = STAGE * 8u;
  constexpr int COLS = VNX * VNY;
  constexpr int CPT = (COLS + NT - 1) / NT;      // columns per consumer thread
  static_assert(STAGE_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(NS >= 3, "two stages are being read while at least one is in flight");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE_BYTES);
  uint64_t* empty = full + NS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s_begin = (int)((long long)blockIdx.x * nstages / gridDim.x);
  const int s_end = (int)((long long)(blockIdx.x + 1) * nstages / gridDim.x);
  if (s_begin >= s_end) return;
  // stages to load: one before and one after the range (z halo of the first / last output plane)
  const int l_begin = s_begin > 0 ? s_begin - 1 : 0;
  const int l_end = s_end < nstages ? s_end + 1 : nstages;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    // ---------------- producer: one lane keeps the ring full ----------------
    if (lane == 0) {
      int slot = 0, round = 0;                   // ring slot of stage t is (t - l_begin) % NS
      for (int t = l_begin; t < l_end; ++t) {
        if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
        if (MODE == 3) { mbar_arrive(full + slot); if (++slot == NS) { slot = 0; ++round; } continue; }
        mbar_expect_tx(full + slot, STAGE_BYTES);
        bulk_g2s(ring + (size_t)slot * STAGE, in + (size_t)t * STAGE, STAGE_BYTES, full + slot);
        if (++slot == NS) { slot = 0; ++round; }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  int soff[CPT];
#pragma unroll
  for (int m = 0; m < CPT; ++m) {
    const int idx = tid + m * NT;
    const int i = idx % VNX, j = idx / VNX;
    soff[m] = (idx < COLS) ? (j + NG) * N0 + (i + NG) : -1;
  }
  const int p_lo = s_begin * PS, p_hi = s_end * PS;   // planes whose output this CTA owns
  int q = l_begin * PS;                                // plane being read as the "z+1" plane
  int kz = q % n2;                                     // its index inside its fab
  int slot = 0, prev_slot = 0;
  uint32_t phase = 0;
  double cc[CPT], zm[CPT];                             // column values at planes q-1 and q-2
  const double* prev = ring;
  bool prev_out = false;

  for (int t = l_begin; t < l_end; ++t) {
    mbar_wait(full + slot, (phase >> slot) & 1u);
    phase ^= 1u << slot;
#pragma unroll
    for (int p = 0; p < PS; ++p) {
      const double* __restrict__ cur = ring + (size_t)(slot * PS + p) * PLANE;
      const bool emit = prev_out && (q - 1 >= p_lo) && (q - 1 < p_hi);
      double* __restrict__ po = out + (size_t)(q - 1) * PLANE;
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        if (soff[m] < 0) continue;
        const int o = soff[m];
        const double zp = cur[o];
        if (MODE == 1 || MODE == 3) { if (emit) po[o] = cc[m]; }
        else if (MODE == 2) { if (emit && zp == 1.2345e300) po[o] = zp; }
        else if (emit) {
          const double c = cc[m];
          const double twoc = __dmul_rn(2.0, c);
          const double tx = sd2(prev[o + 1], twoc, prev[o - 1]);
          const double ty = sd2(prev[o + N0], twoc, prev[o - N0]);
          const double tz = sd2(zp, twoc, zm[m]);
          const double S = __dadd_rn(__dadd_rn(tx, ty), tz);
          po[o] = fma ? __fma_rn(f, S, c) : __dadd_rn(c, __dmul_rn(f, S));
        }
        zm[m] = cc[m];
        cc[m] = zp;
      }
      prev = cur;
      prev_out = (kz >= NG) && (kz < n2 - NG);
      ++q;
      if (++kz == n2) kz = 0;
    }
    if (t > l_begin) {                           // stage t-1 is no longer read by this warp
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + prev_slot);
    }
    prev_slot = slot;
    if (++slot == NS) slot = 0;
  }
}

template <int VNX, int VNY, int NG, int PS, int NS, int NCW, int MODE = 0>
static int launch_cfg(const double* in, double* out, int nstages, int n2, double f, int fma, cudaStream_t s) {
  constexpr size_t smem = (size_t)NS * PS * (VNX + 2 * NG) * (VNY + 2 * NG) * 8 + 2 * NS * 8;
  static bool configured = false;
  auto kern = k_heat_stream<VNX, VNY, NG, PS, NS, NCW, MODE>;
  if (!configured) {
    SGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = rt().sm_count;
  if (grid > nstages) grid = nstages;
  kern<<<grid, (NCW + 1) * 32, smem, s>>>(in, out, nstages, n2, f, fma);
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
#define SGC_STREAM_CASE(VX, VY, PS_, NS_, NCW_)                                                      \
  if (vnx == VX && vny == VY && n2 % PS_ == 0)                                                       \
    return launch_cfg<VX, VY, 1, PS_, NS_, NCW_>(in, out, (int)(planes / PS_), n2, f, fma, s)
  if (const char* m = getenv("SGC_STREAM_MODE")) {      // diagnostics only
    if (vnx == 64 && vny == 64) {
      const int ns = (int)(planes);
      if (m[0] == '1') return launch_cfg<64, 64, 1, 1, 6, 16, 1>(in, out, ns, n2, f, fma, s);
      if (m[0] == '2') return launch_cfg<64, 64, 1, 1, 6, 16, 2>(in, out, ns, n2, f, fma, s);
      if (m[0] == '3') return launch_cfg<64, 64, 1, 1, 6, 16, 3>(in, out, ns, n2, f, fma, s);
    }
  }
  SGC_STREAM_CASE(64, 64, 1, 6, 16);
  SGC_STREAM_CASE(32, 32, 2, 10, 16);
  SGC_STREAM_CASE(16, 16, 6, 12, 8);
  SGC_STREAM_CASE(8, 8, 10, 24, 2);
#undef SGC_STREAM_CASE
  return 1;
}

}  // namespace sgc
