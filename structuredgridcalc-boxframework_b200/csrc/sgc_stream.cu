// sgc_stream.cu — the bandwidth-tuned 7-point sweep for uniform levels (BASELINE configs 2-5).
//
// Data model exploited (split-x device layout, sgc_device.h): the main blocks of all fabs of a level
// are contiguous, each a stack of z-planes of n1 dense rows of the VALID x columns, so the whole
// level is one homogeneous sequence of planes: plane q = (box q / n2, fab plane q % n2); the two
// x-ghost columns of a plane are one contiguous strip of n1 values per side in the fab's xg block.
//
// Kernel shape (persistent, one CTA per SM, warp specialised):
//   * the plane sequence of the launch is cut into `gridDim.x` contiguous ranges of stages (a stage
//     = PS planes); CTA c streams its range through a ring of NS stages in shared memory;
//   * a producer warp issues, per stage, three 1-D bulk async copies (cp.async.bulk, the TMA
//     engine: SASS UBLKCP) — the planes and the two x-ghost strips — completing on one `full`
//     mbarrier per ring slot; NS-2 stages (~140 KB) are in flight per SM, which covers HBM latency
//     at full bandwidth without tying up registers;
//   * consumer warps never synchronise with each other: every thread owns CPT columns (x,y) and
//     marches along z with the centre and z-1 values rolling through registers; z+1, x±1 and y±1
//     come from shared memory (conflict-free: lanes of a warp read consecutive doubles); a ring
//     slot goes back to the producer through an `empty` mbarrier counting one arrival per warp;
//   * results go straight to global memory as dense rows: a warp stores 32 consecutive doubles,
//     rows are 512 B aligned for a 64^3 box, so every DRAM sector is written whole (the reference
//     layout would leave two partial sectors per row: 2.5 vs 5.3 TB/s, profiles/micro);
//   * ghost planes are loaded (they carry the z halo) but produce no output, so crossing from one
//     fab to the next needs no special case; a range that starts or ends inside a fab re-reads one
//     stage of its neighbour range (2 x 148 planes per launch, < 1 % extra traffic).
//
// HBM traffic per cell-update: n1*n2*(vnx+2)/(vnx*vny*vnz)*8 B read (66^3/64^3 -> 8.77 B) + 8 B
// written; the algorithmic figure used for the roofline is 16 B.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "sgc_device.h"
#include "sgc_pipe.cuh"

#ifndef SGC_NCW64
#define SGC_NCW64 16     // consumer warps of the 64x64 instance
#endif

namespace sgc {

constexpr int kGatherWarps = 4;   // helper warps of the narrow-box fused mode (GATHER)

// Narrow-box instances have few consumer warps per CTA (one or four columns per thread): when the ring
// is small enough two CTAs share an SM, which doubles the warps that hide shared-memory / FP64 latency.
__host__ __device__ constexpr size_t stream_smem_bytes(int vnx, int vny, int ng, int ps, int ns, int op, int esz = 8) {
  return (size_t)ns * ps * ((size_t)vnx * (vny + 2 * ng) + 2 * (vny + 2 * ng) + (op == 1 ? (size_t)vnx * vny : 0)) * esz +
         2 * (size_t)ns * 8;
}
__host__ __device__ constexpr int stream_ctas_per_sm(int vnx, int vny, int ng, int ps, int ns, int op, int esz = 8) {
  const size_t per_cta = stream_smem_bytes(vnx, vny, ng, ps, ns, op, esz) + 1024;     // + the per-CTA reservation
  return per_cta <= 74 * 1024 ? 3 : (per_cta <= 112 * 1024 ? 2 : 1);
}


// NCW consumer warps + 1 producer warp.  Consumer warps never synchronise with each other: a slot
// of the ring is handed back to the producer through an `empty` mbarrier that counts one arrival
// per consumer warp, so warps drift apart by up to NS-2 stages and stalls of one warp (store
// back-pressure, a late plane) do not stop the others.
// MODE: 0 = the sweep.  1..3 are roofline diagnostics of the same pipeline (env SGC_STREAM_MODE,
// never used by the product path): 1 = plain copy of the centre value, 2 = loads only, 3 = stores only.
//
// PUSH = fused ghost fill ("pull y/z, push x"): inside a step loop the separate exchange launch
// disappears.
//   * y and z halos are PULLED by the producer: instead of a fab's own ghost rows / ghost planes it
//     bulk-copies the neighbouring fab's first / last valid row / plane (dense, whole sectors) into
//     the same place of the ring slot — zero work for the consumer warps;
//   * x halos are PUSHED by the consumers: the thread that computes the first / last cell of a row
//     (one lane per warp) also stores it into the x neighbour's ghost strip, one predicated store
//     per cell from the register that holds the result.  (Pulling x would read one 8-byte value per
//     512-byte row, which costs a 128-byte DRAM access each: ncu shows 583 MB read for 104 MB of
//     ghost cells in the stand-alone exchange kernel.)
// Where the plan has no local motion item (push.nbr == -1: physical boundary or a box owned by
// another GPU) the fab's own ghost cells are used, exactly as without PUSH; they are kept current by
// the BC / NCCL unpack.  Edge and corner ghosts are not needed by the 7-point stencil and are left
// alone; sgc_heat_run restores the reference's ghost contents by running its last two steps unfused.

// OP: 0 = heat / Jacobi (unew = u + f*S), 1 = leap-frog wave (unp1 = 2 un - unm1 + f*Tx + f*Ty + f*Tz,
// WavePatch.cpp:226-244).  For the wave every ring slot also carries the valid rows of unm1 for the
// planes that are emitted when the slot arrives (one more bulk copy per plane; read back with LDS).
//
// GATHER (narrow boxes, VNX < 32): pushing x faces costs a partial-warp store on every 8th cell and
// pulled y rows would be 64-128 byte bulk copies, so kGatherWarps extra helper warps GATHER both
// instead: taking the stages round-robin, the 32 lanes of a gather warp load the x-halo values (last
// / first valid column of the x neighbours) and the y-ghost rows straight from the neighbouring fabs
// with ordinary loads (one batch per stage), store them into the ring slot, and arrive on the slot's
// `full` barrier next to the TMA transaction count.  With
// 16^3 boxes a CTA walks the boxes of an x row back to back, so those lines are in L2 (they were, or
// are about to be, streamed by the same CTA) and cost no extra DRAM traffic.
// T: double, or float (the reference's USE_SINGLE_PRECISION Real; plain sweep only: the fused ghost fill, the
// peer-memory halo and the edge strips are double-only).
template <int VNX, int VNY, int NG, int PS, int NS, int NCW, int MODE = 0, bool PUSH = false, int OP = 0,
          bool GATHER = false, typename T = double>
__global__ void __launch_bounds__((NCW + 1 + (GATHER ? kGatherWarps : 0)) * 32, stream_ctas_per_sm(VNX, VNY, NG, PS, NS, OP, (int)sizeof(T)))
k_heat_stream(const T* __restrict__ in_main, const T* __restrict__ in_xg, T* __restrict__ out_main,
              int nstages, int n2, double f_, int fma, int seam, int jump, PushArgs push,
              const T* __restrict__ aux_main) {
  static_assert(sizeof(T) == 8 || (!PUSH && !GATHER && OP == 0 && MODE == 0), "float: the plain heat sweep only");
  constexpr uint32_t ESZ = (uint32_t)sizeof(T);
  const T f = (T)f_;
  // the level-wide pointers of the fused modes, in the element type (never dereferenced when T is float)
  const T* const lvl_in = reinterpret_cast<const T*>(push.in_level);
  T* const lvl_out = reinterpret_cast<T*>(push.out_level);
  // The launch covers the box range as a virtual stage sequence [0, nstages); stages >= seam live
  // `jump` stages further on in memory (a second box range, e.g. the two boundary box layers of a
  // z-slab).  Both ranges start and end on fab boundaries, so the seam sits between two ghost planes.
  constexpr int NT = NCW * 32;                   // consumer threads
  constexpr int N1 = VNY + 2 * NG;
  constexpr int PLANE = VNX * N1;                // elements per main plane (dense rows, y ghosts included)
  constexpr int STRIP = N1;                      // elements per x-ghost strip of one plane and side: the ghost column NEXT to
                                                 // the valid ones (xg layer NG-1 of the low side, layer 0 of the high side)
  constexpr int AUXP = (OP == 1) ? VNX * VNY : 0;   // elements of the third level carried per plane (valid rows)
  constexpr int STAGE = PS * (PLANE + 2 * STRIP + AUXP);   // ring slot: planes | low strips | high strips | aux rows
  constexpr int AUX_OFF = PS * (PLANE + 2 * STRIP);
  constexpr uint32_t MAIN_BYTES = PS * PLANE * ESZ, STRIP_BYTES = PS * STRIP * ESZ, AUX_BYTES = PS * AUXP * ESZ;
  constexpr int COLS = VNX * VNY;
  constexpr int CPT = (COLS + NT - 1) / NT;      // columns per consumer thread
  static_assert(NG == 1 || (NG == 2 && !PUSH && !GATHER && OP == 0), "wider ghost layers: the plain heat sweep only");
  static_assert(MAIN_BYTES % 16 == 0 && STRIP_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(NT % VNX == 0, "a thread keeps the same x for all its columns");
  static_assert(NS >= 3, "two stages are being read while at least one is in flight");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* ring = reinterpret_cast<T*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * STAGE * ESZ);
  uint64_t* empty = full + NS;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s_begin = (int)((long long)blockIdx.x * nstages / gridDim.x);
  const int s_end = (int)((long long)(blockIdx.x + 1) * nstages / gridDim.x);
  if (s_begin >= s_end) return;
  if (push.err && *reinterpret_cast<const volatile unsigned int*>(push.err)) return;   // an earlier wait of this phase timed out
  // stages to load: one before and one after the range (z halo of the first / last output plane)
  const int l_begin = s_begin > 0 ? s_begin - 1 : 0;
  const int l_end = s_end < nstages ? s_end + 1 : nstages;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, GATHER ? 2 : 1); mbar_init(empty + i, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == NCW) {
    // ---------------- producer: one lane keeps the ring full ----------------
    if (lane == 0) {
      int slot = 0, round = 0;                   // ring slot of stage t is (t - l_begin) % NS
      long long tr = l_begin + (l_begin >= seam ? jump : 0);   // real stage index
      const long long xfab = (long long)2 * NG * STRIP * n2;   // xg block elements per fab: [side][layer][k][j]
      int pb = -1, dym = -1, dyp = -1, dzm = -1, dzp = -1;     // cached pull sources of box pb
      int rzm = -1, rzp = -1;                                  // remote z neighbours of box pb (peer memory)
      bool waited[2] = {false, false};
      // where ghost plane kz of box b comes from: the z neighbour's last (low side) / first (high side)
      // valid plane — a local fab, a fab in the neighbouring GPU's memory, or (no neighbour) the own ghosts
      auto zsource = [&](int b, int kz) -> const T* {
        const long long fabm = (long long)PLANE * n2;
        const int vnz_ = n2 - 2 * NG;
        const bool low = kz < NG;
        const int dl = low ? dzm : dzp, rr = low ? rzm : rzp;
        const long long skz = low ? kz + vnz_ : kz - vnz_;
        if (dl >= 0) return lvl_in + dl * fabm + skz * PLANE;
        if (rr >= 0) {
          const int ps = rr >> 28;
          if (!waited[ps]) {           // the peer must have finished writing the level we read
            if (!wait_peer_counter(push.peer_flag[ps], push.wait_value, push.wait_budget_ns)) *push.err = 1u;   // report, never hang
            __threadfence_system();
            asm volatile("fence.proxy.async;" ::: "memory");
            waited[ps] = true;
          }
          return reinterpret_cast<const T*>(push.peer_in[ps]) + (long long)(rr & 0x0fffffff) * fabm + skz * PLANE;
        }
        return lvl_in + b * fabm + (long long)kz * PLANE;
      };
      for (int t = l_begin; t < l_end; ++t, ++tr) {
        if (t == seam && t != l_begin) tr += jump;
        if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
        if (MODE == 3) {
          mbar_arrive(full + slot);
        } else {
          const long long pl = tr * PS;                        // first plane of the stage
          const long long box = pl / n2;
          const int kz0 = (int)(pl - box * n2);
          T* dst = ring + (size_t)slot * STAGE;
          const T* xg = in_xg + box * xfab + ((long long)(NG - 1) * n2 + kz0) * STRIP;      // low side, innermost layer
          if (GATHER) {
            // planes only: a ghost plane whole (from its z source), an interior plane's valid rows;
            // the gather warp fills the y-ghost rows and the x strips
            const int b = push.box0 + (int)box;
            if (b != pb) {
              pb = b;
              const int* __restrict__ nb = push.nbr + (size_t)b * 27;
              dzm = __ldg(nb + 13 - 9); dzp = __ldg(nb + 13 + 9);
              rzm = push.rz ? __ldg(push.rz + 2 * b) : -1;
              rzp = push.rz ? __ldg(push.rz + 2 * b + 1) : -1;
            }
            const long long fabm = (long long)PLANE * n2;
            uint32_t bytes = 0;
#pragma unroll
            for (int p = 0; p < PS; ++p) bytes += (kz0 + p < NG || kz0 + p >= n2 - NG) ? PLANE * ESZ : VNY * VNX * ESZ;
            mbar_expect_tx(full + slot, bytes);
#pragma unroll
            for (int p = 0; p < PS; ++p) {
              const int kz = kz0 + p;
              T* d = dst + p * PLANE;
              const T* own = lvl_in + b * fabm + (long long)kz * PLANE;
              if (kz < NG || kz >= n2 - NG) bulk_g2s(d, zsource(b, kz), PLANE * ESZ, full + slot);
              else bulk_g2s(d + NG * VNX, own + NG * VNX, VNY * VNX * ESZ, full + slot);
            }
            if (++slot == NS) { slot = 0; ++round; }
            continue;
          }
          mbar_expect_tx(full + slot, MAIN_BYTES + 2 * STRIP_BYTES + AUX_BYTES);
          if (OP == 1) {
            // unm1 valid rows of planes pl-1 .. pl+PS-2 (the planes emitted while this stage is read);
            // the very first plane of the level has no predecessor: load plane 0 instead (never used)
#pragma unroll
            for (int p = 0; p < PS; ++p) {
              const long long ap = (pl + p > 0) ? pl + p - 1 : 0;
              bulk_g2s(dst + AUX_OFF + p * AUXP, aux_main + ap * PLANE + NG * VNX, AUXP * ESZ, full + slot);
            }
          }
          if (!PUSH || !(push.mask & 6)) {
            bulk_g2s(dst, in_main + (size_t)tr * (PS * PLANE), MAIN_BYTES, full + slot);
          } else {
            // pull: y-ghost rows and z-ghost planes come from the neighbours' valid cells
            const int b = push.box0 + (int)box;
            if (b != pb) {
              pb = b;
              const int* __restrict__ nb = push.nbr + (size_t)b * 27;
              const bool py = (push.mask & 4) != 0, pz = (push.mask & 2) != 0;
              dym = py ? __ldg(nb + 13 - 3) : -1; dyp = py ? __ldg(nb + 13 + 3) : -1;
              dzm = pz ? __ldg(nb + 13 - 9) : -1; dzp = pz ? __ldg(nb + 13 + 9) : -1;
              rzm = (pz && push.rz) ? __ldg(push.rz + 2 * b) : -1;
              rzp = (pz && push.rz) ? __ldg(push.rz + 2 * b + 1) : -1;
            }
            const long long fabm = (long long)PLANE * n2;
#pragma unroll
            for (int p = 0; p < PS; ++p) {
              const int kz = kz0 + p;
              T* d = dst + p * PLANE;
              const T* own = lvl_in + b * fabm + (long long)kz * PLANE;
              if (kz < NG || kz >= n2 - NG) {
                bulk_g2s(d, zsource(b, kz), PLANE * ESZ, full + slot);
              } else if (dym < 0 && dyp < 0) {
                bulk_g2s(d, own, PLANE * ESZ, full + slot);
              } else {
                constexpr uint32_t YG = NG * VNX * ESZ, YV = VNY * VNX * ESZ;
                const T* lo = dym >= 0 ? lvl_in + dym * fabm + (long long)kz * PLANE + VNY * VNX : own;
                const T* hi = dyp >= 0 ? lvl_in + dyp * fabm + (long long)kz * PLANE + NG * VNX
                                            : own + (VNY + NG) * VNX;
                bulk_g2s(d, lo, YG, full + slot);                              // ghost rows 0..NG-1
                bulk_g2s(d + NG * VNX, own + NG * VNX, YV, full + slot);      // valid rows
                bulk_g2s(d + (VNY + NG) * VNX, hi, YG, full + slot);          // ghost rows VNY+NG..
              }
            }
          }
          bulk_g2s(dst + PS * PLANE, xg, STRIP_BYTES, full + slot);
          bulk_g2s(dst + PS * PLANE + PS * STRIP, xg + (long long)n2 * STRIP, STRIP_BYTES, full + slot);   // high side, layer 0
        }
        if (++slot == NS) { slot = 0; ++round; }
      }
    }
    return;
  }

  if (GATHER && warp > NCW) {
    // ---------------- gather warps: x strips and y-ghost rows from the neighbouring fabs ----------------
    // kGatherWarps warps take the stages round-robin; each issues all loads of its stage in one batch
    // (they are L2 round trips) before it stores them, so a stage costs one latency, spread over 4 warps.
    const int gw = warp - NCW - 1;
    const long long fabm = (long long)PLANE * n2, xfab = (long long)2 * NG * STRIP * n2;
    int pb = -1, dxm = -1, dxp = -1, dym = -1, dyp = -1;
    constexpr int NXH = 2 * PS * VNY;              // x-halo values of a stage: (side, plane, valid row)
    constexpr int NYH = 2 * PS * VNX;              // y-ghost values of a stage: (side, plane, x)
    constexpr int IX = (NXH + 31) / 32, IY = (NYH + 31) / 32;
    for (int t = l_begin + gw; t < l_end; t += kGatherWarps) {
      const int rel = t - l_begin, slot = rel % NS, round = rel / NS;
      const long long tr = t + (t >= seam ? jump : 0);
      if (round > 0) mbar_wait(empty + slot, (round - 1) & 1u);
      const long long pl = tr * PS;
      const long long box = pl / n2;
      const int kz0 = (int)(pl - box * n2);
      const int b = push.box0 + (int)box;
      if (b != pb) {
        pb = b;
        const int* __restrict__ nb = push.nbr + (size_t)b * 27;
        dxm = __ldg(nb + 12); dxp = __ldg(nb + 14); dym = __ldg(nb + 10); dyp = __ldg(nb + 16);
      }
      T* dst = ring + (size_t)slot * STAGE;
      const T* ownmain = lvl_in + b * fabm;
      const T* ownxg = lvl_in + push.xg_base + b * xfab;
      T vx[IX], vy[IY];
      int ox[IX], oy[IY];                          // destination offsets in the slot, -1 = nothing to do
#pragma unroll
      for (int it = 0; it < IX; ++it) {
        const int idx = lane + 32 * it;
        const int side = idx / (PS * VNY), rem = idx % (PS * VNY), p = rem / VNY, r = rem % VNY + NG;
        const int kz = kz0 + p;
        const bool live = idx < NXH && kz >= NG && kz < n2 - NG;   // edge ghosts are not needed by the stencil
        const int D = side ? dxp : dxm;
        const T* src = D >= 0 ? lvl_in + D * fabm + ((long long)kz * N1 + r) * VNX + (side ? 0 : VNX - 1)
                                   : ownxg + ((long long)side * n2 + kz) * N1 + r;
        ox[it] = live ? PS * PLANE + side * PS * STRIP + p * STRIP + r : -1;
        vx[it] = live ? *src : (T)0;
      }
#pragma unroll
      for (int it = 0; it < IY; ++it) {
        const int idx = lane + 32 * it;
        const int side = idx / (PS * VNX), rem = idx % (PS * VNX), p = rem / VNX, x = rem % VNX;
        const int kz = kz0 + p;
        const bool live = idx < NYH && kz >= NG && kz < n2 - NG;   // ghost planes are copied whole by the producer
        const int D = side ? dyp : dym;
        const int grow = side ? VNY + NG : 0;      // ghost row being filled (NG == 1)
        const T* src = D >= 0 ? lvl_in + D * fabm + ((long long)kz * N1 + (side ? NG : VNY)) * VNX + x
                                   : ownmain + ((long long)kz * N1 + grow) * VNX + x;
        oy[it] = live ? p * PLANE + grow * VNX + x : -1;
        vy[it] = live ? *src : (T)0;
      }
#pragma unroll
      for (int it = 0; it < IX; ++it) if (ox[it] >= 0) dst[ox[it]] = vx[it];
#pragma unroll
      for (int it = 0; it < IY; ++it) if (oy[it] >= 0) dst[oy[it]] = vy[it];
      __syncwarp();
      if (lane == 0) mbar_arrive(full + slot);
    }
    return;
  }

  // ---------------- consumers ----------------
  // Thread tid owns the CPT columns (xi, j0 + m*ROWS), m = 0..CPT-1: the same x for all of them, so
  // every shared-memory / global offset of column m is the offset of column 0 plus a compile-time
  // constant (m*NT elements), which the compiler folds into the instructions' immediate fields.
  static_assert(COLS % NT == 0, "columns divide evenly among the consumer threads");
  constexpr int ROWS = NT / VNX;                        // rows covered by one m-slice
  const int xi = tid % VNX, j0 = tid / VNX;
  const bool at_lo = (xi == 0), at_hi = (xi == VNX - 1);
  const int so0 = (j0 + NG) * VNX + xi;                 // offset of column 0 inside a plane
  const int sr0 = j0 + NG;                              // its fab-relative row
  // fused ghost fill: which faces of the box this thread's cells lie on
  const int fx = at_lo ? -1 : (at_hi ? 1 : 0);
  const long long fab_xg = (long long)2 * NG * STRIP * n2;
  T* xdst = nullptr;      // x-face destination strip of the current fab (threads on an x face only)
  // edge strips of the output level ([box][valid plane][4][VNY] = copies of columns 0, 1, VNX-2, VNX-1)
  constexpr bool XE = (VNX >= 32 && OP == 0 && MODE == 0 && sizeof(T) == 8);
  const int xe_s = xi == 0 ? 0 : (xi == 1 ? 1 : (xi == VNX - 2 ? 2 : (xi == VNX - 1 ? 3 : -1)));    // strip this column is a copy of
  T* const xe_col = (XE && xe_s >= 0 && push.xe_out) ? reinterpret_cast<T*>(push.xe_out) + xe_s * VNY + j0 : nullptr;
  int pbox = -1;               // box whose destination is cached
  int bcur = 0;                // box of plane q (tracked incrementally)

  const int p_lo = s_begin * PS, p_hi = s_end * PS;   // planes whose output this CTA owns
  int q = l_begin * PS;                                // plane being read as the "z+1" plane (virtual index)
  long long qr = (long long)(l_begin + (l_begin >= seam ? jump : 0)) * PS;   // its real index in memory
  int kz = (int)(qr % n2);                             // its index inside its fab
  bcur = push.box0 + (int)(qr / n2);
  int bprev = bcur;                                    // box of plane q-1
  int slot = 0, prev_slot = 0;
  uint32_t phase = 0;
  T cc[CPT], zm[CPT];                             // column values at planes q-1 and q-2
  const T* prev = ring;                           // plane q-1 and its two x-ghost strips
  const T* prev_lo = ring;
  const T* prev_hi = ring;
  bool prev_out = false;
  int kz_prev = 0;                                     // fab plane index of plane q-1

  for (int t = l_begin; t < l_end; ++t) {
    if (t == seam && t != l_begin) { qr += (long long)jump * PS; bcur = push.box0 + (int)(qr / n2); }
    mbar_wait(full + slot, (phase >> slot) & 1u);
    phase ^= 1u << slot;
    const T* __restrict__ sbase = ring + (size_t)slot * STAGE;
#pragma unroll
    for (int p = 0; p < PS; ++p) {
      const T* __restrict__ cur = sbase + p * PLANE + so0;      // column 0 of plane q
      const bool emit = prev_out && (q - 1 >= p_lo) && (q - 1 < p_hi);
      if (!emit) {
#pragma unroll
        for (int m = 0; m < CPT; ++m) { zm[m] = cc[m]; cc[m] = cur[m * NT]; }
      } else {
        T* __restrict__ po = out_main + (size_t)(qr - 1) * PLANE + so0;
        const T* __restrict__ pc = prev + so0;
        const T* __restrict__ pa = sbase + AUX_OFF + p * AUXP + (so0 - NG * VNX);   // OP 1: unm1 of plane q-1
        // x neighbours: from the plane, or from the strip for the first / last column
        const T* __restrict__ pxm = at_lo ? prev_lo + sr0 : pc - 1;
        const T* __restrict__ pxp = at_hi ? prev_hi + sr0 : pc + 1;
        const int xmstep = at_lo ? ROWS : NT, xpstep = at_hi ? ROWS : NT;   // strips advance by rows, planes by NT
        T* __restrict__ xrow = nullptr;
        if (PUSH && !GATHER) {
          if (bprev != pbox) {           // new fab: look up the x neighbour this thread pushes into
            pbox = bprev;
            const int Dx = (fx && (push.mask & 1)) ? __ldg(push.nbr + (size_t)pbox * 27 + 13 + fx) : -1;
            // a +x push lands in the neighbour's LOW strip, a -x push in its HIGH strip
            xdst = Dx >= 0 ? lvl_out + push.xg_base + Dx * fab_xg + (fx > 0 ? 0 : (long long)n2 * N1) + sr0 : nullptr;
          }
          if (xdst) xrow = xdst + (long long)kz_prev * N1;
        }
        T* __restrict__ erow = nullptr;
        if (XE) { if (xe_col) erow = xe_col + ((long long)bprev * (n2 - 2 * NG) + (kz_prev - NG)) * (4 * VNY); }
#pragma unroll
        for (int m = 0; m < CPT; ++m) {
          const T zp = cur[m * NT];
          const T c = cc[m];
          T r;
          if (MODE == 1 || MODE == 3) r = c;
          else if (MODE == 2) r = zp;
          else {
            const T twoc = Ar<T>::mul((T)2, c);
            const T tx = sd2(pxp[m * xpstep], twoc, pxm[m * xmstep]);
            const T ty = sd2(pc[m * NT + VNX], twoc, pc[m * NT - VNX]);
            const T tz = sd2(zp, twoc, zm[m]);
            if (OP == 0) {
              const T S = Ar<T>::add(Ar<T>::add(tx, ty), tz);
              r = (fma & 1) ? Ar<T>::fma(f, S, c) : Ar<T>::add(c, Ar<T>::mul(f, S));
            } else {
              r = Ar<T>::sub(twoc, pa[m * NT]);
              if (fma & 2) {               // SGC_WAVE_DIRSUM: (2u - um1) + f*(Tz + (Ty + Tx))
                const T sum = Ar<T>::add(tz, Ar<T>::add(ty, tx));
                r = (fma & 1) ? Ar<T>::fma(f, sum, r) : Ar<T>::add(r, Ar<T>::mul(f, sum));
              } else if (fma & 1) { r = Ar<T>::fma(f, tx, r); r = Ar<T>::fma(f, ty, r); r = Ar<T>::fma(f, tz, r); }
              else { r = Ar<T>::add(r, Ar<T>::mul(f, tx)); r = Ar<T>::add(r, Ar<T>::mul(f, ty)); r = Ar<T>::add(r, Ar<T>::mul(f, tz)); }
            }
          }
          if (MODE != 2 || zp == (T)1.2345e30) po[m * NT] = r;
          if (PUSH && !GATHER) { if (xrow) xrow[m * ROWS] = r; }
          if (XE) { if (erow) erow[m * ROWS] = r; }
          zm[m] = c;
          cc[m] = zp;
        }
      }
      prev = sbase + p * PLANE;
      prev_lo = sbase + PS * PLANE + p * STRIP;
      prev_hi = sbase + PS * PLANE + PS * STRIP + p * STRIP;
      prev_out = (kz >= NG) && (kz < n2 - NG);
      kz_prev = kz;
      bprev = bcur;
      ++q; ++qr;
      if (++kz == n2) { kz = 0; ++bcur; }
    }
    if (t > l_begin) {                           // stage t-1 is no longer read by this warp
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + prev_slot);
    }
    prev_slot = slot;
    if (++slot == NS) slot = 0;
  }
}

template <int VNX, int VNY, int NG, int PS, int NS, int NCW, int MODE = 0, bool PUSH = false, int OP = 0, bool GATHER = false,
          typename T = double>
static int launch_cfg(const T* in_main, const T* in_xg, T* out_main, int nstages, int n2, double f,
                      int fma, int seam, int jump, cudaStream_t s, PushArgs push = PushArgs(),
                      const T* aux_main = nullptr) {
  constexpr int E = (int)sizeof(T);
  static_assert((PS * VNX * (VNY + 2 * NG) * E) % 16 == 0 && (PS * (VNY + 2 * NG) * E) % 16 == 0, "bulk copies move multiples of 16 bytes");
  constexpr size_t smem = stream_smem_bytes(VNX, VNY, NG, PS, NS, OP, E);
  static_assert(smem <= 232448, "ring exceeds the 227 KB of shared memory a CTA can have");
  static bool configured = false;
  auto kern = k_heat_stream<VNX, VNY, NG, PS, NS, NCW, MODE, PUSH, OP, GATHER, T>;
  if (!configured) {
    SGC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = rt().sm_count * stream_ctas_per_sm(VNX, VNY, NG, PS, NS, OP, E);
  if (rt().sweep_grid_cap > 0 && rt().sweep_grid_cap < rt().sm_count) grid = rt().sweep_grid_cap * stream_ctas_per_sm(VNX, VNY, NG, PS, NS, OP, E);
  if (grid > nstages) grid = nstages;
  kern<<<grid, (NCW + 1 + (GATHER ? kGatherWarps : 0)) * 32, smem, s>>>(in_main, in_xg, out_main, nstages, n2, f, fma, seam, jump, push, aux_main);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// Sweeps boxes [a0,a1) and [b0,b1) (a1 <= b0; the second range may be empty) in ONE launch.
// Returns 1 when the geometry is not covered (caller falls back to the table-driven kernel).
int launch_heat_stream2(sgc_level_s* unew, sgc_level_s* u, double f, int fma, int a0, int a1, int b0, int b1,
                        cudaStream_t s, const PushArgs* fuse) {
  if (!u->uniform || !unew->uniform || u->ncomp != unew->ncomp || u->ng != unew->ng || u->ng < 1 || u->ng > 2) return 1;
  if (u->esz != unew->esz || (u->esz != 8 && u->esz != 4)) return 1;
  if (u->ng == 2 && (fuse || u->esz != 8)) return 1;           // two ghost layers: the plain double sweep
  // Several components: component c of fab b is the virtual fab b*ncomp + c of the plane sequence (the main
  // blocks [comp][k][j][i] and xg blocks [comp][side][k][j] of a fab follow each other).  The fused modes and
  // float levels take the plain sweep only.
  const int nc = u->ncomp;
  if (fuse && (nc != 1 || u->esz != 8)) return 1;
  const FabGeom& g = u->geom[0];
  if (memcmp(g.n, unew->geom[0].n, sizeof g.n) != 0) return 1;
  if (a1 == a0) { a0 = b0; a1 = b1; b0 = b1 = a1; }
  if (b1 == b0) b0 = b1 = a1;
  const int vnx = g.vn[0], vny = g.vn[1], n2 = g.n[2];
  const long long planesA = (long long)(a1 - a0) * nc * n2, planesB = (long long)(b1 - b0) * nc * n2;
  const long long planes = planesA + planesB, gap = (long long)(b0 - a1) * nc * n2;
  if (planes <= 0) return 0;
  if (planes + gap > 0x7fffffffLL) return 1;
  if (getenv("SGC_NO_STREAM")) return 1;
  if (u->esz == 4) {
    const float* in_main = (const float*)u->d_arena + (long long)a0 * nc * g.cs;
    const float* in_xg = (const float*)u->d_arena + u->xg_base + (long long)a0 * nc * g.xcs;
    float* out = (float*)unew->d_arena + (long long)a0 * nc * g.cs;
    unew->xe_valid = 0;
    // stages of an even number of planes keep every bulk copy a multiple of 16 bytes (a strip of 66 floats is 264)
#define SGC_STREAM_CASE_F(VX, VY, PS_, NS_, NCW_)                                                                        \
  if (vnx == VX && vny == VY && n2 % PS_ == 0)                                                                           \
    return launch_cfg<VX, VY, 1, PS_, NS_, NCW_, 0, false, 0, false, float>(in_main, in_xg, out, (int)(planes / PS_), n2, f, fma, \
                                                                          (int)(planesA / PS_), (int)(gap / PS_), s)
    SGC_STREAM_CASE_F(64, 64, 2, 6, 16);
    SGC_STREAM_CASE_F(32, 32, 2, 10, 8);
    SGC_STREAM_CASE_F(16, 16, 6, 6, 8);
#undef SGC_STREAM_CASE_F
    return 1;
  }
  const double* in_main = (const double*)u->d_arena + (long long)a0 * nc * g.cs;
  const double* in_xg = (const double*)u->d_arena + u->xg_base + (long long)a0 * nc * g.xcs;
  double* out = (double*)unew->d_arena + (long long)a0 * nc * g.cs;
  if (u->ng == 2) {
    // levels with two ghost layers (Copier plans exist for any width, Copier.H:521-528): the 7-point sweep reads the inner
    // layer only; same pipeline, planes of (vny + 4) rows, the x-ghost strip = the innermost column of each side
    unew->xe_valid = 0;
#define SGC_STREAM_CASE_G2(VX, VY, PS_, NS_, NCW_)                                                                \
  if (vnx == VX && vny == VY && n2 % PS_ == 0)                                                                    \
    return launch_cfg<VX, VY, 2, PS_, NS_, NCW_>(in_main, in_xg, out, (int)(planes / PS_), n2, f, fma, (int)(planesA / PS_), \
                                                 (int)(gap / PS_), s)
    SGC_STREAM_CASE_G2(64, 64, 1, 6, 16);
    SGC_STREAM_CASE_G2(32, 32, 2, 10, 8);
    SGC_STREAM_CASE_G2(16, 16, 4, 6, 8);
#undef SGC_STREAM_CASE_G2
    return 1;
  }
  PushArgs pa;
  memset(&pa, 0, sizeof pa);
  if (fuse) pa = *fuse;
  pa.out_level = (double*)unew->d_arena;
  pa.in_level = (const double*)u->d_arena;
  pa.xg_base = unew->xg_base;
  pa.box0 = a0;
  // the sweep keeps the edge strips of its output current when it covers the whole level
  const bool whole = (a0 == 0 && a1 == unew->nbox) || (a0 == 0 && b1 == unew->nbox && a1 == b0);
  static const bool no_xe = getenv("SGC_NO_XE") != nullptr;
  pa.xe_out = (whole && unew->xe_elems && !unew->xe_extern && !no_xe) ? (double*)unew->d_arena + unew->xe_base : nullptr;
  unew->xe_valid = 0;
  if (const char* pm = getenv("SGC_PUSH_MASK")) pa.mask = atoi(pm);      // diagnostics
#define SGC_STREAM_CASE(VX, VY, PS_, NS_, NCW_)                                                      \
  if (vnx == VX && vny == VY && n2 % PS_ == 0) {                                                     \
    int st_;                                                                                         \
    if (fuse)                                                                                        \
      st_ = launch_cfg<VX, VY, 1, PS_, NS_, NCW_, 0, true, 0, (VX < 32)>(in_main, in_xg, out, (int)(planes / PS_), n2, \
                                                                         f, fma, (int)(planesA / PS_),   \
                                                                         (int)(gap / PS_), s, pa);       \
    else                                                                                             \
      st_ = launch_cfg<VX, VY, 1, PS_, NS_, NCW_>(in_main, in_xg, out, (int)(planes / PS_), n2, f, fma, \
                                                  (int)(planesA / PS_), (int)(gap / PS_), s, pa);    \
    if (st_ == 0 && pa.xe_out && VX >= 32) unew->xe_valid = 2;                                       \
    return st_;                                                                                      \
  }
  if (const char* m = getenv("SGC_STREAM_MODE")) {      // diagnostics only
    if (vnx == 64 && vny == 64) {
      const int ns = (int)(planes), sm = (int)planesA, jp = (int)gap;
      if (m[0] == '1') return launch_cfg<64, 64, 1, 1, 6, SGC_NCW64, 1>(in_main, in_xg, out, ns, n2, f, fma, sm, jp, s);
      if (m[0] == '2') return launch_cfg<64, 64, 1, 1, 6, SGC_NCW64, 2>(in_main, in_xg, out, ns, n2, f, fma, sm, jp, s);
      if (m[0] == '3') return launch_cfg<64, 64, 1, 1, 6, SGC_NCW64, 3>(in_main, in_xg, out, ns, n2, f, fma, sm, jp, s);
    }
  }
  SGC_STREAM_CASE(64, 64, 1, 6, SGC_NCW64);
  SGC_STREAM_CASE(32, 32, 2, 10, 8);
  // 16^2: a 6-slot ring (93 KB) lets two CTAs share an SM: 16 consumer warps instead of 8 hide the latency
  // of the one-column-per-thread loop; measured 222 -> 277 Gcell/s on 32 768 boxes (a 12-slot ring, or a
  // 4-slot one with three CTAs and a 48-register cap, are both slower)
  SGC_STREAM_CASE(16, 16, 6, 6, 8);
  SGC_STREAM_CASE(8, 8, 10, 24, 2);
#undef SGC_STREAM_CASE
  return 1;
}

// Leap-frog wave step over all boxes with the streaming pipeline; 1 when the geometry is not covered.
int launch_wave_stream(sgc_level_s* unp1, sgc_level_s* un, sgc_level_s* unm1, double f, int fma, cudaStream_t s) {
  if (!un->uniform || un->ncomp != 1 || unp1->ncomp != 1 || unm1->ncomp != 1 || un->ng != 1 || unp1->ng != 1 ||
      unm1->ng != 1 || un->esz != 8 || unp1->esz != 8 || unm1->esz != 8 || getenv("SGC_NO_STREAM"))
    return 1;
  const FabGeom& g = un->geom[0];
  if (memcmp(g.n, unp1->geom[0].n, sizeof g.n) != 0 || memcmp(g.n, unm1->geom[0].n, sizeof g.n) != 0 || !unp1->uniform ||
      !unm1->uniform)
    return 1;
  const int vnx = g.vn[0], vny = g.vn[1], n2 = g.n[2];
  const long long planes = (long long)un->nbox * n2;
  if (planes > 0x7fffffffLL) return 1;
  const double* in_main = (const double*)un->d_arena;
  const double* in_xg = in_main + un->xg_base;
  double* out = (double*)unp1->d_arena;
  const double* aux = (const double*)unm1->d_arena;
  PushArgs none;
  memset(&none, 0, sizeof none);
#define SGC_WAVE_CASE(VX, VY, PS_, NS_, NCW_)                                                                   \
  if (vnx == VX && vny == VY && n2 % PS_ == 0)                                                                  \
    return launch_cfg<VX, VY, 1, PS_, NS_, NCW_, 0, false, 1>(in_main, in_xg, out, (int)(planes / PS_), n2, f, fma, \
                                                              (int)(planes / PS_), 0, s, none, aux)
  SGC_WAVE_CASE(64, 64, 1, 3, 16);
  SGC_WAVE_CASE(32, 32, 2, 6, 8);
  SGC_WAVE_CASE(16, 16, 6, 7, 8);
#undef SGC_WAVE_CASE
  return 1;
}

// does a streaming instance exist for this pair of levels?
bool heat_stream_covers(const sgc_level_s* unew, const sgc_level_s* u) {
  if (!u->uniform || !unew->uniform || u->ncomp != 1 || unew->ncomp != 1 || u->ng != 1 || unew->ng != 1) return false;
  if (u->esz != 8 || unew->esz != 8) return false;
  const FabGeom& g = u->geom[0];
  if (memcmp(g.n, unew->geom[0].n, sizeof g.n) != 0 || getenv("SGC_NO_STREAM")) return false;
  const int vx = g.vn[0], vy = g.vn[1], n2 = g.n[2];
  return (vx == 64 && vy == 64) || (vx == 32 && vy == 32 && n2 % 2 == 0) || (vx == 16 && vy == 16 && n2 % 6 == 0) ||
         (vx == 8 && vy == 8 && n2 % 10 == 0);
}

int launch_heat_stream(sgc_level_s* unew, sgc_level_s* u, double f, int fma, int box_begin, int box_end,
                       cudaStream_t s) {
  return launch_heat_stream2(unew, u, f, fma, box_begin, box_end, box_end, box_end, s, nullptr);
}

}  // namespace sgc
