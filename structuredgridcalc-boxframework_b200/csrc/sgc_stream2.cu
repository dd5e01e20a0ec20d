// sgc_stream2.cu — TWO heat / Jacobi steps per pass over the level (temporal blocking of the
// streaming sweep of sgc_stream.cu), used inside sgc_heat_run.
//
// The one-step sweep is HBM-bound at ~86 % of the measured copy bandwidth, so the only way to go
// faster is to move fewer bytes per cell-update.  This kernel reads level t once and writes level
// t+2 once: the intermediate level t+1 never leaves the SM.  It is bit-identical to two
// { exchange ; sweep } steps of the reference (LevelData.H:455-566 + the application's sweep loop)
// on the valid cells, because every t+1 value it needs — including the ones that belong to the
// neighbouring boxes — is recomputed with the same operations in the same order from the same
// inputs.
//
// The CTA marches DOWN a PILLAR of z-stacked boxes (HC = nbz * VZ planes; nbz = 1: a single box) through HC+4
// "virtual planes" l = 0 .. HC+3, i.e. z = HC+1 .. -2 counted from the pillar's bottom (why downwards: see
// heat_zhalf below; why pillars: the halo planes of a box ARE its z neighbour's outermost planes, and a pillar
// reads and sweeps them once).  For every virtual
// plane a TILE of level t is assembled in a ring slot — entirely from VALID cells (ghost cells are
// not read where a neighbour exists, so no exchange is needed between the two steps or between passes):
//     rows y = -2..V+1 of the V valid columns      (own valid rows + 2 rows of each y neighbour)
//     x strips x = -2,-1,V,V+1 for y = 0..V-1      (the x neighbours' EDGE STRIPS, see below)
//     the 4 corner cells (-1|V, -1|V)              (the xy-diagonal neighbours; 8-byte cp.async)
//   planes z < 0 / z >= HC are the same tile of the z neighbour of the pillar's bottom / top box (a local fab
//   or, on a z-slab partition, a fab in the neighbouring GPU's memory over NVLink, gated by its progress counter).
//   Faces WITHOUT a neighbour (physical boundaries; template flag BND) read the box's own, never
//   exchanged ghost cells instead — see the comment at the producer roles.
// All NCW warps are consumers (thread = one x and CPT consecutive rows, all state in registers);
// six of them also bring in one piece of each tile, the last 4V/32 warps also compute the flanks:
//   stage 1: when the tile of plane z arrives, t+1 on plane z+1 for the V x V cells and for the four
//            1-cell-wide flanks (x = -1, V and y = -1, V) — the values the x / y neighbours would have
//            exchanged.  Results go to registers (z neighbours of stage 2) and to one of two t+1 plane
//            buffers in shared memory (x / y neighbours of stage 2).
//   stage 2: t+2 on plane z+2 from the t+1 planes z+3, z+2, z+1; stored as dense rows.
//   One __syncthreads() per plane orders the t+1 buffer hand-over between the warps.
// EDGE STRIPS (xe): a per-level side array [box][kz][4][V] holding copies of the box's own columns
// x = 0, 1, V-2, V-1, written by stage 2 straight from the result registers (and built once per
// sgc_heat_run by k_build_xe).  Reading a neighbour's edge column out of its main block would
// fetch one 32-byte sector per 8-byte value; the strips make the x halo dense.
//
// HBM traffic per plane pass (V = 64): 36.9 KB tile + 32 KB rows + 2 KB strips for TWO cell-updates
// of 4096 cells = 8.7 B per cell-update (x (HC+4)/HC for the halo planes of a pillar; ncu: 8.8 B) against 16.8 B
// for the one-step sweep; the algorithmic figure stays 16 B per cell-update.
//
// Instances: V = 64 (16 warps, 8 rows per thread, one CTA per SM), V = 32 (8 warps, 4 rows per thread, three
// CTAs per SM) and, on request only, V = 16 (4 warps, 2 rows per thread: slower than the one-step sweep).
// The V = 64 kernel uses the whole register file (128 x 512) without spilling; with four warps per scheduler its
// dependency latency (FP64 chains, shared-memory loads, the per-plane barrier) and its HBM traffic (2.35 GB per
// launch at 0.87-0.89 of the measured copy bandwidth) are in balance.  Its speed is sensitive to anything that
// costs a register or lengthens the serial path from the barrier to the first row
// (DESIGN.md §4.1d has the history; profiles/micro/ab_t2_round2b.sh times variants on one GPU).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sgc_device.h"
#include "sgc_pipe.cuh"

namespace sgc {

// The reference's update, association of its loop: ((u[+e] - 2u) + u[-e]) per direction, x -> y -> z, then
// u + f*S.  2u is exact in binary floating point, so fma(-2, u, p) rounds p - 2u ONCE: the same bits as the
// reference's multiply-then-subtract (short of an overflow of 2u), one FP64 instruction less per update.
//
// The z term is formed in two parts, which is why the march goes DOWN: zh = u[+e] - 2u needs the plane ABOVE,
// the value a column has held in a register since it arrived two tiles ago; the plane BELOW (u[-e]) — the
// arriving tile for stage 1, the fresh stage-1 result for stage 2 — is only added at the end.  The held value
// is therefore dead before the new one is needed and the new one is born in its register.  Marching upwards
// both are operands of the same expression, every new value lands in a different register from the one it
// replaces, and the compiler rotates 32 doubles back with register moves on every second plane.
__device__ __forceinline__ double heat_zhalf(double c, double zabove) { return __fma_rn(-2.0, c, zabove); }
template <bool FMA>
__device__ __forceinline__ double heat_finish(double c, double xm, double xp, double ym, double yp, double zh, double zbelow,
                                              double f) {
  const double tx = __dadd_rn(__fma_rn(-2.0, c, xp), xm);
  const double ty = __dadd_rn(__fma_rn(-2.0, c, yp), ym);
  const double tz = __dadd_rn(zh, zbelow);
  const double S = __dadd_rn(__dadd_rn(tx, ty), tz);
  return FMA ? __fma_rn(f, S, c) : __dadd_rn(c, __dmul_rn(f, S));
}

// shared-memory load from a 32-bit shared address (lets the x-neighbour addresses be formed with ONE
// integer multiply-add: base + column * per-lane byte stride)
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

// CTAs per SM: the 64^2 instance fills the SM (221 KB of shared memory, the whole register file); the 32^2
// instance (256 threads, 62 KB) runs three to a SM
constexpr int t2_ctas_per_sm(int v) { return v >= 64 ? 1 : (v >= 32 ? 3 : 8); }

template <int V, int NS, int NCW, bool FMA, bool BND>
__global__ void __launch_bounds__(NCW * 32, t2_ctas_per_sm(V)) k_heat_t2(T2Args a) {
  constexpr int NT = NCW * 32;                  // threads: every warp is a consumer
  constexpr int N1 = V + 2;                     // fab rows (one ghost layer)
  constexpr int FABP = V * N1;                  // elements of one fab plane in the main block
  constexpr int NR = (V + 4) * V;               // tile rows y = -2..V+1
  constexpr int ESTR = V + 4;                   // allocation of one x strip in the tile
  constexpr int SLOT = NR + 4 * ESTR;           // ring slot, elements
  // t+1 plane buffer: the tile layout shifted down by one row (rows y = -1..V, the x = -1 strip at
  // E(1) - V, the x = V strip at E(2) - V), so that "buffer - V" is addressed with TILE offsets
  constexpr int T1SZ = (NR + 2 * ESTR + 2 + V + 15) / 16 * 16;
  constexpr int G = NT / V;                     // thread groups along y
  constexpr int CPT = V / G;                    // consecutive rows per thread
  constexpr uint32_t TILE_BYTES = (uint32_t)(V * V + 4 * V + 4 * V) * 8u;   // rows + 2x2 y rows + 4 strips
  static_assert(NT % V == 0 && V % G == 0, "thread mapping");
  constexpr int FW = 4 * V / 32;                // flank warps: the last FW warps own one flank cell per thread
  static_assert((4 * V) % 32 == 0 && FW <= NCW, "flank cells");
  static_assert(2 * NCW >= 6, "six roles bring in the tile, at most two per warp");
  constexpr int NROLE = 6;                      // warps that bring in a piece of each tile
  // arrivals per tile: warps 0-4 + the four corner copies of warp 5; with physical boundaries (BND) the x
  // strips may come by 8-byte cp.async copies, so every lane of warps 3 and 4 arrives
  constexpr int NARRIVE = BND ? 3 + 64 + 4 : 5 + 4;
  static_assert((NS & (NS - 1)) == 0 && NS >= 4, "ring: two tiles being read, two in flight");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);
  double* t1 = ring + (size_t)NS * SLOT;
  uint64_t* full = reinterpret_cast<uint64_t*>(t1 + 2 * T1SZ);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // PILLARS.  Boxes stacked in z whose faces towards each other are ordinary local neighbours form a pillar
  // (box p + k * npil, k = 0 .. nbz-1 from the bottom; the host checks the structure, launch_heat_t2).  The march
  // runs down a whole pillar: HC = nbz * vz planes + the two halo planes above and below, so the tiles of a box's
  // halo planes — which ARE its z neighbour's outermost planes — are read and swept once per pillar, not once
  // per box (512^3 in 64^3 boxes: 516 tiles per pillar instead of 8 x 68).  nbz = 1 is the box-by-box march.
  // Only the periodic 64^2 instance has the registers for it (kPil; the boundary instance spills 52 bytes with
  // it and runs 9 % slower than box by box); the others are launched with nbz = 1.
  constexpr bool kPil = V >= 64 && !BND;
  const int vz = a.vz, npil = a.npil, nbz = kPil ? a.nbz : 1, HC = nbz * vz, VP = HC + 4;
  const int topoff = (nbz - 1) * npil;                   // top box of pillar p = p + topoff
  const long long fabm = (long long)FABP * (vz + 2);
  // ranges start on even virtual planes (vz is even: heat_t2_covers), which makes the parity of a plane's
  // t+1 buffer the parity of the loop counter
  const int s_begin = 2 * (int)((long long)blockIdx.x * (a.nplanes / 2) / gridDim.x);
  const int s_end = 2 * (int)((long long)(blockIdx.x + 1) * (a.nplanes / 2) / gridDim.x);
  if (s_begin >= s_end) return;
  if (a.err && *reinterpret_cast<const volatile unsigned int*>(a.err)) return;   // an earlier wait of this phase timed out
  const int l_begin = s_begin >= 4 ? s_begin - 4 : 0;     // 4 tiles prime the two stages

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(full + i, NARRIVE);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto E = [](int s) { return NR + s * ESTR + 2; };       // tile offset of strip s at y = 0

  // ---------------- tile roles (six, one piece of the tile each; warp r plays role r, and role r + NCW if NCW < 6) ----------------
  // Tile t goes to ring slot (t - l_begin) % NS.  The address arithmetic of a tile is spread over six
  // roles so that no warp lags at the per-plane barrier: lane 0 of the warp playing role r issues
  //   r = 0: the V own rows (+ announces the tile's bytes)   r = 1 / 2: the two rows of the -y / +y neighbour
  //   r = 3 / 4: the two edge strips of the -x / +x neighbour
  // and lanes 0-3 of role 5 fetch the four corner cells with 8-byte cp.async copies, each arriving on
  // the slot's mbarrier when it lands (cp.async.mbarrier.arrive.noinc).  issue() starts all of that for
  // the tile two iterations ahead; the slot completes on nine arrivals + the bulk copies' bytes.
  // BND: a face of a box without a motion item in the plan (a physical boundary, or a trimmed direction)
  // keeps its ghost cells for ever, each level its own: step t reads the ghosts of a.ghost[0], step t+1
  // those of a.ghost[1].  Where a neighbour is missing the tile takes the level-t ghost cells instead of
  // the neighbour's valid cells, and the slot of the second halo layer (x = -2 / V+1 strip, row y = -2 /
  // V+1, plane z = -2 / vz+1) — not needed there — carries the level-t+1 ghost cells, which the consumers
  // copy where they would have computed a flank / halo-plane value.
  // per role a warp plays (one for NCW >= 6; with fewer warps a warp plays roles `warp` and `warp + NCW`):
  // S = source box of the current tiles (bit 30 remote, 29 peer slot, 28 z face missing), N = this role's
  // neighbour of S (BND: -1 = missing), ct = BND corner role: where the lane's corner cell comes from
  int Sa = -1, Na = 0, cta = 0, Sb = -1, Nb = 0, ctb = 0;
  int Ka = 0, Kb = 0;                 // plane (valid index) of the role's source box the previous tile came from
  unsigned waited = 0;                // bit per peer slot
  auto issue = [&](int slot, int tpil, int tl, const int role, int& S, int& N, int& ct, int& K) {   // tile = virtual plane tl of pillar tpil
    const int h = HC + 1 - tl;                             // plane of the pillar, from its bottom; the march goes down
    const int cls = h < 0 ? 0 : (h >= HC ? 2 : 1);         // below the pillar / inside / above
    // first tile, or the source box changes here (kPil: the previous tile was its plane 0)
    const bool fresh = S < 0 || tl == 0 || (kPil ? K <= 0 : (tl == 2 || tl == vz + 2));
    if (!kPil) K = cls == 0 ? h + vz : (cls == 2 ? h - vz : h);
    if (fresh) {
      if (cls == 1) {
        if (kPil) { const int kb = h / vz; K = h - kb * vz; S = tpil + kb * npil; }
        else S = tpil;
      } else {
        const int tbox = cls == 0 ? tpil : tpil + topoff;  // the pillar's bottom / top box
        if (kPil) K = cls == 0 ? h + vz : h - HC;
        S = __ldg(a.nbr + (size_t)tbox * 27 + (cls == 0 ? 4 : 22));
        if (S < 0) {
          const int rr = a.rz ? __ldg(a.rz + 2 * tbox + (cls == 0 ? 0 : 1)) : -1;
          if (BND && rr < 0) S = tbox | 0x10000000;          // no z neighbour at all: the box's own ghost planes
          else {                                 // z neighbour lives on another GPU (checked on the host)
            const int ps = rr >> 28;
            S = (rr & 0x0fffffff) | (ps ? 0x60000000 : 0x40000000);     // bit 30: remote, bit 29: peer slot
            if (!(waited >> ps & 1u)) {          // the peer must have finished writing the level we read
              if (lane == 0) {
                if (!wait_peer_counter(ps ? a.peer_flag[1] : a.peer_flag[0], a.wait_value, a.wait_budget_ns)) *a.err = 1u;
              }
              __syncwarp();
              __threadfence_system();
              asm volatile("fence.proxy.async;" ::: "memory");
              waited |= 1u << ps;
            }
          }
        }
      }
      // this role's in-plane neighbour of S (on a z-slab partition the peer's local indices equal ours)
      const int* __restrict__ nb = a.nbr + (size_t)(S & 0x0fffffff) * 27;
      if (role == 0) N = S & 0x0fffffff;
      else if (role < 5) N = __ldg(nb + (role == 1 ? 10 : (role == 2 ? 16 : (role == 3 ? 12 : 14))));
      else {
        // corner (dx,dy) by lane bits: the valid cell of the diagonal box, reached over two faces
        const int cx = (lane & 1) ? 14 : 12, cy = (lane & 2) ? 16 : 10;
        const int X = __ldg(nb + cx), Y = __ldg(nb + cy);
        if (!BND) N = __ldg(a.nbr + (size_t)X * 27 + cy);
        else if (X >= 0 && Y >= 0) { N = __ldg(a.nbr + (size_t)X * 27 + cy); ct = 0; }
        else if (X >= 0) { N = X; ct = 1; }                  // y face missing: the x neighbour's ghost row
        else if (Y >= 0) { N = Y; ct = 2; }                  // x face missing: the y neighbour's ghost column
        else { N = 0; ct = 3; }                              // neither flank is computed: not needed
      }
    } else if (kPil) --K;
    const bool zmiss = BND && (S & 0x10000000);
    const int kz = K;                                               // plane of the source box (valid index)
    // fab plane the tile is read from; a missing z face: the lower / upper GHOST plane of the box and of its
    // in-plane neighbours (whose cells on that plane are the z neighbours of the flank cells), taken from
    // the level-t ghosts for the inner halo plane and from the level-t+1 ghosts for the outer one
    const int fp = !zmiss ? kz + 1 : (cls == 0 ? 0 : vz + 1);
    const long long prow = ((long long)fp * N1 + 1) * V;            // first valid row of that fab plane
    double* dst = ring + (size_t)slot * SLOT;
    const bool rem = (S & 0x40000000) != 0, ps1 = (S & 0x20000000) != 0;
    // BND: the levels whose ghost cells are level t (g0) and level t+1 (g1), on the GPU that owns S
    const double* g0 = !BND ? nullptr : (!rem ? a.ghost[0] : a.peer_ghost[ps1 ? 1 : 0][0]);
    const double* g1 = !BND ? nullptr : (!rem ? a.ghost[1] : a.peer_ghost[ps1 ? 1 : 0][1]);
    const double* base = zmiss ? ((tl == 1 || tl == HC + 2) ? g0 : g1)
                               : (!rem ? a.in_level : (ps1 ? a.peer_in[1] : a.peer_in[0]));
    if (role == 5) {
      if (lane < 4) {
        const bool xh = lane & 1, yh = lane & 2;
        const double* src = base + N * fabm + prow + (long long)(yh ? 0 : V - 1) * V + (xh ? 0 : V - 1);
        if (BND && ct == 1) src = g0 + N * fabm + ((long long)fp * N1 + (yh ? N1 - 1 : 0)) * V + (xh ? 0 : V - 1);
        if (BND && ct == 2)
          src = g0 + a.xg_base + (long long)N * (2 * N1 * (vz + 2)) + ((long long)((xh ? 1 : 0) * (vz + 2) + fp)) * N1 + 1 +
                (yh ? 0 : V - 1);
        if (!BND || (ct != 3 && !zmiss)) cp_async8(dst + E(xh ? 2 : 1) + (yh ? V : -1), src);
        cp_async_arrive(full + slot);
      }
    } else if (BND && role >= 3) {
      // x strips: the neighbour's edge strips (two bulk copies), or — no neighbour — the ghost columns of S
      // itself in the two ghost levels, 8 bytes at a time (a ghost strip starts 8 bytes off a 16-byte
      // boundary, which bulk copies cannot do)
      const bool hi = role == 4;
      if (zmiss) {
        // ghost plane: the x neighbour's ghost-plane column next to this box (not part of its edge strips)
        if (N >= 0) {
#pragma unroll
          for (int y = lane; y < V; y += 32)
            cp_async8(dst + E(hi ? 2 : 1) + y, base + N * fabm + prow + (long long)y * V + (hi ? 0 : V - 1));
        }
      } else if (N >= 0) {
        if (lane == 0) {
          const double* e = base + a.xe_in + ((long long)N * vz + kz) * (4 * V) + (hi ? 0 : 2 * V);
          mbar_expect_tx_only(full + slot, 2 * V * 8u);
          bulk_g2s(dst + E(hi ? 2 : 0), e, V * 8u, full + slot);
          bulk_g2s(dst + E(hi ? 3 : 1), e + V, V * 8u, full + slot);
        }
      } else {
        const long long go = a.xg_base + (long long)(S & 0x0fffffff) * (2 * N1 * (vz + 2)) +
                             ((long long)((hi ? 1 : 0) * (vz + 2) + fp)) * N1 + 1;
#pragma unroll
        for (int y = lane; y < V; y += 32) {
          cp_async8(dst + E(hi ? 2 : 1) + y, g0 + go + y);      // level t ghosts at x = -1 / V
          cp_async8(dst + E(hi ? 3 : 0) + y, g1 + go + y);      // level t+1 ghosts, in the unused outer strip
        }
      }
      cp_async_arrive(full + slot);
    } else if (lane == 0) {
      if (role == 0) {
        if (BND) mbar_expect_tx(full + slot, V * V * 8u);
        else mbar_expect_tx(full + slot, TILE_BYTES);             // this warp's arrival + the tile's byte count
        bulk_g2s(dst + 2 * V, base + N * fabm + prow, V * V * 8u, full + slot);                           // rows 0..V-1
      } else if (role == 1 || role == 2) {
        const bool hi = role == 2;
        if (BND) mbar_expect_tx_only(full + slot, 2 * V * 8u);
        if (!BND || N >= 0) {
          if (!hi) bulk_g2s(dst, base + N * fabm + prow + (long long)(V - 2) * V, 2 * V * 8u, full + slot);   // rows -2,-1
          else bulk_g2s(dst + (V + 2) * V, base + N * fabm + prow, 2 * V * 8u, full + slot);                  // rows V,V+1
        } else {
          // no y neighbour: the ghost row of S in the level-t ghosts, and the level-t+1 ghosts in the outer row
          const long long ro = (long long)(S & 0x0fffffff) * fabm + ((long long)fp * N1 + (hi ? N1 - 1 : 0)) * V;
          bulk_g2s(dst + (hi ? (V + 2) * V : V), g0 + ro, V * 8u, full + slot);
          bulk_g2s(dst + (hi ? (V + 3) * V : 0), g1 + ro, V * 8u, full + slot);
        }
      } else {
        const double* e = base + a.xe_in + ((long long)N * vz + kz) * (4 * V);
        if (role == 3) {
          bulk_g2s(dst + E(0), e + 2 * V, V * 8u, full + slot);     // x = -2  <- x neighbour's column V-2
          bulk_g2s(dst + E(1), e + 3 * V, V * 8u, full + slot);     // x = -1  <- column V-1
        } else {
          bulk_g2s(dst + E(2), e, V * 8u, full + slot);             // x = V   <- column 0
          bulk_g2s(dst + E(3), e + V, V * 8u, full + slot);         // x = V+1 <- column 1
        }
      }
      if (role != 0) mbar_arrive(full + slot);
    }
  };
  // a warp's role(s) for one tile
  auto issue_roles = [&](int slot, int tpil, int tl) {
    if (warp < NROLE) issue(slot, tpil, tl, warp, Sa, Na, cta, Ka);
    if (NCW < NROLE && warp + NCW < NROLE) issue(slot, tpil, tl, warp + NCW, Sb, Nb, ctb, Kb);
  };
  {
    const int b0 = l_begin / VP, l0 = l_begin - b0 * VP;
    issue_roles(0, b0, l0);
    if (l_begin + 1 < s_end) issue_roles(1, l0 + 1 == VP ? b0 + 1 : b0, l0 + 1 == VP ? 0 : l0 + 1);
  }

  // ---------------- flank role (the last FW warps): one cell of the x / y flanks per thread ----------------
  const bool flank = warp >= NCW - FW;
  double hA = 0.0, hB = 0.0;          // the flank cell at the two latest planes of level t (roles alternate)
  // tile offsets of the flank cell and its four in-plane neighbours.  Flank cell number
  // (warp - first flank warp) * 32 + lane: side 0..3 = x = -1, x = V, y = -1, y = V; position p along it
  auto flank_offsets = [&](int& hc, int& hxm, int& hxp, int& hym, int& hyp, int& side, int& p) {
    const int fidx = (warp - (NCW - FW)) * 32 + lane;
    side = V >= 32 ? ((warp - (NCW - FW)) * 32) / V : fidx / V;      // warp-uniform for V >= 32
    p = V >= 32 ? ((warp - (NCW - FW)) * 32) % V + lane : fidx % V;
    if (side == 0)      { hc = E(1) + p; hxm = E(0) + p; hxp = (p + 2) * V; hym = hc - 1; hyp = hc + 1; }             // x = -1
    else if (side == 1) { hc = E(2) + p; hxm = (p + 2) * V + V - 1; hxp = E(3) + p; hym = hc - 1; hyp = hc + 1; }     // x = V
    else if (side == 2) { hc = V + p; hym = p; hyp = 2 * V + p;                                                       // y = -1
                      hxm = p > 0 ? hc - 1 : E(1) - 1; hxp = p < V - 1 ? hc + 1 : E(2) - 1; }
    else            { hc = (V + 2) * V + p; hym = (V + 1) * V + p; hyp = (V + 3) * V + p;                         // y = V
                      hxm = p > 0 ? hc - 1 : E(1) + V; hxp = p < V - 1 ? hc + 1 : E(2) + V; }
  };
  // Per-thread constants of the plane loop.  What lies between the per-plane barrier and a warp's first row is
  // serial latency every warp pays 68 times per box, so the 64^2 instance keeps these in registers across the
  // loop (kKeep); the 32^2 / 16^2 instances (80 / 64 registers per thread) would spill and re-derive them.
  constexpr bool kKeep = V >= 64;
  int hc0 = 0, hxm0 = 0, hxp0 = 0, hym0 = 0, hyp0 = 0, side0 = 0, p0 = 0;
  if (kKeep && flank) flank_offsets(hc0, hxm0, hxp0, hym0, hyp0, side0, p0);

  // ---------------- regular role (every thread): x = xi, rows y0 .. y0+CPT-1 ----------------
  const int xi = tid % V, y0 = (tid / V) * CPT;
  const int oC0 = (y0 + 2) * V + xi;                              // tile offset of (xi, y0)
  // x neighbours: the plane (offset -1 / +1), or the x = -1 / x = V strip for the first / last column of
  // a row.  Per lane that is a base offset and a byte stride per row (V*8 in the plane, 8 in a strip);
  // the strides are made opaque to the compiler so that an access costs one multiply-add + one load
  // (it would otherwise select between per-column constants: three integer instructions per access).
  int oxm0 = xi == 0 ? E(1) + y0 : oC0 - 1, sxm = xi == 0 ? 8 : V * 8;
  int oxp0 = xi == V - 1 ? E(2) + y0 : oC0 + 1, sxp = xi == V - 1 ? 8 : V * 8;
  asm volatile("" : "+r"(sxm), "+r"(sxp));
  // edge-strip index of this thread's column (-1: none), opaque and therefore kept in a register: left to itself
  // the compiler re-derives it from the thread index on every plane — through a JUMP TABLE (a constant-bank load
  // and an indirect branch in all 16 warps; hoisting it took 5 % off the launch)
  int xs = xi == 0 ? 0 : (xi == 1 ? 1 : (xi == V - 2 ? 2 : (xi == V - 1 ? 3 : -1)));
  asm volatile("" : "+r"(xs));
  // Column state, all in registers: level t at the two latest planes (tA / tB) and level t+1 at the two
  // latest planes (uA / uB).  The roles "newest / previous" alternate from plane to plane — the loop
  // below is unrolled by two so that no value is ever moved between registers.
  double tA[CPT], tB[CPT], uA[CPT], uB[CPT];
#pragma unroll
  for (int m = 0; m < CPT; ++m) tA[m] = tB[m] = uA[m] = uB[m] = 0.0;
  const double f = a.f;
  int pil = l_begin / VP, l = l_begin - pil * VP;
  // the plane stage 2 writes, two virtual planes behind the arriving tile: plane okz of box obx (for l < 4, where
  // nothing is written yet, okz >= vz counts down towards the top box's last plane), stepped at the end of a plane
  int obx_run = 0, okz_run = 0;       // kPil only; a box-by-box march derives them from (pil, l)
  if (kPil) {
    const int ho = HC + 3 - l, kb = ho / vz < nbz - 1 ? ho / vz : nbz - 1;
    obx_run = pil + kb * npil;
    okz_run = ho - kb * vz;
  }
  // kKeep: running bases of that box in the output level (its main block / its edge strips, at this thread's
  // column and first row), advanced when the box changes instead of multiplied out of the box index
  double* obox = a.out_level + (kKeep ? (kPil ? obx_run : pil) * fabm + (oC0 + (N1 - 1) * V) : 0);
  double* ebox = a.out_level + a.xe_out + (kKeep ? (long long)(kPil ? obx_run : pil) * vz * (4 * V) + (oC0 / V - 2) : 0);
  const int n_iter = s_end - l_begin, lead = s_begin - l_begin;

  // One plane: tile i arrives as the plane BELOW stage 1's.  X = level t at the previous plane (centre of
  // stage 1), Y = the plane before, i.e. above (replaced by the arriving plane); U = level t+1 at the previous
  // plane (centre of stage 2), W = the plane before (replaced by the stage-1 result).
#ifdef SGC_T2_DIAG
  // diagnostic build (profiles/micro/ab_t2_round2b.sh): cycles each warp of one CTA waits for the tile and at the
  // per-plane barrier — the warps that wait least at the barrier are the ones the others wait for
  long long diag_tile = 0, diag_bar = 0;
  const long long diag_t0 = clock64();
#endif
  auto plane = [&](int i, double (&X)[CPT], double (&Y)[CPT], double (&U)[CPT], double (&W)[CPT], double& hX, double& hY) {
    const int slot = i & (NS - 1);
#ifdef SGC_T2_DIAG
    const long long d0 = clock64();
    mbar_wait(full + slot, (i / NS) & 1u);
    const long long d1 = clock64();
    __syncthreads();
    const long long d2 = clock64();
    diag_tile += d1 - d0; diag_bar += d2 - d1;
#else
    mbar_wait(full + slot, (i / NS) & 1u);
    __syncthreads();                 // t+1 buffer hand-over; every warp has left tile i-2 and t+1 plane i-3
#endif
    const bool more = i + 2 < n_iter;
    if (more) issue_roles((i + 2) & (NS - 1), l + 2 >= VP ? pil + 1 : pil, l + 2 >= VP ? l + 2 - VP : l + 2);
    const double* __restrict__ C = ring + (size_t)slot * SLOT;
    const double* __restrict__ P = ring + (size_t)((i + NS - 1) & (NS - 1)) * SLOT;
    const bool do1 = (l >= 2) && (i >= 2);
    const bool do2 = (l >= 4) && (i >= lead);
    // t+1 plane buffers, addressed with tile offsets
    const int par = i & 1;                                        // parity of the virtual plane v = l_begin + i
    double* __restrict__ Tw = t1 + (par ^ 1) * T1SZ - V;          // virtual plane v-1 (written now)
    const double* __restrict__ Tr = t1 + par * T1SZ - V;          // virtual plane v-2 (written last iteration)
    if (flank) {
      int hc = hc0, hxm = hxm0, hxp = hxp0, hym = hym0, hyp = hyp0, side = side0, p = p0;
      if (!kKeep) flank_offsets(hc, hxm, hxp, hym, hyp, side, p);
      const double zh = heat_zhalf(hX, hY);                       // the held plane is consumed here ...
      const double zp = C[hc];                                    // ... and the arriving one may take its register
      if (do1) {
        double r = heat_finish<FMA>(hX, P[hxm], P[hxp], P[hym], P[hyp], zh, zp, f);
        if (BND) {
          // no neighbour on this flank's side: the "flank" is this box's ghost cell; its level-t+1 value was
          // brought in the slot of the outer halo layer
          const int code = side == 0 ? 12 : (side == 1 ? 14 : (side == 2 ? 10 : 16));
          const int alt = side == 0 ? E(0) + p : (side == 1 ? E(3) + p : (side == 2 ? p : (V + 3) * V + p));
          if (__ldg(a.nbr + (size_t)pil * 27 + code) < 0) r = P[alt];        // the same for every box of a pillar (host-checked)
        }
        Tw[hc] = r;
      }
      hY = zp;
    }
    if (!do1) {
#pragma unroll
      for (int m = 0; m < CPT; ++m) Y[m] = C[oC0 + m * V];
    } else {
      // output plane inside the fab's valid region.  While the stages are being primed (do2 false) the
      // stage-2 values are meaningless: they go to a scratch plane and a spare strip plane behind the edge
      // strips, which keeps the stores free of branches.
      const int obx = kPil ? obx_run : pil, okz = kPil ? okz_run : vz + 3 - l;
      // (xi, y0) of fab plane okz: fab-relative offset = tile offset + (N1 - 1) * V
      double* __restrict__ po =
          do2 ? (kKeep ? obox + (long long)okz * FABP : obox + obx * fabm + (long long)okz * FABP + (oC0 + (N1 - 1) * V))
              : a.out_level + a.dump + (oC0 - 2 * V);
      // BND: a missing z face: level t+1 on the ghost plane above / below the pillar is not computed but is the
      // ghost plane of the level-t+1 ghosts, which arrived as the tile of the outer halo plane: it sits in
      // Y (plane HC+1, read two planes ago: l == 2) resp. in the arriving tile (plane -2: l == HC+3)
      bool ovr_held = false, ovr_arriving = false;
      if (BND && (l == 2 || l == HC + 3)) {
        const bool top = l == 2;                                  // the march starts above the pillar
        const int ebx = top ? pil + topoff : pil;                 // the pillar's top / bottom box
        const bool miss = __ldg(a.nbr + (size_t)ebx * 27 + (top ? 22 : 4)) < 0 && !(a.rz && __ldg(a.rz + 2 * ebx + (top ? 1 : 0)) >= 0);
        ovr_held = miss && top; ovr_arriving = miss && !top;
      }
      // edge strips of the output plane: strip xs of this thread's column, if it is one of the four
      double* __restrict__ pe =
          (do2 ? (kKeep ? ebox + (long long)okz * (4 * V) : ebox + ((long long)obx * vz + okz) * (4 * V) + (oC0 / V - 2))
               : a.out_level + a.xe_out + a.xe_spare_plane * (4 * V) + (oC0 / V - 2)) + (xs < 0 ? 0 : xs) * V;
      const double yfirst = P[oC0 - V], ylast = P[oC0 + CPT * V];
      const double yfirst1 = Tr[oC0 - V], ylast1 = Tr[oC0 + CPT * V];
      const uint32_t pxm = smem_u32(P + oxm0), pxp = smem_u32(P + oxp0);      // x neighbours of row y0, level t
      const uint32_t qxm = smem_u32(Tr + oxm0), qxp = smem_u32(Tr + oxp0);    // and level t+1
      // Row m's held values are consumed FIRST (zh = Y - 2X, wh = W - 2U) and only then replaced — Y[m] by the
      // arriving plane's value, W[m] by the stage-1 result — so each new value is born in the register of the
      // value it replaces.  Row m+1's zh, arriving value and x neighbours are formed / requested before the
      // arithmetic of row m (software prefetch).
      double zhN = heat_zhalf(X[0], Y[0]);
      const double y0old = BND ? Y[0] : 0.0;
      Y[0] = C[oC0];
      double xmN = lds64(pxm), xpN = lds64(pxp), yoldN = y0old;
      double oprev = 0.0;
#pragma unroll
      for (int m = 0; m < CPT; ++m) {
        const double zh = zhN, xm = xmN, xp = xpN, yold = yoldN;
        if (m < CPT - 1) {
          zhN = heat_zhalf(X[m + 1], Y[m + 1]);
          if (BND) yoldN = Y[m + 1];
          Y[m + 1] = C[oC0 + (m + 1) * V];
          xmN = lds64(pxm + (m + 1) * sxm);
          xpN = lds64(pxp + (m + 1) * sxp);
        }
        const double wh = heat_zhalf(U[m], W[m]);
        W[m] = heat_finish<FMA>(X[m], xm, xp, m > 0 ? X[m - 1] : yfirst, m < CPT - 1 ? X[m + 1] : ylast, zh, Y[m], f);
        if (BND) W[m] = ovr_held ? yold : (ovr_arriving ? Y[m] : W[m]);
        Tw[oC0 + m * V] = W[m];
        // stage 2
        const double xm1 = lds64(qxm + m * sxm), xp1 = lds64(qxp + m * sxp);
        const double o = heat_finish<FMA>(U[m], xm1, xp1, m > 0 ? U[m - 1] : yfirst1, m < CPT - 1 ? U[m + 1] : ylast1,
                                          wh, W[m], f);
        po[m * V] = o;
        // edge strips: two rows per 16-byte store (only the four edge columns of 64 store at all)
        if (m & 1) { if (xs >= 0) *reinterpret_cast<double2*>(pe + m - 1) = make_double2(oprev, o); } else oprev = o;
      }
    }
    if (++l == VP) {                       // next pillar: from the bottom box of this one to the top box of the next
      l = 0; ++pil;
      if (kPil) { okz_run = vz + 3; obx_run += topoff + 1; }
      if (kKeep) { obox += (topoff + 1) * fabm; ebox += (long long)(topoff + 1) * vz * (4 * V); }
    } else if (kPil && --okz_run < 0) {    // the box below in the same pillar
      okz_run = vz - 1; obx_run -= npil;
      if (kKeep) { obox -= npil * fabm; ebox -= (long long)npil * vz * (4 * V); }
    }
  };
  for (int i = 0; i < n_iter; i += 2) {
    plane(i, tA, tB, uA, uB, hA, hB);
    if (i + 1 < n_iter) plane(i + 1, tB, tA, uB, uA, hB, hA);
  }
#ifdef SGC_T2_DIAG
  if (blockIdx.x == 5 && lane == 0 && V == 64 && !BND)
    printf("T2DIAG warp %d planes %d total %lld tile_wait %lld barrier_wait %lld\n", warp, n_iter, clock64() - diag_t0, diag_tile, diag_bar);
#endif
}

// edge strips of a level from its main blocks: xe[box][kz][s][y], s = columns 0, 1, V-2, V-1
__global__ void k_build_xe(const double* __restrict__ level, double* __restrict__ xe, int nbox, int vx, int vy, int vz) {
  const long long n = (long long)nbox * vz * 4 * vy;
  const int n1 = vy + 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i % vy);
    const int s = (int)((i / vy) % 4);
    const int kz = (int)((i / (4 * vy)) % vz);
    const long long b = i / ((long long)4 * vy * vz);
    const int x = s == 0 ? 0 : (s == 1 ? 1 : (s == 2 ? vx - 2 : vx - 1));
    xe[i] = level[b * ((long long)vx * n1 * (vz + 2)) + ((long long)(kz + 1) * n1 + (y + 1)) * vx + x];
  }
}

template <int V, int NS, int NCW>
static int launch_t2_cfg(const T2Args& a, int fma, cudaStream_t s) {
  constexpr size_t smem = ((size_t)NS * ((V + 4) * V + 4 * (V + 4)) + 2 * (((V + 4) * V + 2 * (V + 4) + 2 + V + 15) / 16 * 16)) * 8 + NS * 8;
  static_assert(smem <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
  static bool configured = false;
  if (!configured) {
    SGC_CUDA(cudaFuncSetAttribute(k_heat_t2<V, NS, NCW, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGC_CUDA(cudaFuncSetAttribute(k_heat_t2<V, NS, NCW, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGC_CUDA(cudaFuncSetAttribute(k_heat_t2<V, NS, NCW, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SGC_CUDA(cudaFuncSetAttribute(k_heat_t2<V, NS, NCW, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  int grid = rt().sm_count * t2_ctas_per_sm(V);
  if (grid > a.nplanes / 8) grid = a.nplanes / 8 > 0 ? a.nplanes / 8 : 1;
  if (a.ghost[0]) {          // levels with faces that have no neighbour (physical boundaries)
    if (fma) k_heat_t2<V, NS, NCW, true, true><<<grid, NCW * 32, smem, s>>>(a);
    else k_heat_t2<V, NS, NCW, false, true><<<grid, NCW * 32, smem, s>>>(a);
  } else {
    if (fma) k_heat_t2<V, NS, NCW, true, false><<<grid, NCW * 32, smem, s>>>(a);
    else k_heat_t2<V, NS, NCW, false, false><<<grid, NCW * 32, smem, s>>>(a);
  }
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

// does the two-step kernel have an instance for this pair of levels?
bool heat_t2_covers(const sgc_level_s* unew, const sgc_level_s* u) {
  if (!u->uniform || !unew->uniform || u->ncomp != 1 || unew->ncomp != 1 || u->ng != 1 || unew->ng != 1) return false;
  if (u->esz != 8 || unew->esz != 8 || !u->xe_elems || !unew->xe_elems || getenv("SGC_NO_T2")) return false;
  const FabGeom& g = u->geom[0];
  if (memcmp(g.n, unew->geom[0].n, sizeof g.n) != 0 || u->nbox != unew->nbox) return false;
  // the 16^2 instance (4 warps, two tile roles per warp, 8 CTAs per SM) is bit-exact but SLOWER than the one-step
  // sweep (163 against 271 Gcell/s on 32 768 boxes: with 2 rows per thread the per-plane barrier, tile issue and
  // flank work dominate), so it only runs on request
  const bool shape = (g.vn[0] == 64 && g.vn[1] == 64) || (g.vn[0] == 32 && g.vn[1] == 32) ||
                     (g.vn[0] == 16 && g.vn[1] == 16 && getenv("SGC_T2_16"));
  return shape && g.vn[2] >= 4 && g.vn[2] % 2 == 0;
}

int launch_build_xe(sgc_level_s* l, cudaStream_t s) {
  const FabGeom& g = l->geom[0];
  const long long n = (long long)l->nbox * g.vn[2] * 4 * g.vn[1];
  const int grid = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  k_build_xe<<<grid, 256, 0, s>>>((const double*)l->d_arena, (double*)l->d_arena + l->xe_base, l->nbox, g.vn[0], g.vn[1],
                                  g.vn[2]);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  l->xe_valid = l->xe_extern ? 0 : 2;
  return 0;
}

// unew(t+2) <- u(t): two steps over all boxes in one launch.  `remote` carries the peer-memory fields.
int launch_heat_t2(sgc_level_s* unew, sgc_level_s* u, double f, int fma, const int* d_nbr, const T2Args* remote,
                   cudaStream_t s) {
  const FabGeom& g = u->geom[0];
  T2Args a;
  memset(&a, 0, sizeof a);
  if (remote) a = *remote;
  a.in_level = (const double*)u->d_arena;
  a.out_level = (double*)unew->d_arena;
  a.xe_in = u->xe_base;
  a.xe_out = unew->xe_base;
  a.nbr = d_nbr;
  a.vz = g.vn[2];
  // pillars (plan->t2_npil / t2_nbz through `remote`); none given, or SGC_T2_NO_PILLARS: every box its own pillar
  if (a.npil <= 0 || a.nbz <= 0 || a.npil * a.nbz != u->nbox || g.vn[0] != 64 || a.ghost[0] || getenv("SGC_T2_NO_PILLARS")) {
    a.npil = u->nbox;
    a.nbz = 1;
  }
  const long long np = (long long)a.npil * ((long long)a.nbz * g.vn[2] + 4);
  if (np > 0x7fffffffLL) return 1;
  a.nplanes = (int)np;
  a.xe_spare_plane = (long long)u->nbox * g.vn[2];
  a.dump = unew->xe_base + (a.xe_spare_plane + 1) * 4 * g.vn[1];     // scratch plane behind the spare strip plane
  a.xg_base = u->xg_base;
  a.f = f;
  // 64^2 boxes: 16 warps, 8 rows per thread; 32^2 boxes: 8 warps, 4 rows per thread, three CTAs per SM
  // (V = 64 with 32 warps of 4 rows — a 64-register cap — spills and is 1.5x slower)
  unew->xe_valid = unew->xe_extern ? 0 : 2;        // stage 2 stores all four edge columns of every plane
  return g.vn[0] == 64 ? launch_t2_cfg<64, 4, 16>(a, fma, s)
                       : (g.vn[0] == 32 ? launch_t2_cfg<32, 4, 8>(a, fma, s) : launch_t2_cfg<16, 4, 4>(a, fma, s));
}

}  // namespace sgc
