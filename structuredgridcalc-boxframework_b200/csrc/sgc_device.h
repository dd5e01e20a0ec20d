// sgc_device.h — device-side data model shared by the .cu files.
//
// Replaces the reference's device data model: CudaFab (CudaFab.H:104-530, an
// alias of one BaseFab), SymbolPair host/device mirrors per fab (BaseFab.cpp:509-528) and the
// __constant__ pointer tables of WavePatch_Cuda.cu:22-24 by ONE arena per level plus a device
// box table of PODs.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "sgc_internal.h"

namespace sgc {

// ---------------------------------------------------------------------------------------------
// Device layout of a fab ("split-x").  The reference stores a fab as one Fortran array over the
// box grown by nghost (BaseFab.H:277-313), so consecutive valid rows are separated by 2*nghost
// ghost cells and every row of a sweep's output starts and ends inside a 32-byte DRAM sector
// shared with ghost cells: measured on B200, writing such rows runs at 2.5 TB/s against 5.3 TB/s
// for dense rows (profiles/micro/store_patterns.cu).  On the device a fab is therefore two blocks:
//   main : x restricted to the VALID columns, y and z over the grown box:
//            main[c*cs + (k*n1 + j)*mx + (i - ngx)],  mx = n0 - 2*ngx
//          rows are dense (64 doubles = 512 B for a 64^3 box), so a sweep writes whole sectors;
//   xg   : the 2*ngx ghost columns in x, as [side/ghost layer g][k][j]:
//            xg[c*xcs + (g*n2 + k)*n1 + j],  g = i for the low side, i - mx for the high side
//          so the x-halo of one z-plane is one contiguous strip per side.
// (i,j,k) are fab-relative (0-based from the grown box's lower corner).  Host memory always uses
// the reference order; sgc_level_upload/download convert.  A dense buffer (pack buffers, the
// upload staging area) is the same struct with ngx = 0.
// ---------------------------------------------------------------------------------------------
struct FabGeom {
  long long off;   // element offset of the main block (comp 0) in the level arena
  long long xoff;  // element offset of the xg block (comp 0)
  long long cs;    // main component stride = mx*n1*n2
  long long xcs;   // xg component stride = 2*ngx*n1*n2
  int lo[3];       // lower corner of the fab box (= valid box grown by nghost)
  int n[3];        // extents of the fab box
  int vlo[3];      // lower corner of the valid box
  int vn[3];       // extents of the valid box
  int ngx;         // ghost columns per side kept in the xg block
  int mx;          // x extent of the main block = n[0] - 2*ngx
};

__host__ __device__ inline long long fab_index(const FabGeom& g, int i, int j, int k, int c) {
  const int vi = i - g.ngx;
  if (vi >= 0 && vi < g.mx) return g.off + c * g.cs + ((long long)k * g.n[1] + j) * g.mx + vi;
  const int gi = vi < 0 ? i : i - g.mx;
  return g.xoff + c * g.xcs + ((long long)gi * g.n[2] + k) * g.n[1] + j;
}

// One strided region copy dst[x,y,z,c] = src[x,y,z,c] — (a piece of) a Motion2Way, a
// BaseFab::copy, a linearOut/linearIn, an upload/download conversion or a physical-BC face copy.
struct CopyItem {
  long long dst;   // element offset of the region origin, first component, in the dst arena
  long long src;   // same in the src arena
  long long dcs, scs;   // component strides
  int nx, ny, nz, nc;   // region extents and number of components
  int dsx, dsy, dsz;    // dst strides (elements) along x, y, z
  int ssx, ssy, ssz;    // src strides
  unsigned flags;  // component mask, bit = absolute destination component (BaseFab.cpp:320)
  int c0;          // first destination component (for the mask test)
};

// Items of 8-byte elements whose rows are even-length, contiguous and 16-byte aligned on both sides are
// WIDE: they are stored first, in units of 16 bytes (offsets, x extent and strides halved), and run by
// the first `nwchunk` CTAs of the launch with 16-byte loads / stores.  "Cells" below are work units:
// 16-byte units [0, nwcell) of the wide items, then elements [nwcell, ncell) of the others.
struct CopyTable {
  CopyItem* d_items = nullptr;
  long long* d_cell_start = nullptr;   // exclusive prefix sum of item unit counts, n+1 entries
  int* d_chunk_first = nullptr;        // first item overlapping chunk c; one past each class a closing entry
  int nitem = 0;
  int nchunk = 0;                      // CTAs of the launch: nwchunk wide + the rest
  int nwitem = 0, nwchunk = 0;
  long long ncell = 0, nwcell = 0;
  bool strided_reads = false;          // a good share of the cells are single elements of long rows (x faces)
  void release();
  int upload(const std::vector<CopyItem>& items, int esz = 0);   // esz 8 enables the wide class
};

// geometry helpers (sgc_device.cu)
FabGeom level_fab_geom(const IBox& valid, int ng, long long off, long long xoff);
FabGeom dense_geom(const IBox& reg, long long off);
bool fab_contains(const FabGeom& g, const IBox& r);
// append the pieces of dst[dreg, dcomp..] = src[sreg, scomp..] (regions in index space, equal extents)
void make_items(std::vector<CopyItem>& out, const FabGeom& gd, const IBox& dreg, int dcomp, const FabGeom& gs,
                const IBox& sreg, int scomp, int nc, unsigned flags);

constexpr int kCopyChunk = 2048;       // cells per CTA of the batched copy kernel
constexpr int kCopyThreads = 256;

// Event-pair log for sgc_profile_* (class 0 = sweep, 1 = exchange copy, 2 = two-step sweep).
struct ProfLog {
  std::vector<cudaEvent_t> a, b;
  int used = 0;
};
struct ProfScope {          // RAII: records the pair around one launch when profiling is on
  int klass; cudaStream_t s; int slot;
  ProfScope(int klass_, cudaStream_t s_);
  ~ProfScope();
};

struct Runtime {
  void* d_stage = nullptr;              // reference-order staging area for upload / download
  size_t stage_bytes = 0;
  bool profiling = false;
  ProfLog prof[3];
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t compute = nullptr;
  cudaStream_t comm = nullptr;
  int prio_hi = 0;
  long long launches = 0;
  int sweep_grid_cap = 0;               // > 0: persistent sweeps use at most this many CTAs (leaves SMs to NCCL)
  // allocations other GPUs may still have mapped (exported level arenas, progress counters): freeing them
  // before every importer has closed its mapping is undefined, so they wait for sgc_comm_finalize /
  // sgc_shutdown, the points every rank passes
  std::vector<void*> deferred_free;
};
Runtime& rt();

// A captured unit of a step loop (a few launches on the compute stream that return the ping-pong levels to
// their starting roles), replayed as one CUDA graph: on the small levels of the reference's own applications
// a step is three ~10 us kernels and the loop is bound by launch latency.  Only pure launch sequences are
// captured: single rank, profiling off (sgc_device.cu: step_graph_ok).
struct StepGraph {
  cudaGraphExec_t exec = nullptr;
  long long launches_before = 0, unit = 0;
  int begin();                 // start capturing on the compute stream
  int end();                   // stop, instantiate; nothing has executed yet
  int replay(int times);
  ~StepGraph();
};
bool step_graph_ok(int nremote_a, int nremote_b = 0);

#define SGC_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::sgc::fail(SGC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define SGC_REQUIRE_RT()                                                                     \
  do {                                                                                      \
    if (!::sgc::rt().ready)                                                                 \
      return ::sgc::fail(SGC_ERR_CUDA, "sgc_init() has not succeeded: no CUDA device bound " \
                                       "(this library has no CPU fallback)");               \
  } while (0)

// Arguments of the fused ghost fill of the streaming sweep (sgc_stream.cu).
struct PushArgs {
  const int* nbr;        // [nbox][27] local box in direction code 13 + dx + 3 dy + 9 dz that has a local
                         // motion item with this box, or -1 (domain boundary, trimmed, or remote owner)
  double* out_level;     // arena base of the output level            (filled in by the launcher)
  const double* in_level;   // arena base of the input level          (filled in by the launcher)
  long long xg_base;     // element offset of the xg regions          (filled in by the launcher)
  int box0;              // local index of the first box of the launch (filled in by the launcher)
  int mask;              // bit 0 = push x faces, bit 1 = pull z planes, bit 2 = pull y rows
  // z halo straight out of the neighbouring GPU's memory (NVLink peer access, z-slab partitions):
  const int* rz;                             // [nbox][2] (-z, +z): (peer slot << 28) | box index on the peer, or -1
  const double* peer_in[2];                  // the peers' input-level arenas (cudaIpc-mapped)
  const unsigned long long* peer_flag[2];    // the peers' progress counters (cudaIpc-mapped)
  unsigned long long wait_value;             // counter value that makes the peers' input level readable
  unsigned int* err;                         // set to 1 when a wait timed out; a launch that finds it set does nothing
  unsigned long long wait_budget_ns;         // how long a wait may take (SGC_P2P_TIMEOUT_MS, default 30 s)
  // edge strips of the OUTPUT level (sgc_level_s::xe_base), or null: the threads that compute the first /
  // last cell of a row also store it there, so that the next exchange reads its x faces densely
  double* xe_out;
};

// Arguments of the two-steps-per-pass sweep (sgc_stream2.cu).
struct T2Args {
  const double* in_level;     // arena base of the level at time t
  double* out_level;          // arena base of the level that receives time t+2
  long long xe_in, xe_out;    // element offsets of the edge-strip blocks in the two arenas
  long long xe_spare_plane;   // strip plane index behind the last box: scratch for stores that must go nowhere
  long long dump;             // element offset (output arena) of a scratch plane for the same purpose
  long long xg_base;          // element offset of the x-ghost region of an arena
  // levels with faces that have no neighbour (non-null ghost[0] selects that instance): the arenas whose
  // ghost cells are read as level t (steps 0, 2, 4 .. of the run: level a) and as level t+1 (level b)
  const double* ghost[2];
  const double* peer_ghost[2][2];            // the same arenas on the z neighbours' GPUs, [peer slot][0/1]
  const int* nbr;             // [nbox][27], as PushArgs::nbr
  int nplanes;                // npil * (nbz * vz + 4) virtual planes
  int vz;                     // valid planes per box
  int npil, nbz;              // pillars: boxes p + k * npil, k = 0 .. nbz-1, are stacked in z (nbz = 1: box by box)
  double f;
  // z neighbours in the neighbouring GPU's memory (z-slab partitions), as PushArgs
  const int* rz;
  const double* peer_in[2];
  const unsigned long long* peer_flag[2];
  unsigned long long wait_value;
  unsigned int* err;                         // as PushArgs::err
  unsigned long long wait_budget_ns;
};

// launchers implemented in sgc_kernels.cu
int launch_copy(const CopyTable& t, void* dst_arena, const void* src_arena, int esz, cudaStream_t s);
int launch_fill(void* p, long long n, int esz, const void* value, cudaStream_t s);
int launch_convert_level(void* arena, void* stage, long long nrows, int mx, int n1, int n2, int ncomp, long long xg_base,
                         bool to_arena, cudaStream_t s);
int launch_fill_fabs(void* arena, const FabGeom* d_geom, int nbox, long long max_elems, int c0, int nc, int esz,
                     const void* value, cudaStream_t s);

}  // namespace sgc

// opaque handle bodies
struct sgc_level_s {
  int nbox = 0, ncomp = 0, ng = 0, esz = 8, gidx0 = 0;
  std::vector<sgc::IBox> boxes;         // valid boxes
  std::vector<long long> off;           // element offset of fab b in the host concatenation
  long long elems = 0;                  // total elements, all fabs and comps
  long long lead = 0;                   // elements of padding before fab 0 in the arena
  void* d_base = nullptr;               // cudaMalloc'ed block
  void* d_arena = nullptr;              // = d_base + lead*esz; main blocks of all fabs, then all xg blocks
  long long main_elems = 0;             // elements in the main region (xg region starts here, padded)
  long long xg_base = 0;                // element offset of the xg region in the arena
  long long arena_elems = 0;            // main + xg (+ padding)
  // edge strips [box][kz][4][vny] (copies of the valid columns 0, 1, vnx-2, vnx-1) read by the two-step
  // sweep (sgc_stream2.cu); allocated behind the arena for levels that kernel covers, else xe_elems = 0
  long long xe_base = 0, xe_elems = 0;
  // which strips mirror the valid cells right now: 0 none, 1 the outer columns (0 and vnx-1: what an
  // exchange needs for its x faces), 2 all four.  Every writer of valid cells either keeps them current
  // (the streaming sweeps, k_build_xe) or resets this (sgc::level_written); handing out the raw device
  // pointer (sgc_level_device_ptr) turns the mechanism off for the level for good.
  int xe_valid = 0;
  bool xe_extern = false;
  bool exported = false;                // the arena has been exported with cudaIpcGetMemHandle (peer GPUs may map it)
  unsigned long long generation = 0;    // unique per created level (peer-memory sessions are keyed on it)
  sgc::CopyTable conv_in, conv_out;     // reference-order staging <-> split-x arena, whole level
  sgc::CopyTable conv_valid;            // arena -> [box][comp] valid cells only (plot output), built on first use
  sgc::CopyTable conv_valid_in;         // the reverse: dense valid cells -> arena (ghost cells untouched)
  long long valid_elems = 0;
  sgc::FabGeom* d_geom = nullptr;
  std::vector<sgc::FabGeom> geom;
  bool uniform = true;                  // all fabs have the same extents
  // sweep work table (box, y0, z0) for the generic kernel, built at creation
  int* d_work = nullptr;
  int nwork = 0;
  int tile_z = 16;                      // planes per work tile
  std::vector<int> work_begin;          // first work item of box b (nbox+1 entries)
  // scratch device buffer for linearOut/In
  void* d_scratch = nullptr;
  size_t scratch_bytes = 0;
};

namespace sgc {
inline void level_written(sgc_level_s* l) { l->xe_valid = 0; }      // valid cells changed behind the strips' back
}

struct sgc_plan_s {
  // geometry signature of the level the plan was built for
  int nbox = 0, ncomp = 0, ng = 0, esz = 0, gidx0 = 0;
  std::vector<sgc::IBox> boxes;
  int c0 = 0, nc = 0;
  int ng_exchange = 0;                  // ghost width the items fill (<= ng: Copier::defineExchangeDBL's a_numGhost)
  int nproc = 1, rank = 0;
  std::vector<sgc_motion_item> items;   // this rank's items, reference order
  int nremote = 0;
  sgc::CopyTable local;                 // local items: arena -> arena
  // the same items with every x-face / x-edge source column read from the level's edge strips (dense)
  // instead of one 8-byte element per row of the neighbour's main block; used when level->xe_valid
  sgc::CopyTable local_xe;
  bool ghost_only = false;              // every item writes ghost cells only (exchange and BC plans)
  // the uniform layout the plan was derived from (sgc_plan_create_exchange), for re-boxing
  bool has_grid = false;
  int dlo[3] = {0, 0, 0}, dhi[3] = {0, 0, 0}, maxbox[3] = {0, 0, 0};
  unsigned periodic = 0, trim = 0;
  // multi-rank: per-peer staging
  struct Peer {
    int rank = -1;
    sgc::CopyTable pack;                // arena -> send buffer   (dense, receiver's item order)
    sgc::CopyTable unpack;              // recv buffer -> arena
    void* d_send = nullptr;
    void* d_recv = nullptr;
    long long send_elems = 0, recv_elems = 0;
    cudaEvent_t ev_arrived = nullptr;   // this peer's message has landed in d_recv (recorded on the comm stream)
    // the same ghost regions filled by READING the peer's arena where it lies (peer-memory session of a step loop):
    // dst = offsets in this rank's arena, src = offsets in the peer's arena (uniform levels: same layout everywhere)
    sgc::CopyTable pull;
  };
  std::vector<Peer> peers;
  int n_boundary_boxes = 0;               // boxes with at least one remote item
  std::vector<char> box_is_boundary;
  cudaEvent_t ev_halo = nullptr, ev_ready = nullptr;    // halo received / data to exchange produced
  bool in_flight = false;
  // fused ghost fill (sweep pushes boundary cells into the neighbours' ghosts): destination local
  // box per (box, direction code), -1 where there is no local motion item
  int* d_push = nullptr;
  bool fusable = false;
  // two-steps-per-pass sweep (sgc_stream2.cu): every box has its 8 in-plane neighbours locally, and its
  // two z neighbours locally (t2_local) or at least in the plan (t2_slab: the missing ones are remote)
  bool t2_local = false, t2_slab = false;
  bool t2_bnd = false;                  // as t2_slab, but some faces have no motion item at all (physical boundary)
  int t2_agreed = -1;                   // multi-rank: every rank can run the two-step passes (-1: not asked yet)
  int t2_npil = 0, t2_nbz = 1;          // pillar structure of the level's boxes under this plan (sgc::pillar_structure)
  // library-owned third level of the two-step passes (single rank): with it an ODD number of passes still ends in the
  // level the step-by-step schedule leaves the result in (a -> b -> scratch -> a), so a run of n steps is (n-2)/2
  // passes + the two plain steps instead of an even number of passes + up to three fused one-step sweeps
  sgc_level_s* t2_scratch = nullptr;
  // peer-memory halo session (sgc_comm.cu): the two ping-pong levels of a step loop, their arenas on
  // the z neighbours mapped through cudaIpc, and progress counters
  struct P2P {
    bool tried = false, ready = false;
    sgc_level_s* lev[3] = {nullptr, nullptr, nullptr};   // the two ping-pong levels + the plan's third level (or null)
    int npeer = 0;
    int peer_rank[2] = {-1, -1};
    void* peer_base[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // [peer slot][level] opened allocations
    void* peer_flag_base[2] = {nullptr, nullptr};
    unsigned long long* d_flag = nullptr;    // this rank's counter (cudaMalloc, exported)
    unsigned int* d_err = nullptr;
    int* d_rz = nullptr;                     // [nbox][2]
    unsigned long long counter = 0;          // value last signalled
    unsigned long long gen[3] = {0, 0, 0};   // generations of the levels (a recycled address is not the same level)
  } p2p;
  // Re-boxed step loop (sgc_device.cu: rebox_setup): a level of many small boxes (16^3, 32^3, 8^3) is
  // advanced on a library-owned pair of levels whose device fabs are 64^2-footprint unions of the user's
  // boxes, so that the two-step sweep runs at its 64^3-box speed; the user's levels receive the result
  // and run the last steps themselves.  Built on first use, lives as long as the plan.
  struct Rebox {
    bool tried = false, ok = false;
    sgc_level_s* lev[2] = {nullptr, nullptr};
    sgc_plan_s* plan = nullptr;
    sgc::CopyTable f2c, f2c_ghost, c2f;      // user's fabs -> union fabs (valid cells; physical ghost faces) and back
    int factor[3] = {1, 1, 1};
  } rebox;
};

struct sgc_event_s {
  cudaEvent_t ev = nullptr;
};

struct sgc_copyset_s {
  // geometry signatures of the two levels the set was recorded on
  int nbox[2] = {0, 0}, ncomp[2] = {0, 0}, ng[2] = {0, 0}, esz = 0;
  std::vector<sgc::IBox> boxes[2];
  sgc::CopyTable table;
  long long ncell = 0;
};

struct sgc_stream_s {
  cudaStream_t s = nullptr;
  void* d_stage = nullptr;      // reference-order staging area of this stream's transfers
  size_t stage_bytes = 0;
};
