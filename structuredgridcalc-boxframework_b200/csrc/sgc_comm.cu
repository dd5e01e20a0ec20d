// sgc_comm.cu — inter-GPU ghost exchange: one process per GPU, ranks := the reference's MPI ranks.
//
// Replaces the reference's message layer: MPI_Isend/MPI_Irecv of one packed host
// buffer per (box, neighbour direction) with tag 27*globalIdx+dir (Copier.H:388-411,
// LevelData.H:488-497) followed by MPI_Waitall and linearIn (LevelData.H:540-561).  Here every
// (peer) pair moves ONE aggregated device buffer per exchange: a batched pack kernel gathers all
// regions a peer needs into a staging buffer in the RECEIVER's motion-item order (each region in
// linearOut wire order comp->k->j->i), ncclSend/ncclRecv inside one group move it over NVLink on
// the comm stream, and a batched unpack kernel scatters it into the ghost cells.
//
// NCCL is resolved with dlopen at sgc_comm_init so that (a) the library loads on machines without
// NCCL (single-GPU use, CPU-only symbol tests) and (b) inside a torch process the already-loaded
// libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>   // types only; no link-time dependency

#include <algorithm>
#include <tuple>
#include <cstdlib>
#include <cstring>

#include "sgc_device.h"
#include "sgc_pipe.cuh"

namespace sgc {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Comm {
  NcclApi api;
  ncclComm_t comm = nullptr;
  int nproc = 1, rank = 0;
  bool ready = false;
};
static Comm& cm() {
  static Comm c;
  return c;
}

static int load_nccl() {
  NcclApi& a = cm().api;
  if (a.handle) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) return fail(SGC_ERR_COMM, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SGC_SYM(field, name)                                                         \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name));              \
  if (!a.field) return fail(SGC_ERR_COMM, std::string("libnccl lacks symbol ") + name)
  SGC_SYM(GetUniqueId, "ncclGetUniqueId");
  SGC_SYM(CommInitRank, "ncclCommInitRank");
  SGC_SYM(CommDestroy, "ncclCommDestroy");
  SGC_SYM(Send, "ncclSend");
  SGC_SYM(Recv, "ncclRecv");
  SGC_SYM(GroupStart, "ncclGroupStart");
  SGC_SYM(GroupEnd, "ncclGroupEnd");
  SGC_SYM(AllReduce, "ncclAllReduce");
  SGC_SYM(GetErrorString, "ncclGetErrorString");
#undef SGC_SYM
  return 0;
}

#define SGC_NCCL(expr)                                                                             \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != ncclSuccess)                                                                         \
      return fail(SGC_ERR_COMM, std::string(#expr) + ": " + cm().api.GetErrorString(_r));          \
  } while (0)

static IBox box_of(const int lo[3], const int hi[3]) {
  IBox b;
  for (int d = 0; d < 3; ++d) { b.lo[d] = lo[d]; b.hi[d] = hi[d]; }
  return b;
}

// Host-side description of what one rank sends to / receives from each peer; pure index algebra,
// shared with the CPU (gloo) tests through sgc_halo_describe().
struct HaloList {
  int peer;
  std::vector<sgc_motion_item> recv;   // my items whose source box lives on `peer`, my order
  std::vector<sgc_motion_item> send;   // peer's items whose source box lives on me, peer's order
};

static int build_halo_lists(const Grid& g, int ng, unsigned periodic, unsigned trim, int rank,
                            const std::vector<sgc_motion_item>& mine, std::vector<HaloList>& out) {
  std::vector<int> peers;
  for (const sgc_motion_item& it : mine)
    if (it.send_proc != rank) peers.push_back(it.send_proc);
  std::sort(peers.begin(), peers.end());
  peers.erase(std::unique(peers.begin(), peers.end()), peers.end());
  for (int q : peers) {
    HaloList h;
    h.peer = q;
    for (const sgc_motion_item& it : mine)
      if (it.send_proc == q) h.recv.push_back(it);
    std::vector<sgc_motion_item> theirs;
    const int st = build_exchange_items(g, ng, periodic, trim, q, theirs);
    if (st) return st;
    for (const sgc_motion_item& it : theirs)
      if (it.send_proc == rank) h.send.push_back(it);
    out.push_back(std::move(h));
  }
  return 0;
}

int comm_plan_setup(sgc_plan_s* p, sgc_level_s* l, const Grid& g, unsigned periodic, unsigned trim) {
  if (!cm().ready || cm().nproc != p->nproc || cm().rank != p->rank)
    return fail(SGC_ERR_COMM, "plan has remote motion items but sgc_comm_init(nproc, rank, id) has not been called "
                              "with the same nproc/rank");
  std::vector<HaloList> lists;
  int st = build_halo_lists(g, p->ng_exchange > 0 ? p->ng_exchange : l->ng, periodic, trim, p->rank, p->items, lists);
  if (st) return st;
  for (const HaloList& h : lists) {
    sgc_plan_s::Peer peer;
    peer.rank = h.peer;
    std::vector<CopyItem> unpack, pack;
    long long pos = 0;
    for (const sgc_motion_item& it : h.recv) {
      const IBox r = box_of(it.recv_lo, it.recv_hi);
      const int ld = it.recv_box - l->gidx0;
      make_items(unpack, l->geom[ld], r, p->c0, dense_geom(r, pos), r, 0, p->nc, it.comp_flags);
      pos += r.size() * p->nc;
    }
    peer.recv_elems = pos;
    if (l->uniform) {
      // pull table: every rank holds the same sequence of fab shapes, so box `send_box` sits in its owner's arena
      // where our own box of the same local index sits in ours
      std::vector<CopyItem> pull;
      const FabGeom& g0 = l->geom[0];
      for (const sgc_motion_item& it : h.recv) {
        const IBox r = box_of(it.recv_lo, it.recv_hi), sr = box_of(it.src_lo, it.src_hi);
        const int ld = it.recv_box - l->gidx0, rb = it.send_box - it.send_proc * g.bpp;
        const FabGeom gp = level_fab_geom(g.box(it.send_box), l->ng, (long long)rb * g0.cs * l->ncomp,
                                          l->xg_base + (long long)rb * g0.xcs * l->ncomp);
        make_items(pull, l->geom[ld], r, p->c0, gp, sr, p->c0, p->nc, it.comp_flags);
      }
      st = peer.pull.upload(pull, l->esz);
      if (st) return st;
    }
    pos = 0;
    for (const sgc_motion_item& it : h.send) {
      const IBox r = box_of(it.src_lo, it.src_hi);
      const int ls = it.send_box - l->gidx0;
      if (ls < 0 || ls >= l->nbox) return fail(SGC_ERR_COMM, "halo: peer asks for a box this rank does not own");
      make_items(pack, dense_geom(r, pos), r, 0, l->geom[ls], r, p->c0, p->nc, ~0u);
      pos += r.size() * p->nc;
    }
    peer.send_elems = pos;
    st = peer.pack.upload(pack, l->esz);
    if (!st) st = peer.unpack.upload(unpack, l->esz);
    if (st) return st;
    SGC_CUDA(cudaMalloc(&peer.d_send, (size_t)std::max<long long>(peer.send_elems, 1) * l->esz));
    SGC_CUDA(cudaMalloc(&peer.d_recv, (size_t)std::max<long long>(peer.recv_elems, 1) * l->esz));
    p->peers.push_back(peer);
  }
  // Message order: the same TOTAL order of rank pairs on every rank (colour of the pair, then the pair), so
  // that the per-peer send/recv groups below can never wait for each other in a cycle; the colour (parity of
  // the lower rank) lets a ring of slabs finish in two or three rounds instead of a chain through all ranks.
  const int me = p->rank;
  std::sort(p->peers.begin(), p->peers.end(), [me](const sgc_plan_s::Peer& x, const sgc_plan_s::Peer& y) {
    auto key = [me](int q) { const int lo = std::min(me, q), hi = std::max(me, q); return std::make_tuple(lo & 1, lo, hi); };
    return key(x.rank) < key(y.rank);
  });
  for (sgc_plan_s::Peer& q : p->peers) SGC_CUDA(cudaEventCreateWithFlags(&q.ev_arrived, cudaEventDisableTiming));
  SGC_CUDA(cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));
  SGC_CUDA(cudaEventCreateWithFlags(&p->ev_halo, cudaEventDisableTiming));
  return 0;
}

void comm_plan_release(sgc_plan_s* p) {
  for (sgc_plan_s::Peer& q : p->peers) {
    q.pack.release();
    q.unpack.release();
    q.pull.release();
    if (q.d_send) cudaFree(q.d_send);
    if (q.d_recv) cudaFree(q.d_recv);
    if (q.ev_arrived) cudaEventDestroy(q.ev_arrived);
  }
  p->peers.clear();
  if (p->ev_ready) cudaEventDestroy(p->ev_ready);
  if (p->ev_halo) cudaEventDestroy(p->ev_halo);
  p->ev_ready = p->ev_halo = nullptr;
}

// LevelData::exchangeBegin's remote half (LevelData.H:577-626): pack every peer's message, then ONE send/recv
// pair per peer, each followed by its own event — the device-side form of the reference's MPI_Waitany loop
// (LevelData.H:507-536): comm_exchange_end unpacks a peer's message as soon as THAT message has arrived, while
// later peers' transfers are still in flight.  SGC_COMM_ONE_GROUP=1 restores the single group for all peers.
int comm_exchange_begin(sgc_level_s* l, sgc_plan_s* p) {
  Runtime& r = rt();
  Comm& c = cm();
  if (!c.ready) return fail(SGC_ERR_COMM, "sgc_comm_init has not been called");
  // the comm stream may start once the data being exchanged has been produced
  SGC_CUDA(cudaEventRecord(p->ev_ready, r.compute));
  SGC_CUDA(cudaStreamWaitEvent(r.comm, p->ev_ready, 0));
  for (sgc_plan_s::Peer& q : p->peers) {
    const int st = launch_copy(q.pack, q.d_send, l->d_arena, l->esz, r.comm);
    if (st) return st;
  }
  static const char* og = getenv("SGC_COMM_ONE_GROUP");
  const bool one_group = og && og[0] == '1';
  if (one_group) SGC_NCCL(c.api.GroupStart());
  for (sgc_plan_s::Peer& q : p->peers) {
    if (!one_group) SGC_NCCL(c.api.GroupStart());
    if (q.send_elems) SGC_NCCL(c.api.Send(q.d_send, (size_t)q.send_elems * l->esz, ncclChar, q.rank, c.comm, r.comm));
    if (q.recv_elems) SGC_NCCL(c.api.Recv(q.d_recv, (size_t)q.recv_elems * l->esz, ncclChar, q.rank, c.comm, r.comm));
    if (!one_group) {
      SGC_NCCL(c.api.GroupEnd());
      SGC_CUDA(cudaEventRecord(q.ev_arrived, r.comm));
    }
  }
  if (one_group) {
    SGC_NCCL(c.api.GroupEnd());
    for (sgc_plan_s::Peer& q : p->peers) SGC_CUDA(cudaEventRecord(q.ev_arrived, r.comm));
  }
  SGC_CUDA(cudaEventRecord(p->ev_halo, r.comm));
  return 0;
}

int comm_exchange_end(sgc_level_s* l, sgc_plan_s* p) {
  Runtime& r = rt();
  for (sgc_plan_s::Peer& q : p->peers) {
    SGC_CUDA(cudaStreamWaitEvent(r.compute, q.ev_arrived, 0));       // this peer's message only
    const int st = launch_copy(q.unpack, l->d_arena, q.d_recv, l->esz, r.compute);
    if (st) return st;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Peer-memory halo session for a fused step loop on a z-slab partition.
//
// The reference moves a halo as: pack on the sender, MPI message, unpack on the receiver
// (LevelData.H:488-561).  On one NVSwitch node the receiving GPU can instead READ the neighbour's
// cells where they lie: the arenas of the two ping-pong levels are exported with cudaIpcGetMemHandle,
// the handles travel once (ncclSend/Recv), and the sweep's TMA producer bulk-copies the neighbour's
// first / last valid plane over NVLink into its ring slot.  Ordering: after every sweep a rank bumps a
// counter in its own memory (k_signal, stream-ordered behind the sweep); the neighbour's producer
// lane spins on that counter (volatile system-scope loads, bounded) before its first remote copy.
// Each kernel only ever waits for a sweep the neighbour launched earlier without waiting on us, so
// there is no cycle; the planes a neighbour reads are exactly those whose computation needed the
// remote plane, so they are never overwritten too early.
// ---------------------------------------------------------------------------------------------
__global__ void k_signal(unsigned long long* flag, unsigned long long value, const unsigned int* err) {
  // after a timed-out wait the phase is dead: do not tell the neighbours that data they cannot trust is ready
  if (*reinterpret_cast<const volatile unsigned int*>(err)) return;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(flag) = value;
  __threadfence_system();
}

// plans that hold a peer-memory session (so that destroying a level can invalidate the sessions that name it)
static std::vector<sgc_plan_s*>& p2p_plans() {
  static std::vector<sgc_plan_s*> v;
  return v;
}

struct DevTmp {              // device scratch freed on every path out of a function
  void* p = nullptr;
  ~DevTmp() { if (p) cudaFree(p); }
};

struct P2PMsg {
  cudaIpcMemHandle_t lev[3];
  int nlev;
  cudaIpcMemHandle_t flag;
  unsigned long long counter;
};

void comm_p2p_release(sgc_plan_s* p) {
  sgc_plan_s::P2P& S = p->p2p;
  if (S.tried && rt().ready) { cudaStreamSynchronize(rt().compute); cudaStreamSynchronize(rt().comm); }
  // closing OUR mappings of the neighbours' memory is always legal
  for (int k = 0; k < 2; ++k) {
    for (int l = 0; l < 3; ++l)
      if (S.peer_base[k][l]) { cudaIpcCloseMemHandle(S.peer_base[k][l]); S.peer_base[k][l] = nullptr; }
    if (S.peer_flag_base[k]) { cudaIpcCloseMemHandle(S.peer_flag_base[k]); S.peer_flag_base[k] = nullptr; }
  }
  // our counter may still be mapped by the neighbours: its memory waits for a point every rank passes
  if (S.d_flag) rt().deferred_free.push_back(S.d_flag);
  if (S.d_err) cudaFree(S.d_err);
  if (S.d_rz) cudaFree(S.d_rz);
  S = sgc_plan_s::P2P();
  std::vector<sgc_plan_s*>& reg = p2p_plans();
  reg.erase(std::remove(reg.begin(), reg.end(), p), reg.end());
}

// sgc_level_destroy: sessions that name the level are over (a later level at the same address is NOT the same level)
void comm_level_destroyed(sgc_level_s* l) {
  const std::vector<sgc_plan_s*> reg = p2p_plans();      // copy: release edits the registry
  for (sgc_plan_s* p : reg)
    for (int k = 0; k < 3; ++k)
      if (p->p2p.lev[k] == l && p->p2p.gen[k] == l->generation) { comm_p2p_release(p); break; }
}

void comm_drain_deferred_frees() {
  for (void* q : rt().deferred_free) cudaFree(q);
  rt().deferred_free.clear();
}

// All ranks meet on stream `s` (a one-int all-reduce, no host synchronisation): no kernel a rank enqueues behind
// it starts before every rank has enqueued its own, which bounds the skew the peer-counter waits ever see
int comm_stream_barrier(cudaStream_t s) {
  Comm& c = cm();
  if (!c.ready || c.nproc < 2) return 0;
  static int* d_one = nullptr;
  if (!d_one) { SGC_CUDA(cudaMalloc(&d_one, 256)); SGC_CUDA(cudaMemset(d_one, 0, 256)); }
  SGC_NCCL(c.api.AllReduce(d_one, d_one, 1, ncclInt, ncclMax, c.comm, s));
  return 0;
}

__global__ void k_wait_peer(const unsigned long long* flag, unsigned long long value, unsigned long long budget_ns,
                            unsigned int* err) {
  if (*reinterpret_cast<const volatile unsigned int*>(err)) return;
  if (!wait_peer_counter(flag, value, budget_ns)) *err = 1u;
  __threadfence_system();
}

// can the remote half of an exchange of `l` be done by reading the neighbours' arenas? (session up, l is one of its two
// levels, every peer of the plan is a session peer and has a pull table)
bool comm_p2p_can_pull(sgc_plan_s* p, sgc_level_s* l) {
  const sgc_plan_s::P2P& S = p->p2p;
  if (!S.ready || (l != S.lev[0] && l != S.lev[1])) return false;
  for (const sgc_plan_s::Peer& q : p->peers) {
    bool found = false;
    for (int k = 0; k < S.npeer; ++k) found = found || S.peer_rank[k] == q.rank;
    if (!found || q.pull.nitem == 0) return false;
  }
  return true;
}

// The remote half of LevelData::exchange inside a step loop on a slab partition, without a message: wait (on the
// device) until the neighbour's progress counter says its level holds the step being exchanged, then ONE batched copy
// per neighbour reads the regions out of its arena over NVLink straight into our ghost cells.  No pack, no NCCL, no
// unpack, no host synchronisation.  Ordering against the neighbour's later writes: see DESIGN.md §5.
int comm_p2p_pull(sgc_plan_s* p, sgc_level_s* l, unsigned long long budget_ns, cudaStream_t s) {
  sgc_plan_s::P2P& S = p->p2p;
  const int li = (l == S.lev[0]) ? 0 : 1;
  for (sgc_plan_s::Peer& q : p->peers) {
    int slot = 0;
    for (int k = 0; k < S.npeer; ++k) if (S.peer_rank[k] == q.rank) slot = k;
    k_wait_peer<<<1, 1, 0, s>>>((const unsigned long long*)S.peer_flag_base[slot], S.counter, budget_ns, S.d_err);
    rt().launches++;
    SGC_CUDA(cudaGetLastError());
    const void* peer_arena = (const char*)S.peer_base[slot][li] + l->lead * l->esz;
    const int st = launch_copy(q.pull, l->d_arena, peer_arena, l->esz, s);
    if (st) return st;
  }
  return 0;
}

int comm_p2p_signal(sgc_plan_s* p, unsigned long long value, cudaStream_t s) {
  k_signal<<<1, 1, 0, s>>>(p->p2p.d_flag, value, p->p2p.d_err);
  rt().launches++;
  SGC_CUDA(cudaGetLastError());
  return 0;
}

int comm_p2p_check(sgc_plan_s* p) {
  unsigned int err = 0;
  SGC_CUDA(cudaMemcpyAsync(&err, p->p2p.d_err, sizeof err, cudaMemcpyDeviceToHost, rt().compute));
  SGC_CUDA(cudaStreamSynchronize(rt().compute));
  if (err) return fail(SGC_ERR_COMM, "peer-memory halo: timed out waiting for a neighbouring GPU's sweep");
  return 0;
}

// min over all ranks of a small integer (collective; used to take the same code path everywhere)
int comm_all_min(int v, int* out) {
  Comm& c = cm();
  *out = v;
  if (!c.ready || c.nproc < 2) return 0;
  Runtime& r = rt();
  DevTmp buf;
  SGC_CUDA(cudaMalloc(&buf.p, 256));
  int* d = (int*)buf.p;
  SGC_CUDA(cudaMemcpy(d, &v, sizeof v, cudaMemcpyHostToDevice));
  SGC_NCCL(c.api.AllReduce(d, d, 1, ncclInt, ncclMin, c.comm, r.comm));
  SGC_CUDA(cudaStreamSynchronize(r.comm));
  SGC_CUDA(cudaMemcpy(out, d, sizeof v, cudaMemcpyDeviceToHost));
  return 0;
}

// Returns 0 when the call went through (p->p2p.ready tells whether the session is usable on EVERY
// rank — decided by an all-reduce, so all ranks take the same path); < 0 on hard errors.
int comm_p2p_setup(sgc_plan_s* p, sgc_level_s* a, sgc_level_s* b, sgc_level_s* third) {
  sgc_plan_s::P2P& S = p->p2p;
  // the session is tied to the LEVELS (generation ids), not to their addresses
  const bool same_third = S.lev[2] == third && (!third || S.gen[2] == third->generation);
  if (S.tried && same_third &&
      ((S.lev[0] == a && S.gen[0] == a->generation && S.lev[1] == b && S.gen[1] == b->generation) ||
       (S.lev[0] == b && S.gen[0] == b->generation && S.lev[1] == a && S.gen[1] == a->generation)))
    return 0;
  Comm& c = cm();
  if (!c.ready || c.nproc < 2) return 0;
  comm_p2p_release(p);
  S.tried = true;
  S.lev[0] = a; S.gen[0] = a->generation;
  S.lev[1] = b; S.gen[1] = b->generation;
  S.lev[2] = third; S.gen[2] = third ? third->generation : 0;
  p2p_plans().push_back(p);
  Runtime& r = rt();
  // 1. which boxes have a z neighbour on another rank, and is every remote FACE a z face?
  int ok = 1;
  std::vector<int> rz((size_t)a->nbox * 2, -1);
  const int bpp = a->nbox;
  for (const sgc_motion_item& it : p->items) {
    if (it.recv_proc == it.send_proc) continue;
    const int codim = abs(it.dir[0]) + abs(it.dir[1]) + abs(it.dir[2]);
    if (codim != 1) continue;                       // edges / corners are not needed inside the loop
    if (it.dir[2] == 0) { ok = 0; break; }          // a remote x or y face: keep the NCCL schedule
    int slot = -1;
    for (int k = 0; k < S.npeer; ++k) if (S.peer_rank[k] == it.send_proc) slot = k;
    if (slot < 0) {
      if (S.npeer == 2) { ok = 0; break; }
      slot = S.npeer++;
      S.peer_rank[slot] = it.send_proc;
    }
    const int rbox = it.send_box - it.send_proc * bpp;
    rz[(size_t)(it.recv_box - a->gidx0) * 2 + (it.dir[2] < 0 ? 0 : 1)] = (slot << 28) | rbox;
  }
  // 2. our side of the session
  P2PMsg mine, theirs[2];
  memset(&mine, 0, sizeof mine);
  if (ok) {
    ok = cudaMalloc(&S.d_flag, 256) == cudaSuccess && cudaMalloc(&S.d_err, 256) == cudaSuccess &&
         cudaMalloc(&S.d_rz, sizeof(int) * rz.size()) == cudaSuccess;
    if (ok) {
      cudaMemset(S.d_flag, 0, 256);
      cudaMemset(S.d_err, 0, 256);
      cudaMemcpy(S.d_rz, rz.data(), sizeof(int) * rz.size(), cudaMemcpyHostToDevice);
      ok = cudaIpcGetMemHandle(&mine.lev[0], a->d_base) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine.lev[1], b->d_base) == cudaSuccess &&
           cudaIpcGetMemHandle(&mine.flag, S.d_flag) == cudaSuccess;
      mine.nlev = 2;
      if (ok && third) { ok = cudaIpcGetMemHandle(&mine.lev[2], third->d_base) == cudaSuccess; mine.nlev = 3; }
      if (ok) { a->exported = b->exported = true; if (third) third->exported = true; }
    }
    cudaGetLastError();
  }
  // 3. agree globally before anything rank-to-rank happens
  DevTmp ok_buf, msg_buf;
  SGC_CUDA(cudaMalloc(&ok_buf.p, 256));
  int* d_ok = (int*)ok_buf.p;
  SGC_CUDA(cudaMemcpy(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
  SGC_NCCL(c.api.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c.comm, r.comm));
  SGC_CUDA(cudaStreamSynchronize(r.comm));
  SGC_CUDA(cudaMemcpy(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost));
  if (ok) {
    // 4. swap handles with the z neighbours and map their memory
    SGC_CUDA(cudaMalloc(&msg_buf.p, sizeof(P2PMsg) * 3));
    P2PMsg* d_msg = (P2PMsg*)msg_buf.p;
    SGC_CUDA(cudaMemcpy(d_msg, &mine, sizeof mine, cudaMemcpyHostToDevice));
    SGC_NCCL(c.api.GroupStart());
    for (int k = 0; k < S.npeer; ++k) {
      SGC_NCCL(c.api.Send(d_msg, sizeof(P2PMsg), ncclChar, S.peer_rank[k], c.comm, r.comm));
      SGC_NCCL(c.api.Recv(d_msg + 1 + k, sizeof(P2PMsg), ncclChar, S.peer_rank[k], c.comm, r.comm));
    }
    SGC_NCCL(c.api.GroupEnd());
    SGC_CUDA(cudaStreamSynchronize(r.comm));
    SGC_CUDA(cudaMemcpy(theirs, d_msg + 1, sizeof(P2PMsg) * 2, cudaMemcpyDeviceToHost));
    for (int k = 0; k < S.npeer && ok; ++k) {
      if (theirs[k].nlev != mine.nlev) ok = 0;           // every rank takes the same decisions; a mismatch is a bug
      for (int l = 0; l < mine.nlev && ok; ++l)
        ok = cudaIpcOpenMemHandle(&S.peer_base[k][l], theirs[k].lev[l], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
      if (ok) ok = cudaIpcOpenMemHandle(&S.peer_flag_base[k], theirs[k].flag, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
    }
    cudaGetLastError();
    // 5. and agree again (a failed mapping on one rank must switch every rank back to NCCL)
    SGC_CUDA(cudaMemcpy(d_ok, &ok, sizeof ok, cudaMemcpyHostToDevice));
    SGC_NCCL(c.api.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c.comm, r.comm));
    SGC_CUDA(cudaStreamSynchronize(r.comm));
    SGC_CUDA(cudaMemcpy(&ok, d_ok, sizeof ok, cudaMemcpyDeviceToHost));
  }
  S.ready = ok != 0;
  S.counter = 0;
  return 0;
}

}  // namespace sgc

using namespace sgc;

extern "C" {

int sgc_comm_unique_id(char id128[128]) {
  if (!id128) return fail(SGC_ERR_ARG, "null id");
  const int st = load_nccl();
  if (st) return st;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  SGC_NCCL(cm().api.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return 0;
}

int sgc_comm_init(int nproc, int rank, const char id128[128]) {
  SGC_REQUIRE_RT();
  if (nproc < 1 || rank < 0 || rank >= nproc) return fail(SGC_ERR_ARG, "sgc_comm_init: bad nproc/rank");
  Comm& c = cm();
  if (c.ready) return fail(SGC_ERR_ARG, "sgc_comm_init: already initialised");
  c.nproc = nproc;
  c.rank = rank;
  if (nproc > 1) {
    if (!id128) return fail(SGC_ERR_ARG, "sgc_comm_init: null id");
    const int st = load_nccl();
    if (st) return st;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    SGC_NCCL(c.api.CommInitRank(&c.comm, nproc, id, rank));
  }
  c.ready = true;
  return 0;
}

int sgc_comm_finalize(void) {
  Comm& c = cm();
  if (!c.ready) return 0;
  if (rt().ready) cudaDeviceSynchronize();
  if (c.comm && rt().ready) {
    // every rank is here: sessions are over, nobody launches kernels that read a neighbour any more, and the
    // allocations neighbours had mapped can go (their mappings were closed with their plans, or die with them)
    int all = 0;
    comm_all_min(1, &all);
    comm_drain_deferred_frees();
  }
  if (c.comm) c.api.CommDestroy(c.comm);
  c.comm = nullptr;
  c.ready = false;
  c.nproc = 1;
  c.rank = 0;
  return 0;
}

int sgc_comm_nproc(void) { return cm().nproc; }
int sgc_comm_rank(void) { return cm().rank; }

/* Host-only description of the aggregated halo messages of rank `rank` (no GPU, no NCCL): for
 * peer number `peer_index` (peers sorted by rank) returns the peer's rank and writes the motion
 * items in wire order: kind 0 = what this rank RECEIVES from the peer (its own items, own order),
 * kind 1 = what it SENDS (the peer's items that read from this rank, peer's order).  Returns the
 * item count, or SGC_ERR_ARG when peer_index is past the last peer.  Used by the CPU (gloo) tests
 * of the multi-rank host logic and by the C++ facade.                                         */
int sgc_halo_describe(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost, unsigned periodic,
                      unsigned trim, int nproc, int rank, int peer_index, int kind, int* peer_rank,
                      sgc_motion_item* items, int cap) {
  if (!dlo || !dhi || !maxbox) return fail(SGC_ERR_ARG, "sgc_halo_describe: null argument");
  Grid g(dlo, dhi, maxbox, nproc);
  std::vector<sgc_motion_item> mine;
  int st = build_exchange_items(g, nghost, periodic, trim, rank, mine);
  if (st) return st;
  std::vector<HaloList> lists;
  st = build_halo_lists(g, nghost, periodic, trim, rank, mine, lists);
  if (st) return st;
  if (peer_index < 0 || peer_index >= (int)lists.size()) return SGC_ERR_ARG;
  const HaloList& h = lists[peer_index];
  if (peer_rank) *peer_rank = h.peer;
  const std::vector<sgc_motion_item>& v = kind ? h.send : h.recv;
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) items[i] = v[i];
  return (int)v.size();
}

}  // extern "C"
