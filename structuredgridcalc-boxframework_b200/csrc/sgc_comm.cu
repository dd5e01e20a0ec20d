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
#include <cstring>

#include "sgc_device.h"

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
  int st = build_halo_lists(g, l->ng, periodic, trim, p->rank, p->items, lists);
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
    pos = 0;
    for (const sgc_motion_item& it : h.send) {
      const IBox r = box_of(it.src_lo, it.src_hi);
      const int ls = it.send_box - l->gidx0;
      if (ls < 0 || ls >= l->nbox) return fail(SGC_ERR_COMM, "halo: peer asks for a box this rank does not own");
      make_items(pack, dense_geom(r, pos), r, 0, l->geom[ls], r, p->c0, p->nc, ~0u);
      pos += r.size() * p->nc;
    }
    peer.send_elems = pos;
    st = peer.pack.upload(pack);
    if (!st) st = peer.unpack.upload(unpack);
    if (st) return st;
    SGC_CUDA(cudaMalloc(&peer.d_send, (size_t)std::max<long long>(peer.send_elems, 1) * l->esz));
    SGC_CUDA(cudaMalloc(&peer.d_recv, (size_t)std::max<long long>(peer.recv_elems, 1) * l->esz));
    p->peers.push_back(peer);
  }
  SGC_CUDA(cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));
  SGC_CUDA(cudaEventCreateWithFlags(&p->ev_halo, cudaEventDisableTiming));
  return 0;
}

void comm_plan_release(sgc_plan_s* p) {
  for (sgc_plan_s::Peer& q : p->peers) {
    q.pack.release();
    q.unpack.release();
    if (q.d_send) cudaFree(q.d_send);
    if (q.d_recv) cudaFree(q.d_recv);
  }
  p->peers.clear();
  if (p->ev_ready) cudaEventDestroy(p->ev_ready);
  if (p->ev_halo) cudaEventDestroy(p->ev_halo);
  p->ev_ready = p->ev_halo = nullptr;
}

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
  SGC_NCCL(c.api.GroupStart());
  for (sgc_plan_s::Peer& q : p->peers) {
    if (q.send_elems) SGC_NCCL(c.api.Send(q.d_send, (size_t)q.send_elems * l->esz, ncclChar, q.rank, c.comm, r.comm));
    if (q.recv_elems) SGC_NCCL(c.api.Recv(q.d_recv, (size_t)q.recv_elems * l->esz, ncclChar, q.rank, c.comm, r.comm));
  }
  SGC_NCCL(c.api.GroupEnd());
  SGC_CUDA(cudaEventRecord(p->ev_halo, r.comm));
  return 0;
}

int comm_exchange_end(sgc_level_s* l, sgc_plan_s* p) {
  Runtime& r = rt();
  SGC_CUDA(cudaStreamWaitEvent(r.compute, p->ev_halo, 0));
  for (sgc_plan_s::Peer& q : p->peers) {
    const int st = launch_copy(q.unpack, l->d_arena, q.d_recv, l->esz, r.compute);
    if (st) return st;
  }
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
