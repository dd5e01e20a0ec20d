// sgc_plan.cpp — host-side index algebra: uniform layout and the ghost-exchange plan.
//
// Mirrors (clean-room, from the behaviour documented in SURVEY.md §3.1/§A and checked against
// the oracle): DisjointBoxLayout::define (lib/src/BoxFramework/DisjointBoxLayout.cpp:93-168),
// NeighborIterator / PeriodicIterator (LayoutIterator.H:522-783) and
// Copier::defineExchangeDBL (Copier.H:521-659).  No CUDA in this file: it runs (and is tested)
// on machines without a GPU.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "sgc_internal.h"

namespace sgc {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int status, const std::string& msg) {
  g_err = msg;
  return status;
}
const char* last_error_cstr() { return g_err.c_str(); }

Grid::Grid(const int dlo[3], const int dhi[3], const int mb[3], int nproc_) : nproc(nproc_), status(0) {
  for (int d = 0; d < 3; ++d) {
    domain.lo[d] = dlo[d];
    domain.hi[d] = dhi[d];
    maxbox[d] = mb[d];
    const int ext = dhi[d] - dlo[d] + 1;
    if (mb[d] < 1 || ext < 1) { status = SGC_ERR_ARG; nb[d] = 0; continue; }
    nb[d] = ext / mb[d];
    if (nb[d] * mb[d] != ext) status = SGC_ERR_ARG;  // boxes must tile the domain (:104-106)
  }
  nbox = nb[0] * nb[1] * nb[2];
  if (nproc < 1 || nbox < 1 || nbox % nproc) status = SGC_ERR_ARG;  // (:125)
  bpp = (status == 0) ? nbox / nproc : 1;
}

IBox Grid::box(int g) const {
  const int c[3] = {g % nb[0], (g / nb[0]) % nb[1], g / (nb[0] * nb[1])};
  IBox b;
  for (int d = 0; d < 3; ++d) {
    b.lo[d] = domain.lo[d] + c[d] * maxbox[d];
    b.hi[d] = b.lo[d] + maxbox[d] - 1;
  }
  return b;
}

// One neighbour direction of one box. `outside`: the offset leaves the box grid (periodic image).
struct Nbr {
  int dir[3];
  int cell[3];  // wrapped grid coordinate of the neighbour box
  bool outside;
};

// Enumerate the 26 directions in x-fastest order and classify them for box-grid cell `c`.
template <class F>
static void for_each_direction(const Grid& g, const int c[3], unsigned periodic, unsigned trim, F&& f) {
  const unsigned mask = trim | SGC_TRIM_CENTER;
  for (int code = 0; code < 27; ++code) {
    Nbr n;
    n.dir[0] = code % 3 - 1;
    n.dir[1] = (code / 3) % 3 - 1;
    n.dir[2] = code / 9 - 1;
    const int codim = std::abs(n.dir[0]) + std::abs(n.dir[1]) + std::abs(n.dir[2]);
    if (mask & (1u << codim)) continue;  // Trim{Center,Face,Edge,Corner} = 1<<norm1
    n.outside = false;
    bool reachable = true;
    for (int d = 0; d < 3; ++d) {
      int p = c[d] + n.dir[d];
      if (p < 0 || p >= g.nb[d]) {
        n.outside = true;
        if (!(periodic & (1u << d))) reachable = false;
        p += (p < 0) ? g.nb[d] : -g.nb[d];
      }
      n.cell[d] = p;
    }
    if (reachable) f(n);
  }
}

int build_exchange_items(const Grid& g, int ng, unsigned periodic, unsigned trim, int rank,
                         std::vector<sgc_motion_item>& out) {
  if (g.status) return fail(SGC_ERR_ARG, "layout: boxes must tile the domain and divide evenly among ranks");
  if (rank < 0 || rank >= g.nproc) return fail(SGC_ERR_ARG, "rank out of range");
  if (ng <= 0) return 0;
  const int first = rank * g.bpp;
  out.reserve(out.size() + (size_t)g.bpp * 26);         // at most 26 items per box: no regrowth (C3: 851 968 items, 100 MB)
  for (int gl = first; gl < first + g.bpp; ++gl) {
    const int c[3] = {gl % g.nb[0], (gl / g.nb[0]) % g.nb[1], gl / (g.nb[0] * g.nb[1])};
    const IBox local = g.box(gl);
    const IBox halo = local.grown(ng);
    // the reference emits all in-grid neighbours of a box before its periodic images
    for (int want_outside = 0; want_outside < 2; ++want_outside) {
      for_each_direction(g, c, periodic, trim, [&](const Nbr& n) {
        if ((int)n.outside != want_outside) return;
        const int gr = n.cell[0] + g.nb[0] * (n.cell[1] + g.nb[1] * n.cell[2]);
        const IBox remote = g.box(gr);
        int shift[3] = {0, 0, 0}, back[3] = {0, 0, 0};
        if (n.outside)
          for (int d = 0; d < 3; ++d) {
            shift[d] = local.lo[d] - remote.lo[d] + n.dir[d] * local.ext(d);
            back[d] = -shift[d];
          }
        const IBox image = remote.shifted(shift);       // where the neighbour appears to be
        const IBox recv = halo.isect(image);            // ghost cells it covers
        const IBox send = local.isect(image.grown(ng)); // what the neighbour needs from us
        const IBox src = recv.shifted(back);            // same cells in the neighbour's frame
        sgc_motion_item it;
        it.recv_box = gl;
        it.send_box = gr;
        for (int d = 0; d < 3; ++d) {
          it.recv_lo[d] = recv.lo[d]; it.recv_hi[d] = recv.hi[d];
          it.src_lo[d] = src.lo[d];   it.src_hi[d] = src.hi[d];
          it.send_lo[d] = send.lo[d]; it.send_hi[d] = send.hi[d];
          it.dir[d] = n.dir[d];
        }
        it.recv_proc = g.owner(gl);
        it.send_proc = g.owner(gr);
        it.tag_send = 27 * gl + (n.dir[0] + 1) + 3 * (n.dir[1] + 1) + 9 * (n.dir[2] + 1);
        it.tag_recv = 27 * gr + (1 - n.dir[0]) + 3 * (1 - n.dir[1]) + 9 * (1 - n.dir[2]);
        it.comp_flags = ~0u;
        out.push_back(it);
      });
    }
  }
  return 0;
}


// The fused one-step sweep (sgc_stream.cu) works from the neighbour table alone.  The two-step sweep
// (sgc_stream2.cu) needs more: LOCAL — every box has its 8 in-plane neighbours and both z neighbours on this
// rank; SLAB — the z neighbours may live on another rank (peer memory); BND — some FACES have no motion item
// at all (a physical boundary or a trimmed direction, whose ghost cells the exchange never writes), which the
// kernel's boundary instance handles.  What it cannot handle is an in-plane neighbour on another rank, or a
// diagonal box missing although both faces towards it exist.
int classify_step_loops(const std::vector<sgc_motion_item>& items, int nbox, int gidx0, std::vector<int>& nbr) {
  nbr.assign((size_t)nbox * 27, -1);
  std::vector<char> cover((size_t)nbox * 27, 0);      // (box, face direction) has a motion item, local or remote
  for (const sgc_motion_item& it : items) {
    const int D = it.recv_box - gidx0;
    if (abs(it.dir[0]) + abs(it.dir[1]) + abs(it.dir[2]) == 1) cover[(size_t)D * 27 + 13 + it.dir[0] + 3 * it.dir[1] + 9 * it.dir[2]] = 1;
    if (it.recv_proc != it.send_proc) continue;
    // the box S an item reads from sits, seen from the receiving box D, in the item's direction; seen from S,
    // D sits in the opposite one
    const int S = it.send_box - gidx0;
    nbr[(size_t)S * 27 + (1 - it.dir[0]) + 3 * (1 - it.dir[1]) + 9 * (1 - it.dir[2])] = D;
  }
  bool ok = true, all_inplane = true, all_z_local = true, all_z = true;
  for (int b = 0; b < nbox && ok; ++b) {
    const int* nb = nbr.data() + (size_t)b * 27;
    const char* cv = cover.data() + (size_t)b * 27;
    for (int d = 0; d < 2 && ok; ++d)
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        const int code = 13 + sgn * (d == 0 ? 1 : 3);
        if (nb[code] < 0) { all_inplane = false; if (cv[code]) ok = false; }     // covered but not local: another rank
      }
    for (int sy = -1; sy <= 1 && ok; sy += 2)
      for (int sx = -1; sx <= 1; sx += 2) {
        const int X = nb[13 + sx], Y = nb[13 + 3 * sy];
        if (X >= 0 && Y >= 0 && nbr[(size_t)X * 27 + 13 + 3 * sy] < 0) ok = false;   // diagonal box over two faces
      }
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const int code = 13 + 9 * sgn;
      if (nb[code] < 0) { all_z_local = false; if (!cv[code]) all_z = false; }
    }
  }
  int mask = SGC_LOOP_FUSED;
  if (ok && all_inplane && all_z_local) mask |= SGC_LOOP_TWO_STEP_LOCAL;
  if (ok && all_inplane && all_z) mask |= SGC_LOOP_TWO_STEP_SLAB;
  if (ok && !(all_inplane && all_z)) mask |= SGC_LOOP_TWO_STEP_BND;
  return mask;
}

void pillar_structure(const std::vector<IBox>& boxes, const std::vector<int>& nbr, int& npil, int& nbz) {
  const int nbox = (int)boxes.size();
  npil = nbox;
  nbz = 1;
  if (nbox < 2 || nbr.size() != (size_t)nbox * 27) return;
  // candidate: the boxes of the lowest layer come first (the layout's x-fastest order)
  int n = 0;
  while (n < nbox && boxes[n].lo[2] == boxes[0].lo[2]) ++n;
  if (n == nbox || nbox % n != 0) return;
  const int vz = boxes[0].ext(2);
  for (int b = n; b < nbox; ++b) {
    const IBox &up = boxes[b], &dn = boxes[b - n];
    if (up.lo[0] != dn.lo[0] || up.lo[1] != dn.lo[1] || up.hi[0] != dn.hi[0] || up.hi[1] != dn.hi[1]) return;
    if (up.lo[2] != dn.lo[2] + vz || up.ext(2) != vz || dn.ext(2) != vz) return;
    if (nbr[(size_t)b * 27 + 4] != b - n || nbr[(size_t)(b - n) * 27 + 22] != b) return;      // the shared face is exchanged, locally
    for (int code = 9; code < 18; ++code) {                                                     // the in-plane directions
      if (code == 13) continue;
      if ((nbr[(size_t)b * 27 + code] < 0) != (nbr[(size_t)(b - n) * 27 + code] < 0)) return;
    }
  }
  npil = n;
  nbz = nbox / n;
}

}  // namespace sgc

extern "C" {

const char* sgc_last_error(void) { return sgc::last_error_cstr(); }
int sgc_abi_version(void) { return SGC_ABI_VERSION; }

int sgc_layout_define(const int dlo[3], const int dhi[3], const int maxbox[3], int nproc, int* boxes6,
                      int* owner, int cap, int numbox[3]) {
  if (!dlo || !dhi || !maxbox) return sgc::fail(SGC_ERR_ARG, "sgc_layout_define: null argument");
  sgc::Grid g(dlo, dhi, maxbox, nproc);
  if (g.status)
    return sgc::fail(SGC_ERR_ARG, "sgc_layout_define: boxes must tile the domain and divide evenly among ranks");
  if (numbox)
    for (int d = 0; d < 3; ++d) numbox[d] = g.nb[d];
  for (int i = 0; i < g.nbox && i < cap; ++i) {
    const sgc::IBox b = g.box(i);
    if (boxes6)
      for (int d = 0; d < 3; ++d) { boxes6[6 * i + d] = b.lo[d]; boxes6[6 * i + 3 + d] = b.hi[d]; }
    if (owner) owner[i] = g.owner(i);
  }
  return g.nbox;
}

int sgc_copier_define(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost, unsigned periodic,
                      unsigned trim, int nproc, int rank, sgc_motion_item* items, int cap) {
  if (!dlo || !dhi || !maxbox) return sgc::fail(SGC_ERR_ARG, "sgc_copier_define: null argument");
  sgc::Grid g(dlo, dhi, maxbox, nproc);
  std::vector<sgc_motion_item> v;
  const int st = sgc::build_exchange_items(g, nghost, periodic, trim, rank, v);
  if (st) return st;
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) items[i] = v[i];
  return (int)v.size();
}

int sgc_step_loop_class(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost, unsigned periodic,
                        unsigned trim, int nproc, int rank) {
  if (!dlo || !dhi || !maxbox) return sgc::fail(SGC_ERR_ARG, "sgc_step_loop_class: null argument");
  sgc::Grid g(dlo, dhi, maxbox, nproc);
  if (g.status) return sgc::fail(SGC_ERR_ARG, "sgc_step_loop_class: boxes must tile the domain and divide among ranks");
  std::vector<sgc_motion_item> v;
  const int st = sgc::build_exchange_items(g, nghost, periodic, trim, rank, v);
  if (st) return st;
  std::vector<int> nbr;
  return sgc::classify_step_loops(v, g.bpp, rank * g.bpp, nbr);
}

int sgc_step_loop_pillars(const int dlo[3], const int dhi[3], const int maxbox[3], int nghost, unsigned periodic,
                          unsigned trim, int nproc, int rank, int pillars[2]) {
  if (!dlo || !dhi || !maxbox || !pillars) return sgc::fail(SGC_ERR_ARG, "sgc_step_loop_pillars: null argument");
  sgc::Grid g(dlo, dhi, maxbox, nproc);
  if (g.status) return sgc::fail(SGC_ERR_ARG, "sgc_step_loop_pillars: boxes must tile the domain and divide among ranks");
  std::vector<sgc_motion_item> v;
  const int st = sgc::build_exchange_items(g, nghost, periodic, trim, rank, v);
  if (st) return st;
  std::vector<int> nbr;
  const int mask = sgc::classify_step_loops(v, g.bpp, rank * g.bpp, nbr);
  std::vector<sgc::IBox> boxes(g.bpp);
  for (int b = 0; b < g.bpp; ++b) boxes[b] = g.box(rank * g.bpp + b);
  sgc::pillar_structure(boxes, nbr, pillars[0], pillars[1]);
  return mask;
}

}  // extern "C"
