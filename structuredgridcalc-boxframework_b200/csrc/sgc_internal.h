// sgc_internal.h — shared declarations of the library behind include/sgc_b200.h.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "sgc_b200.h"

namespace sgc {

struct IBox {
  int lo[3], hi[3];
  long long size() const {
    return (long long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
  }
  bool empty() const { return hi[0] < lo[0] || hi[1] < lo[1] || hi[2] < lo[2]; }
  int ext(int d) const { return hi[d] - lo[d] + 1; }
  IBox grown(int n) const {
    IBox b = *this;
    for (int d = 0; d < 3; ++d) { b.lo[d] -= n; b.hi[d] += n; }
    return b;
  }
  IBox shifted(const int s[3]) const {
    IBox b = *this;
    for (int d = 0; d < 3; ++d) { b.lo[d] += s[d]; b.hi[d] += s[d]; }
    return b;
  }
  IBox isect(const IBox& o) const {
    IBox b = *this;
    for (int d = 0; d < 3; ++d) {
      if (o.lo[d] > b.lo[d]) b.lo[d] = o.lo[d];
      if (o.hi[d] < b.hi[d]) b.hi[d] = o.hi[d];
    }
    return b;
  }
  bool contains(const IBox& o) const {
    for (int d = 0; d < 3; ++d)
      if (o.lo[d] < lo[d] || o.hi[d] > hi[d]) return false;
    return true;
  }
  static IBox from6(const int* p) {
    IBox b;
    for (int d = 0; d < 3; ++d) { b.lo[d] = p[d]; b.hi[d] = p[3 + d]; }
    return b;
  }
};

// Uniform box grid of a DisjointBoxLayout (DisjointBoxLayout.cpp:93-168).
struct Grid {
  IBox domain;
  int maxbox[3];
  int nb[3];
  int nbox;
  int nproc;
  int bpp;  // boxes per rank
  int status;  // 0 ok
  Grid(const int dlo[3], const int dhi[3], const int mb[3], int nproc_);
  IBox box(int g) const;
  int owner(int g) const { return g / bpp; }
};

// Host-side plan builder (Copier.H:521-659). Appends to `out` the items of rank `rank`.
int build_exchange_items(const Grid& g, int ng, unsigned periodic, unsigned trim, int rank,
                         std::vector<sgc_motion_item>& out);

// Which device-resident step loops a rank's exchange plan admits (host only).  Fills `nbr`, the table
// [nbox][27] of the LOCAL box in direction code 13 + dx + 3 dy + 9 dz of each local box (-1: none, or owned
// by another rank), and returns a mask of SGC_LOOP_* bits (sgc_b200.h).
int classify_step_loops(const std::vector<sgc_motion_item>& items, int nbox, int gidx0, std::vector<int>& nbr);

// Pillars of the two-step sweep (host only).  The local boxes form npil pillars of nbz boxes — box p + k * npil sits
// exactly on top of box p + (k-1) * npil, both are each other's LOCAL z neighbour in `nbr` (so the plan exchanges
// that face), and every box of a pillar has the same in-plane neighbours present / missing — or else nbz = 1 and
// npil = nbox (every box its own pillar).  `boxes` are the valid boxes in local order.
void pillar_structure(const std::vector<IBox>& boxes, const std::vector<int>& nbr, int& npil, int& nbz);

void set_error(const std::string& msg);
int fail(int status, const std::string& msg);

}  // namespace sgc
