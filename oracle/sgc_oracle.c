/* TEST INFRASTRUCTURE — the ORACLE.  Not part of the product path.
 *
 * Plain-C restatement of the reference's algorithm for the hot path (stencil sweep over
 * each Box of a box-decomposed level + the ghost-cell exchange that precedes it).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.
 *
 * Parity is PINNED: tests/test_oracle_vs_reference.py checks every function below against
 * the real reference library compiled from /root/reference (oracle/_ref/libsgcref*.so, via
 * oracle/ref_shim.cpp) and tests/test_oracle_golden.py checks it against the committed
 * fixtures in tests/golden/ generated from that library (oracle/gen_golden.py), including
 * the survey's known-answer hash 978934eb8172d8e9.
 *
 * All file:line citations are relative to /root/reference/Structured_gridcalc/.
 * Compile with -ffp-contract=off: fused multiply-adds are written explicitly (fma())
 * where the reference's g++ build (-ffp-contract=fast by default) emits them, selected at
 * run time by `contract`, so both reference builds can be matched bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { int lo[3], hi[3]; } obox;

/* ---------------------------------------------------------------- index space (L1) */

/* Box::size, lib/src/BoxFramework/Box.H:532-538 (int product of extents) */
static long box_size(const obox* b)
{
  return (long)(b->hi[0] - b->lo[0] + 1)*(b->hi[1] - b->lo[1] + 1)*(b->hi[2] - b->lo[2] + 1);
}
/* Box::isEmpty, Box.H:560-566 */
__attribute__((unused)) static int box_empty(const obox* b)
{
  return b->hi[0] < b->lo[0] || b->hi[1] < b->lo[1] || b->hi[2] < b->lo[2];
}
/* Box::grow(int), Box.H:355-361 */
static void box_grow(obox* b, int n)
{
  for (int d = 0; d < 3; ++d) { b->lo[d] -= n; b->hi[d] += n; }
}
/* Box::operator&=, Box.H:468-474 (max of lows, min of highs; may yield an empty box) */
static void box_isect(obox* b, const obox* o)
{
  for (int d = 0; d < 3; ++d)
    {
      if (o->lo[d] > b->lo[d]) b->lo[d] = o->lo[d];
      if (o->hi[d] < b->hi[d]) b->hi[d] = o->hi[d];
    }
}
/* Box::shift(IntVect), Box.H:407-413 */
static void box_shift(obox* b, const int* s)
{
  for (int d = 0; d < 3; ++d) { b->lo[d] += s[d]; b->hi[d] += s[d]; }
}
/* Box::adjBox, Box.H:435-460: four sign cases of (ncell, side) */
void sgco_adjbox(const int* lo, const int* hi, int ncell, int dir, int side, int* out6)
{
  obox b;
  memcpy(b.lo, lo, 12); memcpy(b.hi, hi, 12);
  if (ncell > 0 && side > 0)       { b.hi[dir] += ncell; b.lo[dir] = b.hi[dir] - ncell + 1; }
  else if (ncell > 0 && side <= 0) { b.lo[dir] -= ncell; b.hi[dir] = b.lo[dir] + ncell - 1; }
  else if (ncell < 0 && side > 0)  { b.lo[dir] = b.hi[dir] + ncell + 1; }
  else if (ncell < 0 && side <= 0) { b.hi[dir] = b.lo[dir] - ncell - 1; }
  memcpy(out6, b.lo, 12); memcpy(out6 + 3, b.hi, 12);
}

/* ------------------------------------------------------------------- layout (L3) */

/* DisjointBoxLayout::define, DisjointBoxLayout.cpp:93-168.  numBox = domain/maxBox (must
 * divide, :104-106); boxes stored x-fastest (:112-115,141-164); rank k owns linear indices
 * [k*bpp,(k+1)*bpp) (:121-139), must divide (:125).  The reference's walk misbehaves for
 * maxBox == 1; box extents >= 2 are required here.  Returns nbox or <0.      */
int sgco_layout(const int* dlo, const int* dhi, const int* maxbox, int nproc,
                int* boxes6, int* owner, int cap, int* numbox3)
{
  int nb[3];
  for (int d = 0; d < 3; ++d)
    {
      const int ext = dhi[d] - dlo[d] + 1;
      nb[d] = ext/maxbox[d];
      if (nb[d]*maxbox[d] != ext || maxbox[d] < 2) return -2;
    }
  const int n = nb[0]*nb[1]*nb[2];
  if (numbox3) { numbox3[0] = nb[0]; numbox3[1] = nb[1]; numbox3[2] = nb[2]; }
  if (nproc < 1 || n % nproc) return -3;
  if (n > cap) return -1;
  const int bpp = n/nproc;
  int idx = 0;
  for (int bz = 0; bz < nb[2]; ++bz)
    for (int by = 0; by < nb[1]; ++by)
      for (int bx = 0; bx < nb[0]; ++bx, ++idx)
        {
          const int b[3] = { bx, by, bz };
          for (int d = 0; d < 3; ++d)
            {
              boxes6[6*idx + d]     = dlo[d] + b[d]*maxbox[d];
              boxes6[6*idx + 3 + d] = dlo[d] + (b[d] + 1)*maxbox[d] - 1;
            }
          owner[idx] = idx/bpp;
        }
  return n;
}

/* ------------------------------------------------------------ exchange plan (L3/L4) */

#define SGCO_ITEM_INTS 30
/* One record per Motion2Way (Copier.H:135-160): [0] recv(local) box global idx, [1] send
 * (remote) box global idx, [2..7] regionRecv, [8..13] regionSendRemote (what
 * Motion2Way::regionSend() returns, Copier.H:111), [14..16] sendDir, [17] local proc,
 * [18] remote proc, [19] tagSend, [20] tagRecv (uniqueTag, Copier.H:388-395),
 * [21] recv local idx, [22] send "local" idx (global - localIdxBegin, LayoutIterator.H:386-391),
 * [23] compRecvFlags (all ones), [24..29] m_regionSend = local ∩ grow(remote,ng)
 * (MPI builds only in the reference, Copier.H:581-583,620-622).                          */
static void emit_item(int* r, int gl, int gr, const obox* recv, const obox* sendRemote,
                      const obox* sendLocal, const int* dir, int pl, int pr, int lbeg)
{
  r[0] = gl; r[1] = gr;
  memcpy(r + 2, recv->lo, 12);  memcpy(r + 5, recv->hi, 12);
  memcpy(r + 8, sendRemote->lo, 12); memcpy(r + 11, sendRemote->hi, 12);
  memcpy(r + 14, dir, 12);
  r[17] = pl; r[18] = pr;
  r[19] = 27*gl + (dir[0] + 1) + 3*(dir[1] + 1) + 9*(dir[2] + 1);
  r[20] = 27*gr + (-dir[0] + 1) + 3*(-dir[1] + 1) + 9*(-dir[2] + 1);
  r[21] = gl - lbeg; r[22] = gr - lbeg;
  r[23] = -1;
  memcpy(r + 24, sendLocal->lo, 12); memcpy(r + 27, sendLocal->hi, 12);
}

/* Copier::defineExchangeDBL, Copier.H:521-659, with NeighborIterator
 * (LayoutIterator.H:522-576: offsets over [-1,1]^3 ∩ box-grid, x fastest, skipping those
 * whose (1<<norm1) is in trim|TrimCenter) and PeriodicIterator (LayoutIterator.H:637-783:
 * offsets over [-1,1]^3 ∩ (box-grid grown by 1 in periodic dirs), outside the grid, wrapped
 * by ±numBox).  Items are emitted per local box: interior neighbours first, then periodic
 * images (Copier.H:574-641).  Returns the item count (records beyond cap are not written). */
int sgco_copier_items(const int* dlo, const int* dhi, const int* maxbox, int ng,
                      unsigned periodic, unsigned trim, int nproc, int procid,
                      int* items, int cap)
{
  int nb[3];
  for (int d = 0; d < 3; ++d)
    {
      const int ext = dhi[d] - dlo[d] + 1;
      nb[d] = ext/maxbox[d];
      if (nb[d]*maxbox[d] != ext) return -2;
    }
  const int n = nb[0]*nb[1]*nb[2];
  if (n % nproc) return -3;
  if (ng <= 0) return 0;                                     /* Copier.H:545 */
  const int bpp = n/nproc, lbeg = procid*bpp;
  const unsigned tr = trim | 1u;                             /* TrimCenter always */
  int cnt = 0;
  for (int gl = lbeg; gl < lbeg + bpp; ++gl)                 /* DataIterator */
    {
      const int base[3] = { gl % nb[0], (gl/nb[0]) % nb[1], gl/(nb[0]*nb[1]) };
      obox local;
      for (int d = 0; d < 3; ++d)
        {
          local.lo[d] = dlo[d] + base[d]*maxbox[d];
          local.hi[d] = local.lo[d] + maxbox[d] - 1;
        }
      for (int pass = 0; pass < 2; ++pass)                   /* 0 interior, 1 periodic */
        {
          if (pass == 1)
            {
              /* Copier.H:603: only boxes touching a periodic face */
              int touches = 0;
              for (int d = 0; d < 3; ++d)
                if ((periodic & (1u << d)) && (base[d] == 0 || base[d] == nb[d] - 1)) touches = 1;
              if (!touches) break;
            }
          for (int oz = -1; oz <= 1; ++oz)
            for (int oy = -1; oy <= 1; ++oy)
              for (int ox = -1; ox <= 1; ++ox)
                {
                  const int o[3] = { ox, oy, oz };
                  const int norm1 = abs(ox) + abs(oy) + abs(oz);
                  if ((1u << norm1) & tr) continue;
                  int outside = 0, legal = 1, wrap[3];
                  for (int d = 0; d < 3; ++d)
                    {
                      const int p = base[d] + o[d];
                      wrap[d] = p;
                      if (p < 0 || p >= nb[d])
                        {
                          outside = 1;
                          if (!(periodic & (1u << d))) legal = 0;
                          wrap[d] = p < 0 ? p + nb[d] : p - nb[d];
                        }
                    }
                  if (!legal || outside != pass) continue;
                  const int gr = wrap[0] + nb[0]*(wrap[1] + nb[1]*wrap[2]);
                  obox remote;
                  for (int d = 0; d < 3; ++d)
                    {
                      remote.lo[d] = dlo[d] + wrap[d]*maxbox[d];
                      remote.hi[d] = remote.lo[d] + maxbox[d] - 1;
                    }
                  int shiftBy[3] = { 0, 0, 0 }, neg[3];
                  if (pass == 1)                              /* Copier.H:611-615 */
                    for (int d = 0; d < 3; ++d)
                      shiftBy[d] = local.lo[d] - remote.lo[d] + o[d]*maxbox[d];
                  box_shift(&remote, shiftBy);
                  obox recv = local;
                  box_grow(&recv, ng);
                  box_isect(&recv, &remote);                  /* :576-579 / :616-618 */
                  obox sendLocal = local, rg = remote;
                  box_grow(&rg, ng);
                  box_isect(&sendLocal, &rg);                 /* :581-583 / :620-622 */
                  obox sendRemote = recv;
                  for (int d = 0; d < 3; ++d) neg[d] = -shiftBy[d];
                  box_shift(&sendRemote, neg);                /* :626-627 */
                  if (cnt < cap)
                    emit_item(items + (long)SGCO_ITEM_INTS*cnt, gl, gr, &recv, &sendRemote,
                              &sendLocal, o, gl/bpp, gr/bpp, lbeg);
                  ++cnt;
                }
        }
    }
  return cnt;
}

/* ------------------------------------------------------------- box array data (L2) */

/* BaseFab storage order, BaseFab.H:277-313 / BaseFab.cpp:490-503:
 * data[c*S + (i-lo0) + (j-lo1)*n0 + (k-lo2)*n0*n1], S = n0*n1*n2.                         */
static long fab_index(const obox* fb, int i, int j, int k, int c)
{
  const long n0 = fb->hi[0] - fb->lo[0] + 1, n1 = fb->hi[1] - fb->lo[1] + 1;
  return c*box_size(fb) + (i - fb->lo[0]) + n0*((j - fb->lo[1]) + n1*(long)(k - fb->lo[2]));
}

static obox mk(const int* lo, const int* hi)
{
  obox b; memcpy(b.lo, lo, 12); memcpy(b.hi, hi, 12); return b;
}

/* BaseFab::setVal(val) / setVal(comp,val), BaseFab.cpp:219-246 */
void sgco_fab_setval(const int* flo, const int* fhi, int ncomp, double* data, int comp, double val)
{
  const obox fb = mk(flo, fhi);
  const long S = box_size(&fb);
  if (comp < 0) for (long n = 0; n < S*ncomp; ++n) data[n] = val;
  else for (long n = 0; n < S; ++n) data[comp*S + n] = val;
}

/* BaseFab::copy(dstBox,dstComp,src,srcBox,srcComp,numComp,flags), BaseFab.cpp:296-329:
 * dst[i,c'] = src[i+off,c], off = srcBox.lo - dstBox.lo; component c' skipped unless
 * c' >= 32 or bit c' (absolute destination index, :320) of flags is set.                  */
void sgco_fab_copy(const int* dflo, const int* dfhi, int dncomp, double* ddata,
                   const int* sflo, const int* sfhi, int sncomp, const double* sdata,
                   const int* dlo, const int* dhi, int dcomp,
                   const int* slo, const int* shi, int scomp, int ncomp, unsigned flags)
{
  (void)dncomp; (void)sncomp; (void)shi;
  const obox dfb = mk(dflo, dfhi), sfb = mk(sflo, sfhi);
  const int off[3] = { slo[0] - dlo[0], slo[1] - dlo[1], slo[2] - dlo[2] };
  for (int ic = 0; ic < ncomp; ++ic)
    {
      const int dc = ic + dcomp, sc = ic + scomp;
      if (!(dc >= 32 || (flags & (1u << dc)))) continue;
      for (int k = dlo[2]; k <= dhi[2]; ++k)
        for (int j = dlo[1]; j <= dhi[1]; ++j)
          for (int i = dlo[0]; i <= dhi[0]; ++i)
            ddata[fab_index(&dfb, i, j, k, dc)] =
              sdata[fab_index(&sfb, i + off[0], j + off[1], k + off[2], sc)];
    }
}

/* BaseFab::copy(box,src), BaseFab.cpp:254-271: same region, all components */
void sgco_fab_copy_same(const int* dflo, const int* dfhi, int ncomp, double* ddata,
                        const int* sflo, const int* sfhi, const double* sdata,
                        const int* rlo, const int* rhi)
{
  sgco_fab_copy(dflo, dfhi, ncomp, ddata, sflo, sfhi, ncomp, sdata,
                rlo, rhi, 0, rlo, rhi, 0, ncomp, ~0u);
}

/* BaseFab::linearOut, BaseFab.cpp:350-375: pack order comp -> k -> j -> i, no header,
 * flag bit tested with the absolute component index (:367).  Returns elements written.    */
long sgco_fab_linear_out(const int* flo, const int* fhi, int ncomp, const double* data,
                         const int* rlo, const int* rhi, int c0, int c1, unsigned flags,
                         double* buffer)
{
  (void)ncomp;
  const obox fb = mk(flo, fhi);
  long p = 0;
  for (int ic = c0; ic < c1; ++ic)
    {
      if (!(ic >= 32 || (flags & (1u << ic)))) continue;
      for (int k = rlo[2]; k <= rhi[2]; ++k)
        for (int j = rlo[1]; j <= rhi[1]; ++j)
          for (int i = rlo[0]; i <= rhi[0]; ++i)
            buffer[p++] = data[fab_index(&fb, i, j, k, ic)];
    }
  return p;
}

/* BaseFab::linearIn, BaseFab.cpp:394-419 */
long sgco_fab_linear_in(const int* flo, const int* fhi, int ncomp, double* data,
                        const int* rlo, const int* rhi, int c0, int c1, unsigned flags,
                        const double* buffer)
{
  (void)ncomp;
  const obox fb = mk(flo, fhi);
  long p = 0;
  for (int ic = c0; ic < c1; ++ic)
    {
      if (!(ic >= 32 || (flags & (1u << ic)))) continue;
      for (int k = rlo[2]; k <= rhi[2]; ++k)
        for (int j = rlo[1]; j <= rhi[1]; ++j)
          for (int i = rlo[0]; i <= rhi[0]; ++i)
            data[fab_index(&fb, i, j, k, ic)] = buffer[p++];
    }
  return p;
}

/* -------------------------------------------------------------- level data (L4) */

/* Level storage across this interface: the fabs of the nbox boxes given in boxes6,
 * concatenated in that order; fab b lives on grow(box_b, ng) with ncomp components
 * (LevelData::define, LevelData.H:256-272).  fab_off[b] = element offset of fab b.        */
long sgco_level_offsets(const int* boxes6, int nbox, int ng, int ncomp, long* fab_off)
{
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      obox fb = mk(boxes6 + 6*b, boxes6 + 6*b + 3);
      box_grow(&fb, ng);
      if (fab_off) fab_off[b] = off;
      off += box_size(&fb)*ncomp;
    }
  return off;
}

/* LevelData::exchange, LevelData.H:455-566 (serial / all-local form, :471-486): for every
 * motion item, dstFab.copy(regionRecv, c0, srcFab, regionSend(), c0, nc, compRecvFlags).
 * boxes6 are the rank's local boxes (index = item[21]/item[22]).  Items whose send box is
 * not local (idx out of range) are skipped and counted in the return value.               */
int sgco_level_exchange(const int* boxes6, int nbox, int ng, int ncomp, int c0, int nc,
                        const int* items, int nitem, int item_ints, double* fabs)
{
  long* off = (long*)malloc(sizeof(long)*(size_t)nbox);
  sgco_level_offsets(boxes6, nbox, ng, ncomp, off);
  int skipped = 0;
  for (int m = 0; m < nitem; ++m)
    {
      const int* r = items + (long)item_ints*m;
      const int ld = r[21], ls = r[22];
      if (ls < 0 || ls >= nbox || ld < 0 || ld >= nbox) { ++skipped; continue; }
      obox dfb = mk(boxes6 + 6*ld, boxes6 + 6*ld + 3), sfb = mk(boxes6 + 6*ls, boxes6 + 6*ls + 3);
      box_grow(&dfb, ng); box_grow(&sfb, ng);
      sgco_fab_copy(dfb.lo, dfb.hi, ncomp, fabs + off[ld], sfb.lo, sfb.hi, ncomp, fabs + off[ls],
                    r + 2, r + 5, c0, r + 8, r + 11, c0, nc, (unsigned)r[23]);
    }
  free(off);
  return skipped;
}

/* ------------------------------------------------------------- stencil apply (L5) */

/* 7-point heat/Jacobi update over one box, the level form of the application loop
 * (laplacian.cpp:94-108 / LBPatch.H:10-24 pattern; SURVEY §8d association order):
 *   unew = u + f*( ((u[i+1]-2u)+u[i-1]) + ((u[j+1]-2u)+u[j-1]) + ((u[k+1]-2u)+u[k-1]) )
 * contract != 0: the final multiply-add is one fma, as g++ -ffp-contract=fast emits it
 * (2*u is exact, so the inner terms are contraction-invariant).                           */
void sgco_heat_box(const int* flo, const int* fhi, const double* u, double* unew,
                   const int* blo, const int* bhi, double f, int contract)
{
  const obox fb = mk(flo, fhi);
  const long sx = 1, sy = fb.hi[0] - fb.lo[0] + 1, sz = sy*(fb.hi[1] - fb.lo[1] + 1);
  for (int k = blo[2]; k <= bhi[2]; ++k)
    for (int j = blo[1]; j <= bhi[1]; ++j)
      for (int i = blo[0]; i <= bhi[0]; ++i)
        {
          const long c = fab_index(&fb, i, j, k, 0);
          const double a = u[c];
          const double S = ((u[c + sx] - 2*a) + u[c - sx]) +
                           ((u[c + sy] - 2*a) + u[c - sy]) +
                           ((u[c + sz] - 2*a) + u[c - sz]);
          unew[c] = contract ? fma(f, S, a) : a + f*S;
        }
}

void sgco_level_heat_sweep(const int* boxes6, int nbox, int ng, const double* u, double* unew,
                           double f, int contract)
{
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      obox fb = mk(boxes6 + 6*b, boxes6 + 6*b + 3);
      box_grow(&fb, ng);
      sgco_heat_box(fb.lo, fb.hi, u + off, unew + off, boxes6 + 6*b, boxes6 + 6*b + 3, f, contract);
      off += box_size(&fb);
    }
}

static double now_ms(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3*ts.tv_sec + 1e-6*ts.tv_nsec;
}

/* nstep x { exchange ; sweep }, ping-pong, as the level heat driver of SURVEY §8c/§8d.
 * fabs: u on entry (1 comp), last-written level on exit.  times_ms = {plan, exchange, sweep}. */
int sgco_level_heat(const int* dlo, const int* dhi, const int* maxbox, int ng, unsigned periodic,
                    unsigned trim, double f, int nstep, int contract, double* fabs, double* times_ms)
{
  int nb3[3];
  const int cap = ((dhi[0]-dlo[0]+1)/maxbox[0])*((dhi[1]-dlo[1]+1)/maxbox[1])*((dhi[2]-dlo[2]+1)/maxbox[2]);
  int* boxes = (int*)malloc(sizeof(int)*6*(size_t)cap);
  int* owner = (int*)malloc(sizeof(int)*(size_t)cap);
  const int nbox = sgco_layout(dlo, dhi, maxbox, 1, boxes, owner, cap, nb3);
  if (nbox < 0) { free(boxes); free(owner); return nbox; }
  double t0 = now_ms();
  const int nitem = sgco_copier_items(dlo, dhi, maxbox, ng, periodic, trim, 1, 0, 0, 0);
  int* items = (int*)malloc(sizeof(int)*SGCO_ITEM_INTS*(size_t)(nitem > 0 ? nitem : 1));
  sgco_copier_items(dlo, dhi, maxbox, ng, periodic, trim, 1, 0, items, nitem);
  double tplan = now_ms() - t0, tex = 0, tsw = 0;
  const long tot = sgco_level_offsets(boxes, nbox, ng, 1, 0);
  double* other = (double*)malloc(sizeof(double)*(size_t)tot);
  memcpy(other, fabs, sizeof(double)*(size_t)tot);
  double* lev[2] = { fabs, other };
  int cur = 0;
  for (int s = 0; s < nstep; ++s)
    {
      t0 = now_ms();
      sgco_level_exchange(boxes, nbox, ng, 1, 0, 1, items, nitem, SGCO_ITEM_INTS, lev[cur]);
      tex += now_ms() - t0;
      t0 = now_ms();
      sgco_level_heat_sweep(boxes, nbox, ng, lev[cur], lev[1 - cur], f, contract);
      tsw += now_ms() - t0;
      cur = 1 - cur;
    }
  if (cur == 1) memcpy(fabs, other, sizeof(double)*(size_t)tot);
  if (times_ms) { times_ms[0] = tplan; times_ms[1] = tex; times_ms[2] = tsw; }
  free(other); free(items); free(boxes); free(owner);
  return nitem;
}

/* The PR1 driver's sweep, application/laplacian/laplacian.cpp:90-110: A on (1..n)^3,
 * B on (2..n-1)^3 zeroed, niter x { for dir: B -= 0.5*((A[i+e] - 2A[i]) + A[i-e]) } as three
 * separate passes (:98-107).  0.5*T is exact, so contraction does not change the result.  */
void sgco_laplacian(int n, int niter, const double* A, double* B)
{
  const long pa = n, pb = n - 2;
  const long sa[3] = { 1, pa, pa*pa };
  memset(B, 0, sizeof(double)*(size_t)(pb*pb*pb));
  for (int it = 0; it < niter; ++it)
    for (int dir = 0; dir < 3; ++dir)
      for (int k = 2; k <= n - 1; ++k)
        for (int j = 2; j <= n - 1; ++j)
          for (int i = 2; i <= n - 1; ++i)
            {
              const long ia = (i - 1) + pa*((j - 1) + pa*(long)(k - 1));
              const long ib = (i - 2) + pb*((j - 2) + pb*(long)(k - 2));
              B[ib] -= 0.5*((A[ia + sa[dir]] - 2*A[ia]) + A[ia - sa[dir]]);
            }
}

/* Zero-gradient physical BC on a single box, WavePatch.cpp:156-166 (CPU) == kernelBC
 * (WavePatch_Cuda.cu:252-264): for dir x,y,z and side -,+: ghost face <- adjacent interior
 * face (faces only; srcBox = adjBox(domain,-1,dir,side), dstBox = srcBox shifted by side).  */
void sgco_bc_zero_gradient(const int* dlo, const int* dhi, int ng, double* u)
{
  obox fb = mk(dlo, dhi);
  box_grow(&fb, ng);
  for (int dir = 0; dir < 3; ++dir)
    for (int side = -1; side < 2; side += 2)
      {
        int s6[6], d6[6];
        sgco_adjbox(dlo, dhi, -1, dir, side, s6);
        memcpy(d6, s6, sizeof d6);
        d6[dir] += side; d6[3 + dir] += side;
        sgco_fab_copy(fb.lo, fb.hi, 1, u, fb.lo, fb.hi, 1, u, d6, d6 + 3, 0, s6, s6 + 3, 0, 1, ~0u);
      }
}

/* Leap-frog wave update, scalar form WavePatch.cpp:226-244:
 *   unp1 = 2*un - unm1 ; for dir: unp1 += factor*((un[+e] - 2un) + un[-e])
 * (three read-modify-write passes; with contraction each pass is one fma).                */
void sgco_wave_box(const int* flo, const int* fhi, const double* un, const double* unm1,
                   double* unp1, const int* blo, const int* bhi, double factor, int contract)
{
  const obox fb = mk(flo, fhi);
  const long st[3] = { 1, fb.hi[0] - fb.lo[0] + 1,
                       (long)(fb.hi[0] - fb.lo[0] + 1)*(fb.hi[1] - fb.lo[1] + 1) };
  for (int k = blo[2]; k <= bhi[2]; ++k)
    for (int j = blo[1]; j <= bhi[1]; ++j)
      for (int i = blo[0]; i <= bhi[0]; ++i)
        {
          const long c = fab_index(&fb, i, j, k, 0);
          double r = 2*un[c] - unm1[c];
          for (int d = 0; d < 3; ++d)
            {
              const double T = (un[c + st[d]] - 2*un[c]) + un[c - st[d]];
              r = contract ? fma(factor, T, r) : r + factor*T;
            }
          unp1[c] = r;
        }
}

/* Leap-frog update in the form of the shipped wave app's pencil path, WavePatch.cpp:194-206
 * (USE_VEX is defined at :31): unp1 = 2 un - unm1 + factor*MD_DIRSUM(dir -> un[+e] - 2 un + un[-e]).
 * MD_DIRSUM (BaseFabMacros.H:287-317) expands to f(3) + (f(2) + (f(1) + f(0))): f(3) has a zero
 * offset and is exactly 0, so the sum is Tz + (Ty + Tx); the final multiply-add is
 * one fma when contracted.                                                                    */
void sgco_wave_box_dirsum(const int* flo, const int* fhi, const double* un, const double* unm1,
                          double* unp1, const int* blo, const int* bhi, double factor, int contract)
{
  const obox fb = mk(flo, fhi);
  const long st[3] = { 1, fb.hi[0] - fb.lo[0] + 1,
                       (long)(fb.hi[0] - fb.lo[0] + 1)*(fb.hi[1] - fb.lo[1] + 1) };
  for (int k = blo[2]; k <= bhi[2]; ++k)
    for (int j = blo[1]; j <= bhi[1]; ++j)
      for (int i = blo[0]; i <= bhi[0]; ++i)
        {
          const long c = fab_index(&fb, i, j, k, 0);
          double T[3];
          for (int d = 0; d < 3; ++d) T[d] = (un[c + st[d]] - 2*un[c]) + un[c - st[d]];
          const double sum = T[2] + (T[1] + T[0]);
          const double t1 = 2*un[c] - unm1[c];
          unp1[c] = contract ? fma(factor, sum, t1) : t1 + factor*sum;
        }
}

/* nstep x { BC ; dirsum update ; rotate } on one box = WavePatch::advance as shipped (valid cells). */
void sgco_wave_dirsum(const int* dlo, const int* dhi, double factor, int nstep, int contract,
                      double* u0, double* u1)
{
  obox fb = mk(dlo, dhi);
  box_grow(&fb, 1);
  const size_t bytes = sizeof(double)*(size_t)box_size(&fb);
  double* buf[3] = { (double*)malloc(bytes), (double*)calloc(1, bytes), (double*)malloc(bytes) };
  memcpy(buf[0], u0, bytes); memcpy(buf[2], u1, bytes);
  int in = 0, inp1 = 1, inm1 = 2;
  for (int s = 0; s < nstep; ++s)
    {
      sgco_bc_zero_gradient(dlo, dhi, 1, buf[in]);
      sgco_wave_box_dirsum(fb.lo, fb.hi, buf[in], buf[inm1], buf[inp1], dlo, dhi, factor, contract);
      const int t = inm1; inm1 = in; in = inp1; inp1 = t;
    }
  memcpy(u0, buf[in], bytes); memcpy(u1, buf[inm1], bytes);
  free(buf[0]); free(buf[1]); free(buf[2]);
}

/* the same update over every box of a level (1 comp, ng ghosts, ghosts of un already filled) */
void sgco_level_wave_sweep(const int* boxes6, int nbox, int ng, const double* un, const double* unm1,
                           double* unp1, double factor, int contract)
{
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      obox fb = mk(boxes6 + 6*b, boxes6 + 6*b + 3);
      box_grow(&fb, ng);
      sgco_wave_box(fb.lo, fb.hi, un + off, unm1 + off, unp1 + off, boxes6 + 6*b, boxes6 + 6*b + 3, factor, contract);
      off += box_size(&fb);
    }
}

/* nstep x { BC ; update ; rotate } on one box (WavePatch::advance, WavePatch.cpp:136-254).
 * u0 = u^n, u1 = u^{n-1} on grow(domain,1); on exit u0 = latest, u1 = previous.           */
void sgco_wave(const int* dlo, const int* dhi, double factor, int nstep, int contract,
               double* u0, double* u1)
{
  obox fb = mk(dlo, dhi);
  box_grow(&fb, 1);
  const size_t bytes = sizeof(double)*(size_t)box_size(&fb);
  double* buf[3] = { (double*)malloc(bytes), (double*)calloc(1, bytes), (double*)malloc(bytes) };
  memcpy(buf[0], u0, bytes); memcpy(buf[2], u1, bytes);
  int in = 0, inp1 = 1, inm1 = 2;
  for (int s = 0; s < nstep; ++s)
    {
      sgco_bc_zero_gradient(dlo, dhi, 1, buf[in]);
      sgco_wave_box(fb.lo, fb.hi, buf[in], buf[inm1], buf[inp1], dlo, dhi, factor, contract);
      const int t = inm1; inm1 = in; in = inp1; inp1 = t;
    }
  memcpy(u0, buf[in], bytes); memcpy(u1, buf[inm1], bytes);
  free(buf[0]); free(buf[1]); free(buf[2]);
}

/* ------------------------------------------------------------- fixtures / hashing */

/* The survey's seedless dyadic init: values k/1024, k in [0,1024). */
double sgco_hash_init(int i, int j, int k)
{
  const unsigned h = ((unsigned)i*73856093u) ^ ((unsigned)j*19349663u) ^ ((unsigned)k*83492791u);
  return (double)(h % 1024u)/1024.0;
}

/* Fill valid cells of every fab with f(i,j,k); ghosts get ghostval. */
void sgco_level_fill_hash(const int* boxes6, int nbox, int ng, double ghostval, double* fabs)
{
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      const long S = box_size(&fb);
      for (long n = 0; n < S; ++n) fabs[off + n] = ghostval;
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i)
            fabs[off + fab_index(&fb, i, j, k, 0)] = sgco_hash_init(i, j, k);
      off += S;
    }
}

/* Gather the valid cells of comp `comp` into one global array over the domain, x fastest. */
void sgco_level_gather(const int* dlo, const int* dhi, const int* boxes6, int nbox, int ng,
                       int ncomp, int comp, const double* fabs, double* global)
{
  const obox dom = mk(dlo, dhi);
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i)
            global[fab_index(&dom, i, j, k, 0)] = fabs[off + fab_index(&fb, i, j, k, comp)];
      off += box_size(&fb)*ncomp;
    }
}

/* FNV-1a 64 over raw bytes. */
uint64_t sgco_fnv1a64(const void* p, long nbytes)
{
  const unsigned char* s = (const unsigned char*)p;
  uint64_t h = 1469598103934665603ull;
  for (long n = 0; n < nbytes; ++n) { h ^= s[n]; h *= 1099511628211ull; }
  return h;
}

/* ------------------------------------------------- lattice Boltzmann application (D3Q19) */

/* LBParameters.H:38-49,71-96,134-140: weights, lattice velocities, opposite directions.   */
#define SGCO_LB_NVEL 19
static const double lb_w[SGCO_LB_NVEL] = {
  1./3.,
  1./18., 1./18., 1./18., 1./18., 1./18., 1./18.,
  1./36., 1./36., 1./36., 1./36., 1./36., 1./36.,
  1./36., 1./36., 1./36., 1./36., 1./36., 1./36. };
static const int lb_e[SGCO_LB_NVEL][3] = {
  { 0, 0, 0 }, { -1, 0, 0 }, { 1, 0, 0 }, { 0, -1, 0 }, { 0, 1, 0 }, { 0, 0, -1 }, { 0, 0, 1 },
  { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 }, { 1, 1, 0 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 },
  { 1, 0, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 }, { 0, 1, 1 } };
static const int lb_opp[SGCO_LB_NVEL] = { 0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15 };
static const double lb_tau = 0.516, lb_cs2 = 1.0/3, lb_force = 0.000001042;   /* LBParameters.H:52-54 */

void sgco_lb_constants(double* w19, int* e57, int* opp19, double* tau_cs2_force)
{
  memcpy(w19, lb_w, sizeof lb_w); memcpy(e57, lb_e, sizeof lb_e); memcpy(opp19, lb_opp, sizeof lb_opp);
  tau_cs2_force[0] = lb_tau; tau_cs2_force[1] = lb_cs2; tau_cs2_force[2] = lb_force;
}

/* LBPatch::collision (LBPatch.H:27-46) with LBPhysics::collision (LBPhysics.H:5-14) on the valid
 * cells of every box, no contraction (the -ffp-contract=off build of the reference):
 *   eu   = (u0*e0 + u1*e1) + u2*e2
 *   feq  = (w*rho)*( ((1 + eu/cs2) + (eu*eu)/((2*cs2)*cs2)) - ((u0*u0 + u1*u1) + u2*u2)/(2*cs2) )
 *   f    = (f + (feq - f)/tau) + ((3*w)*e0)*force                                          */
void sgco_lb_collision(const int* boxes6, int nbox, int ng, double* fi, const double* macro)
{
  long offf = 0, offm = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      const long S = box_size(&fb);
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i)
            {
              const long c = fab_index(&fb, i, j, k, 0);
              const double rho = macro[offm + c];
              const double u0 = macro[offm + S + c], u1 = macro[offm + 2*S + c], u2 = macro[offm + 3*S + c];
              for (int q = 0; q < SGCO_LB_NVEL; ++q)
                {
                  const double eu = (u0*lb_e[q][0] + u1*lb_e[q][1]) + u2*lb_e[q][2];
                  const double feq = (lb_w[q]*rho)*(((1 + eu/lb_cs2) + (eu*eu)/(2*lb_cs2*lb_cs2))
                                                    - ((u0*u0 + u1*u1) + u2*u2)/(2*lb_cs2));
                  double* f = fi + offf + q*S + c;
                  *f = (*f + (feq - *f)/lb_tau) + 3*lb_w[q]*lb_e[q][0]*lb_force;
                }
            }
      offf += S*SGCO_LB_NVEL; offm += S*4;
    }
}

/* LBPatch::macroscopic (LBPatch.H:9-23) with LBPhysics::macroscopic (LBPhysics.H:17-35):
 * rho = sum f_q, u_d = (sum f_q*e_q[d])/rho, sums in the order q = 0..18 starting from 0.   */
void sgco_lb_macroscopic(const int* boxes6, int nbox, int ng, const double* fi, double* macro)
{
  long offf = 0, offm = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      const long S = box_size(&fb);
      for (int k = lo[2]; k <= hi[2]; ++k)
        for (int j = lo[1]; j <= hi[1]; ++j)
          for (int i = lo[0]; i <= hi[0]; ++i)
            {
              const long c = fab_index(&fb, i, j, k, 0);
              double m[4] = { 0, 0, 0, 0 };
              for (int q = 0; q < SGCO_LB_NVEL; ++q)
                {
                  const double f = fi[offf + q*S + c];
                  m[0] += f; m[1] += f*lb_e[q][0]; m[2] += f*lb_e[q][1]; m[3] += f*lb_e[q][2];
                }
              macro[offm + c] = m[0];
              for (int d = 1; d < 4; ++d) macro[offm + d*S + c] = m[d]/m[0];
            }
      offf += S*SGCO_LB_NVEL; offm += S*4;
    }
}

/* LBPatch::stream (LBPatch.H:49-64) without the swap: for every box and direction q,
 * prev(box, q) = fi(box shifted by e[opp(q)], q) — each valid cell pulls from the cell at -e_q.
 * Ghost cells of prev are not written.                                                      */
void sgco_lb_stream(const int* boxes6, int nbox, int ng, const double* fi, double* prev)
{
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      for (int q = 0; q < SGCO_LB_NVEL; ++q)
        {
          const int* sh = lb_e[lb_opp[q]];
          const int slo[3] = { lo[0] + sh[0], lo[1] + sh[1], lo[2] + sh[2] };
          const int shi[3] = { hi[0] + sh[0], hi[1] + sh[1], hi[2] + sh[2] };
          sgco_fab_copy(fb.lo, fb.hi, SGCO_LB_NVEL, prev + off, fb.lo, fb.hi, SGCO_LB_NVEL, fi + off,
                        lo, hi, q, slo, shi, q, 1, ~0u);
        }
      off += box_size(&fb)*SGCO_LB_NVEL;
    }
}

/* The bounce-back ghost fill of LBLevel::advance (LBLevel.cpp:65-112): a box whose top face is the
 * domain's top copies, for every cell c of its top valid layer, f_6(c) -> f_5(c+e6),
 * f_13 -> f_12(c+e13), f_14 -> f_11(c+e14), f_17 -> f_16(c+e17), f_18 -> f_15(c+e18); ELSE a box on
 * the domain's bottom does the mirror image (5->6, 12->13, 11->14, 16->17, 15->18 along e5, e12,
 * e11, e16, e15).  A box touching both only gets the top treatment (the reference's else-if).     */
void sgco_lb_bounce(const int* dlo, const int* dhi, const int* boxes6, int nbox, int ng, double* fi)
{
  static const int top[5][2] = { { 6, 5 }, { 13, 12 }, { 14, 11 }, { 17, 16 }, { 18, 15 } };   /* src, dst */
  (void)dlo;
  long off = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      int side = 0;
      if (hi[2] == dhi[2]) side = 1; else if (lo[2] == dlo[2]) side = -1;
      if (side)
        {
          const int z = side > 0 ? hi[2] : lo[2];
          const int slo[3] = { lo[0], lo[1], z }, shi[3] = { hi[0], hi[1], z };
          for (int m = 0; m < 5; ++m)
            {
              const int sc = side > 0 ? top[m][0] : top[m][1], dc = side > 0 ? top[m][1] : top[m][0];
              const int* e = lb_e[sc];
              const int tlo[3] = { slo[0] + e[0], slo[1] + e[1], slo[2] + e[2] };
              const int thi[3] = { shi[0] + e[0], shi[1] + e[1], shi[2] + e[2] };
              double* fab = fi + off;
              /* distinct components and disjoint cells: the in-place copy is order-independent */
              sgco_fab_copy(fb.lo, fb.hi, SGCO_LB_NVEL, fab, fb.lo, fb.hi, SGCO_LB_NVEL, fab,
                            tlo, thi, dc, slo, shi, sc, 1, ~0u);
            }
        }
      off += box_size(&fb)*SGCO_LB_NVEL;
    }
}

/* nstep x LBLevel::advance (LBLevel.cpp:7-127): collision; exchange(PeriodicX|PeriodicY,
 * TrimCorner) of all 19 components; bounce-back; stream + swap; macroscopic.  fi/prev/macro are
 * levels on the uniform layout (ng = 1); on return fi is the reference's m_curr.               */
int sgco_lb_advance(const int* dlo, const int* dhi, const int* maxbox, int nstep,
                    double* fi, double* prev, double* macro)
{
  int nb3[3];
  const int ng = 1;
  const int cap = ((dhi[0]-dlo[0]+1)/maxbox[0])*((dhi[1]-dlo[1]+1)/maxbox[1])*((dhi[2]-dlo[2]+1)/maxbox[2]);
  int* boxes = (int*)malloc(sizeof(int)*6*(size_t)cap);
  int* owner = (int*)malloc(sizeof(int)*(size_t)cap);
  const int nbox = sgco_layout(dlo, dhi, maxbox, 1, boxes, owner, cap, nb3);
  if (nbox < 0) { free(boxes); free(owner); return nbox; }
  const unsigned periodic = 3u, trim = 8u;       /* PeriodicX|PeriodicY, TrimCorner (LayoutIterator.H:27-41) */
  const int nitem = sgco_copier_items(dlo, dhi, maxbox, ng, periodic, trim, 1, 0, 0, 0);
  int* items = (int*)malloc(sizeof(int)*SGCO_ITEM_INTS*(size_t)(nitem > 0 ? nitem : 1));
  sgco_copier_items(dlo, dhi, maxbox, ng, periodic, trim, 1, 0, items, nitem);
  const long tot = sgco_level_offsets(boxes, nbox, ng, SGCO_LB_NVEL, 0);
  double* lev[2] = { fi, prev };
  int cur = 0;
  for (int s = 0; s < nstep; ++s)
    {
      sgco_lb_collision(boxes, nbox, ng, lev[cur], macro);
      sgco_level_exchange(boxes, nbox, ng, SGCO_LB_NVEL, 0, SGCO_LB_NVEL, items, nitem, SGCO_ITEM_INTS, lev[cur]);
      sgco_lb_bounce(dlo, dhi, boxes, nbox, ng, lev[cur]);
      sgco_lb_stream(boxes, nbox, ng, lev[cur], lev[1 - cur]);
      cur = 1 - cur;
      sgco_lb_macroscopic(boxes, nbox, ng, lev[cur], macro);
    }
  if (cur == 1)
    for (long n = 0; n < tot; ++n) { const double t = fi[n]; fi[n] = prev[n]; prev[n] = t; }
  free(items); free(boxes); free(owner);
  return nitem;
}

/* What the reference's plot writer hands to CGNS for a level (LevelData.cpp:150-193): for every
 * box, for every component, the VALID cells (ghosts stripped), x fastest.                        */
long sgco_level_valid_out(const int* boxes6, int nbox, int ng, int ncomp, const double* fabs, double* out)
{
  long off = 0, n = 0;
  for (int b = 0; b < nbox; ++b)
    {
      const int* lo = boxes6 + 6*b; const int* hi = lo + 3;
      obox fb = mk(lo, hi);
      box_grow(&fb, ng);
      for (int c = 0; c < ncomp; ++c)
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i)
              out[n++] = fabs[off + fab_index(&fb, i, j, k, c)];
      off += box_size(&fb)*ncomp;
    }
  return n;
}
