/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * ref_shim.cpp: a thin extern "C" driver written against the REFERENCE library's own
 * C++ API (Box / BaseFab / DisjointBoxLayout / LevelData / Copier and the MD_* loop
 * macros).  It is compiled by oracle/build_oracle.sh together with the reference's
 * *unmodified* sources where they lie under /root/reference (never copied here) into
 * oracle/_ref/libsgcref*.so.  Tests use it to pin oracle/sgc_oracle.c; bench.py uses
 * it as the CPU baseline ("kind": "reference").
 *
 * Reference call sites driven here:
 *   Copier::defineExchangeDBL      lib/src/BoxFramework/Copier.H:521-659
 *   LevelData::exchange / Begin    lib/src/BoxFramework/LevelData.H:455-566,577-618
 *   BaseFab::copy/linearOut/In     lib/src/BoxFramework/BaseFab.cpp:254-419
 *   stencil loop form              application/laplacian/laplacian.cpp:90-110,
 *                                  application/wave/WavePatch.cpp:226-244 (scalar form)
 *   DisjointBoxLayout::define      lib/src/BoxFramework/DisjointBoxLayout.cpp:93-168
 */
#include <cstring>
#include <cstdint>
#include <vector>
#include <cmath>

#include "Parameters.H"
#include "IntVect.H"
#include "Box.H"
#include "BoxIterator.H"
#include "BaseFab.H"
#include "BaseFabMacros.H"
#include "DisjointBoxLayout.H"
#include "LayoutIterator.H"
#include "Copier.H"
#include "LevelData.H"
#include "Stopwatch.H"

#ifdef _OPENMP
#include <omp.h>
#define SHIM_BOXLOOP MD_BOXLOOP_OMP
#else
#define SHIM_BOXLOOP MD_BOXLOOP
#endif

namespace {

/* The rank count / rank id are protected statics set only by initMPI(); a derived
 * class may set them so a serial build can *describe* an N-rank layout.            */
struct LayoutRanks : public DisjointBoxLayout
{
  static void set(int nproc, int id) { s_numProc = nproc; s_procID = id; }
};

inline Box mkbox(const int* lo, const int* hi)
{
  return Box(IntVect(lo[0], lo[1], lo[2]), IntVect(hi[0], hi[1], hi[2]));
}

inline void putbox(int* out, const Box& b)
{
  for (int d = 0; d != 3; ++d) { out[d] = b.loVect(d); out[3 + d] = b.hiVect(d); }
}

/* Concatenated level storage used across the shim boundary: fabs in local box order,
 * each ncomp*(box grown by ng).size() doubles, reference memory order.              */
void levelIn(LevelData<FArrayBox>& ld, const DisjointBoxLayout& dbl, const Real* src)
{
  for (DataIterator dit(dbl); dit.ok(); ++dit)
    {
      FArrayBox& fab = ld[dit];
      std::memcpy(fab.dataPtr(), src, sizeof(Real)*fab.size());
      src += fab.size();
    }
}

void levelOut(const LevelData<FArrayBox>& ld, const DisjointBoxLayout& dbl, Real* dst)
{
  for (DataIterator dit(dbl); dit.ok(); ++dit)
    {
      const FArrayBox& fab = ld[dit];
      std::memcpy(dst, fab.dataPtr(), sizeof(Real)*fab.size());
      dst += fab.size();
    }
}

/* One 7-point heat/Jacobi sweep over every box of the level, in the loop form the
 * reference applications use (MD_ARRAY_RESTRICT + MD_BOXLOOP over dbl[dit]).
 * unew = u + f*( (u[i+1]-2u+u[i-1]) + (u[j+1]-2u+u[j-1]) + (u[k+1]-2u+u[k-1]) )     */
void heatSweep(LevelData<FArrayBox>& unew, LevelData<FArrayBox>& u,
               const DisjointBoxLayout& dbl, const Real f)
{
  for (DataIterator dit(dbl); dit.ok(); ++dit)
    {
      const Box box = dbl[dit];
      FArrayBox& fa = u[dit];
      FArrayBox& fb = unew[dit];
      MD_ARRAY_RESTRICT(a, fa);
      MD_ARRAY_RESTRICT(b, fb);
      {
        SHIM_BOXLOOP(box, i)
          {
            b[MD_IX(i, 0)] = a[MD_IX(i, 0)] + f*(
              (a[0][i2][i1][i0+1] - 2*a[MD_IX(i, 0)] + a[0][i2][i1][i0-1]) +
              (a[0][i2][i1+1][i0] - 2*a[MD_IX(i, 0)] + a[0][i2][i1-1][i0]) +
              (a[0][i2+1][i1][i0] - 2*a[MD_IX(i, 0)] + a[0][i2-1][i1][i0]));
          }
      }
    }
}

}  // namespace

extern "C" {

/* bench.py forces the thread count: torchrun exports OMP_NUM_THREADS=1 to its children */
int sgcref_set_threads(int n)
{
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

int sgcref_openmp_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Layout: boxes6[nbox*6] (lo,hi), owner[nbox].  Returns nbox (or -1 if too small). */
int sgcref_layout(const int* dlo, const int* dhi, const int* maxbox, int nproc,
                  int* boxes6, int* owner, int cap)
{
  LayoutRanks::set(nproc, 0);
  DisjointBoxLayout dbl(mkbox(dlo, dhi), IntVect(maxbox[0], maxbox[1], maxbox[2]));
  const int n = dbl.size();
  if (n > cap) { LayoutRanks::set(1, 0); return -1; }
  for (int i = 0; i != n; ++i)
    {
      putbox(boxes6 + 6*i, dbl.getLinear(i).box);
      owner[i] = dbl.getLinear(i).proc;
    }
  LayoutRanks::set(1, 0);
  return n;
}

/* Exchange plan as seen from rank `procid` of `nproc`.  One record of 24 ints per
 * motion item: [0] recv box global idx, [1] send box global idx, [2..7] regionRecv,
 * [8..13] regionSend (the remote box's region, = Motion2Way::regionSend()),
 * [14..16] sendDir, [17] local proc, [18] remote proc, [19] tagSend, [20] tagRecv,
 * [21] recv local idx, [22] send local idx, [23] compRecvFlags.
 * Returns the number of items (items beyond `cap` are counted, not written).        */
int sgcref_copier_items(const int* dlo, const int* dhi, const int* maxbox, int ng,
                        int ncomp, unsigned periodic, unsigned trim,
                        int nproc, int procid, int* items24, int cap)
{
  LayoutRanks::set(nproc, procid);
  DisjointBoxLayout dbl(mkbox(dlo, dhi), IntVect(maxbox[0], maxbox[1], maxbox[2]));
  Copier cp;
  cp.defineExchangeDBL<Real>(dbl, ng, 0, ncomp, periodic, trim);
  const int n = cp.numMotionItem();
  for (int m = 0; m < n && m < cap; ++m)
    {
      const Motion2Way& it = cp[m];
      int* r = items24 + 24*m;
      r[0] = it.bidxRecv().globalIndex();
      r[1] = it.bidxSend().globalIndex();
      putbox(r + 2, it.regionRecv());
      putbox(r + 8, it.regionSend());
      for (int d = 0; d != 3; ++d) r[14 + d] = it.sendDir()[d];
      r[17] = dbl.proc(it.bidxRecv());
      r[18] = dbl.proc(it.bidxSend());
      r[19] = it.uniqueTag(it.bidxRecv(), it.sendDir());
      r[20] = it.uniqueTag(it.bidxSend(), -it.sendDir());
      r[21] = it.bidxRecv().localIndex();
      r[22] = it.bidxSend().localIndex();
      r[23] = (int)it.compRecvFlags();
    }
  LayoutRanks::set(1, 0);
  return n;
}

/* Serial exchange on a whole level (data in/out through `fabs`).
 * mode 0 = LevelData::exchange, 1 = exchangeBegin + exchangeEnd.                    */
int sgcref_level_exchange(const int* dlo, const int* dhi, const int* maxbox, int ng,
                          int ncomp, unsigned periodic, unsigned trim, int mode,
                          Real* fabs)
{
  LayoutRanks::set(1, 0);
  DisjointBoxLayout dbl(mkbox(dlo, dhi), IntVect(maxbox[0], maxbox[1], maxbox[2]));
  LevelData<FArrayBox> ld(dbl, ncomp, ng);
  levelIn(ld, dbl, fabs);
  Copier cp;
  cp.defineExchangeLD(ld, periodic, trim);
  if (mode == 0) ld.exchange(cp);
  else { ld.exchangeBegin(cp); ld.exchangeEnd(cp); }
  levelOut(ld, dbl, fabs);
  return cp.numMotionItem();
}

/* nstep x { exchange ; heat sweep } with ping-pong; `fabs` holds u (1 comp, ng ghosts)
 * on entry and the last-written level on exit.  times_ms[0..2] = plan build,
 * total exchange, total sweep (Stopwatch, ms).                                      */
int sgcref_level_heat(const int* dlo, const int* dhi, const int* maxbox, int ng,
                      unsigned periodic, unsigned trim, Real f, int nstep,
                      Real* fabs, double* times_ms)
{
  LayoutRanks::set(1, 0);
  DisjointBoxLayout dbl(mkbox(dlo, dhi), IntVect(maxbox[0], maxbox[1], maxbox[2]));
  LevelData<FArrayBox> lev[2] = { LevelData<FArrayBox>(dbl, 1, ng),
                                  LevelData<FArrayBox>(dbl, 1, ng) };
  levelIn(lev[0], dbl, fabs);
  levelIn(lev[1], dbl, fabs);   // ghosts of both copies start identical (BC values)
  Stopwatch<> tPlan, tEx, tSw;
  tPlan.start();
  Copier cp;
  cp.defineExchangeLD(lev[0], periodic, trim);
  tPlan.stop();
  int cur = 0;
  for (int s = 0; s != nstep; ++s)
    {
      tEx.start();
      lev[cur].exchange(cp);
      tEx.stop();
      tSw.start();
      heatSweep(lev[1 - cur], lev[cur], dbl, f);
      tSw.stop();
      cur = 1 - cur;
    }
  levelOut(lev[cur], dbl, fabs);
  if (times_ms)
    {
      times_ms[0] = tPlan.time();
      times_ms[1] = tEx.time();
      times_ms[2] = tSw.time();
    }
  return cp.numMotionItem();
}

/* The PR1 driver's VLA-macro sweep (laplacian.cpp:90-110) on caller-supplied A.
 * A on (1..n)^3, B on (2..n-1)^3, B zeroed first, `niter` x 3 directional passes.   */
double sgcref_laplacian(int n, int niter, const Real* A, Real* B)
{
  Box allocBox(IntVect::Unit, n*IntVect::Unit);
  FArrayBox fabA(allocBox, 1);
  std::memcpy(fabA.dataPtr(), A, sizeof(Real)*fabA.size());
  Box computeBox(allocBox);
  computeBox.grow(-1);
  FArrayBox fabB(computeBox, 1);
  Stopwatch<> timer;
  timer.start();
  fabB.setVal(0.);
  MD_ARRAY_RESTRICT(arrA, fabA);
  MD_ARRAY_RESTRICT(arrB, fabB);
  for (int iter = 0; iter != niter; ++iter)
    for (int dir = 0; dir != g_SpaceDim; ++dir)
      {
        const int MD_ID(o, dir);
        MD_BOXLOOP(computeBox, i)
          {
            arrB[MD_IX(i, 0)] -= 0.5*(arrA[MD_OFFSETIX(i,+,o, 0)] -
                                      2*arrA[MD_IX(i, 0)] +
                                      arrA[MD_OFFSETIX(i,-,o, 0)]);
          }
      }
  timer.stop();
  std::memcpy(B, fabB.dataPtr(), sizeof(Real)*fabB.size());
  return timer.time();
}

/* Same sweep through BoxIterator + operator() (laplacian.cpp:58-79). */
double sgcref_laplacian_boxiter(int n, int niter, const Real* A, Real* B)
{
  Box allocBox(IntVect::Unit, n*IntVect::Unit);
  FArrayBox fabA(allocBox, 1);
  std::memcpy(fabA.dataPtr(), A, sizeof(Real)*fabA.size());
  Box computeBox(allocBox);
  computeBox.grow(-1);
  FArrayBox fabB(computeBox, 1);
  Stopwatch<> timer;
  timer.start();
  fabB.setVal(0.);
  for (int iter = 0; iter != niter; ++iter)
    for (int dir = 0; dir != g_SpaceDim; ++dir)
      {
        const IntVect ivo((dir == 0), (dir == 1), (dir == 2));
        for (BoxIterator bit(computeBox); bit.ok(); ++bit)
          {
            const IntVect iv = *bit;
            fabB(iv, 0) -= 0.5*(fabA(iv+ivo, 0) - 2*fabA(iv, 0) + fabA(iv-ivo, 0));
          }
      }
  timer.stop();
  std::memcpy(B, fabB.dataPtr(), sizeof(Real)*fabB.size());
  return timer.time();
}

/* Single-box leap-frog wave step in the reference's scalar form
 * (WavePatch.cpp:156-166 BC by adjBox copies, :226-244 update), nstep steps.
 * u0 = u^n, u1 = u^{n-1}, both on grow(domain,1); on exit u0 = latest, u1 = previous. */
void sgcref_wave(const int* dlo, const int* dhi, Real factor, int nstep,
                 Real* u0, Real* u1)
{
  const Box domain = mkbox(dlo, dhi);
  Box gb(domain);
  gb.grow(1);
  FArrayBox u[3] = { FArrayBox(gb, 1), FArrayBox(gb, 1), FArrayBox(gb, 1) };
  std::memcpy(u[0].dataPtr(), u0, sizeof(Real)*u[0].size());
  std::memcpy(u[2].dataPtr(), u1, sizeof(Real)*u[2].size());
  u[1].setVal(0.);
  int in = 0, inp1 = 1, inm1 = 2;
  for (int s = 0; s != nstep; ++s)
    {
      for (int dir = 0; dir != g_SpaceDim; ++dir)
        for (int side = -1; side < 2; side += 2)
          {
            Box srcBox(domain);
            srcBox.adjBox(-1, dir, side);
            Box dstBox(srcBox);
            dstBox.shift(side, dir);
            u[in].copy(dstBox, 0, u[in], srcBox, 0, 1);
          }
      MD_ARRAY_RESTRICT(arrunp1, u[inp1]);
      MD_ARRAY_RESTRICT(arrun, u[in]);
      MD_ARRAY_RESTRICT(arrunm1, u[inm1]);
      {
        MD_BOXLOOP(domain, i)
          {
            arrunp1[MD_IX(i, 0)] = 2*arrun[MD_IX(i, 0)] - arrunm1[MD_IX(i, 0)];
          }
      }
      for (int dir = 0; dir != g_SpaceDim; ++dir)
        {
          const int MD_ID(o, dir);
          MD_BOXLOOP(domain, i)
            {
              arrunp1[MD_IX(i, 0)] += factor*(arrun[MD_OFFSETIX(i,+,o, 0)] -
                                              2*arrun[MD_IX(i, 0)] +
                                              arrun[MD_OFFSETIX(i,-,o, 0)]);
            }
        }
      const int t = inm1; inm1 = in; in = inp1; inp1 = t;
    }
  std::memcpy(u0, u[in].dataPtr(), sizeof(Real)*u[in].size());
  std::memcpy(u1, u[inm1].dataPtr(), sizeof(Real)*u[inm1].size());
}

/* BaseFab primitives on caller memory (alias construction, BaseFab.H:59-76). */
void sgcref_fab_copy(const int* dfb_lo, const int* dfb_hi, int dncomp, Real* ddata,
                     const int* sfb_lo, const int* sfb_hi, int sncomp, Real* sdata,
                     const int* dlo, const int* dhi, int dcomp,
                     const int* slo, const int* shi, int scomp, int ncomp,
                     unsigned flags)
{
  FArrayBox dst(mkbox(dfb_lo, dfb_hi), dncomp, ddata);
  FArrayBox src(mkbox(sfb_lo, sfb_hi), sncomp, sdata);
  dst.copy(mkbox(dlo, dhi), dcomp, src, mkbox(slo, shi), scomp, ncomp, flags);
}

void sgcref_fab_copy_same(const int* dfb_lo, const int* dfb_hi, int ncomp, Real* ddata,
                          const int* sfb_lo, const int* sfb_hi, Real* sdata,
                          const int* rlo, const int* rhi)
{
  FArrayBox dst(mkbox(dfb_lo, dfb_hi), ncomp, ddata);
  FArrayBox src(mkbox(sfb_lo, sfb_hi), ncomp, sdata);
  dst.copy(mkbox(rlo, rhi), src);
}

void sgcref_fab_linear_out(const int* fb_lo, const int* fb_hi, int ncomp, Real* data,
                           const int* rlo, const int* rhi, int c0, int c1,
                           unsigned flags, Real* buffer)
{
  FArrayBox fab(mkbox(fb_lo, fb_hi), ncomp, data);
  fab.linearOut(buffer, mkbox(rlo, rhi), c0, c1, flags);
}

void sgcref_fab_linear_in(const int* fb_lo, const int* fb_hi, int ncomp, Real* data,
                          const int* rlo, const int* rhi, int c0, int c1,
                          unsigned flags, const Real* buffer)
{
  FArrayBox fab(mkbox(fb_lo, fb_hi), ncomp, data);
  fab.linearIn(buffer, mkbox(rlo, rhi), c0, c1, flags);
}

void sgcref_fab_setval(const int* fb_lo, const int* fb_hi, int ncomp, Real* data,
                       int comp, Real val)
{
  FArrayBox fab(mkbox(fb_lo, fb_hi), ncomp, data);
  if (comp < 0) fab.setVal(val); else fab.setVal(comp, val);
}

/* Box::adjBox (Box.H:435-460) for the BC-region parity test. */
void sgcref_adjbox(const int* lo, const int* hi, int ncell, int dir, int side, int* out6)
{
  Box b = mkbox(lo, hi);
  b.adjBox(ncell, dir, side);
  putbox(out6, b);
}

}  // extern "C"
