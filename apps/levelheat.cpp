// levelheat — box-decomposed level driver: nstep x { LevelData::exchange ; 7-point heat sweep } on a
// periodic domain, the loop the reference's only multi-box application runs per step
// (application/latticeBoltzmann/LBLevel.cpp:12-14: Copier + exchange, then per-box sweeps).
// Written against the facade exactly as one would against the reference's headers; prints the
// reference-style timing lines and one JSON line.
//   usage: levelheat [domain [box [nstep]]]      default 256 64 100
// With RANK/WORLD_SIZE set (one process per GPU) the level is split into z-slabs.
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "BaseFab.H"
#include "BaseFabMacros.H"
#include "Copier.H"
#include "DisjointBoxLayout.H"
#include "LevelData.H"
#include "StencilOps.H"
#include "Stopwatch.H"

static unsigned long long fnv1a(const void* p, size_t n, unsigned long long h)
{
  const unsigned char* s = static_cast<const unsigned char*>(p);
  for (size_t i = 0; i != n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, const char* argv[])
{
  const int n = argc > 1 ? std::atoi(argv[1]) : 256;
  const int bs = argc > 2 ? std::atoi(argv[2]) : 64;
  const int nstep = argc > 3 ? std::atoi(argv[3]) : 100;
  DisjointBoxLayout::initMPI(argc, argv);
  const int np = DisjointBoxLayout::numProc();

  Box domain(IntVect::Zero, IntVect(n - 1, n - 1, n*np - 1));
  DisjointBoxLayout dbl(domain, bs*IntVect::Unit);
  LevelData<FArrayBox> u(dbl, 1, 1), v(dbl, 1, 1);

  // SURVEY §8c seedless dyadic initial condition, written on the host mirrors
  for (DataIterator dit(dbl); dit.ok(); ++dit)
    {
      FArrayBox& fab = u[dit];
      fab.setVal(0.);
      MD_ARRAY_RESTRICT(a, fab);
      MD_BOXLOOP(dbl[dit], i)
        {
          const unsigned h = ((unsigned)i0*73856093u) ^ ((unsigned)i1*19349663u) ^ ((unsigned)i2*83492791u);
          a[MD_IX(i, 0)] = (Real)(h % 1024u)/1024.0;
        }
      v[dit].copy(fab.box(), fab);
    }

  Stopwatch<> tPlan, tRun;
  tPlan.start();
  Copier copier;
  copier.defineExchangeLD(u, PeriodicX | PeriodicY | PeriodicZ);
  u.exchange(copier);                       // creates the device plan, warms up
  SGC_SAFE_CALL(sgc_sync());
  tPlan.stop();

  sgc_event_t e0, e1;
  SGC_SAFE_CALL(sgc_event_create(&e0));
  SGC_SAFE_CALL(sgc_event_create(&e1));
  tRun.start();
  SGC_SAFE_CALL(sgc_event_record(e0));
  LevelData<FArrayBox>& result = StencilOps::heatRun(u, v, copier, 0.125, nstep);
  SGC_SAFE_CALL(sgc_event_record(e1));
  SGC_SAFE_CALL(sgc_sync());
  tRun.stop();
  float ms = 0;
  SGC_SAFE_CALL(sgc_event_elapsed_ms(e0, e1, &ms));

  // hash of the valid cells of the local boxes, box by box, x fastest (rank-local)
  unsigned long long h = 1469598103934665603ull;
  for (DataIterator dit(dbl); dit.ok(); ++dit)
    {
      const FArrayBox& fab = result[dit];
      const Box& b = dbl[dit];
      for (int k = b.loVect(2); k <= b.hiVect(2); ++k)
        for (int j = b.loVect(1); j <= b.hiVect(1); ++j)
          h = fnv1a(&fab(IntVect(b.loVect(0), j, k), 0), sizeof(Real)*b.dimensions()[0], h);
    }
  if (DisjointBoxLayout::procID() == 0)
    {
      std::cout << std::left << std::setw(40) << "Motion items: " << copier.numMotionItem() << std::endl;
      std::cout << std::left << std::setw(40) << "Time for plan + first exchange (ms): " << tPlan.time() << std::endl;
      std::cout << std::left << std::setw(40) << "Time for step loop (ms): " << ms << "  (host wall " << tRun.time() << ")" << std::endl;
    }
  const double cells = (double)n*n*n*nstep;
  std::printf("{\"driver\": \"levelheat\", \"rank\": %d, \"nproc\": %d, \"domain\": %d, \"box\": %d, \"nstep\": %d, "
              "\"gcell_per_s_per_gpu\": %.3f, \"local_fnv1a64\": \"%016llx\"}\n",
              DisjointBoxLayout::procID(), np, n, bs, nstep, cells/(ms*1e-3)/1e9, h);
  sgc_event_destroy(e0);
  sgc_event_destroy(e1);
  DisjointBoxLayout::finalizeMPI();
  return 0;
}
