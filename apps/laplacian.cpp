// laplacian — the PR1 reference driver (application/laplacian/laplacian.cpp) on the B200 facade.
// Same set-up (A(i,j,k) = i+j+k on (1..n)^3, B on (2..n-1)^3, niter x 3 directional updates
// B -= 0.5*(A[+e]-2A+A[-e])), same checks (B == 0 everywhere) and the same printed lines, plus one
// JSON line with Gcell-updates/s and the fraction of the HBM roofline.  The sweep is the device
// operator StencilOps::laplacianApply instead of the hand-written MD_BOXLOOP nests.
//   usage: laplacian [n [niter]]        (reference: n = 64, niter = 16, both hard-coded)
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "BaseFab.H"
#include "BaseFabMacros.H"
#include "BoxIterator.H"
#include "StencilOps.H"
#include "Stopwatch.H"

int main(int argc, char* argv[])
{
  const int n = argc > 1 ? std::atoi(argv[1]) : 128;
  const int niter = argc > 2 ? std::atoi(argv[2]) : 16;
  Box allocBox(IntVect::Unit, n*IntVect::Unit);
  FArrayBox fabA(allocBox, 1);

  {
    MD_ARRAY_RESTRICT(arrA, fabA);
    MD_BOXLOOP(allocBox, i) { arrA[MD_IX(i, 0)] = i0 + i1 + i2; }
  }

  Stopwatch<> timer;
  std::cout << std::left << std::setw(40) << "Clock resolution (ns): " << timer.resolution() << std::endl;

  Box computeBox(allocBox);
  computeBox.grow(-1);
  FArrayBox fabB(computeBox, 1);

  // warm-up (also uploads A)
  fabB.setVal(0.);
  StencilOps::laplacianApply(fabB, fabA, 0.5, niter);
  SGC_SAFE_CALL(sgc_sync());

  int stat = 0;
  double ms[2] = {0, 0};
  const char* label[2] = {"Time for device sweep, per-iter (ms): ", "Time for device sweep, fused (ms): "};
  for (int fuse = 0; fuse != 2; ++fuse)
    {
      fabB.setVal(0.);
      SGC_SAFE_CALL(sgc_sync());
      sgc_event_t e0, e1;
      SGC_SAFE_CALL(sgc_event_create(&e0));
      SGC_SAFE_CALL(sgc_event_create(&e1));
      timer.reset();
      timer.start();
      SGC_SAFE_CALL(sgc_event_record(e0));
      StencilOps::laplacianApply(fabB, fabA, 0.5, niter, fuse != 0);
      SGC_SAFE_CALL(sgc_event_record(e1));
      SGC_SAFE_CALL(sgc_sync());
      timer.stop();
      float dev = 0;
      SGC_SAFE_CALL(sgc_event_elapsed_ms(e0, e1, &dev));
      ms[fuse] = dev;
      std::cout << std::left << std::setw(40) << label[fuse] << dev << "  (host wall " << timer.time() << ")" << std::endl;
      for (BoxIterator bit(computeBox); bit.ok(); ++bit)
        if (fabB(*bit, 0) != 0.) ++stat;
      CH_assert(stat == 0);
      sgc_event_destroy(e0);
      sgc_event_destroy(e1);
    }
  const double cells = (double)computeBox.size()*niter;
  std::printf("{\"driver\": \"laplacian\", \"n\": %d, \"niter\": %d, \"nonzero_cells\": %d, \"gcell_per_s_per_iter\": %.3f, "
              "\"gcell_per_s_fused\": %.3f, \"hbm_gbs_per_iter_24B\": %.1f}\n",
              n, niter, stat, cells/(ms[0]*1e-3)/1e9, cells/(ms[1]*1e-3)/1e9, cells*24/(ms[0]*1e-3)/1e9);
  return stat;
}
