// wave — the reference's second shipped driver (application/wave/wave.cpp + WavePatch) on the B200
// facade: 3-D scalar wave equation, leap-frog, single box = whole domain h^3, three time levels as
// three LevelData with one ghost layer, reflective (zero-gradient) BC, c = 1, dx = 1/h, cfl = 0.125.
// The per-step work of WavePatch::advance (WavePatch.cpp:136-254) — BC face copies, then the update —
// runs as two device launches per step inside a cudaEvent bracket (the shape of
// driverAdvanceIterGroup, WavePatch_Cuda.cu:352-378); SGC_WAVE_DIRSUM reproduces the rounding of
// the shipped AVX pencil path bit for bit.
//   usage: wave [h [iters]]          (reference: ./wave [-np x] [h [iters]], defaults 32 100)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "BaseFab.H"
#include "BaseFabMacros.H"
#include "DisjointBoxLayout.H"
#include "LevelData.H"
#include "StencilOps.H"
#include "Stopwatch.H"

int main(int argc, char* argv[])
{
  const int h = argc > 1 ? std::atoi(argv[1]) : 32;
  const int numIter = argc > 2 ? std::atoi(argv[2]) : 100;
  const Real dx = 1./h, c = 1., cfl = 0.125;
  const Real dt = dx*cfl/c;
  const Real factor = std::pow(dt*c/dx, 2)/g_SpaceDim;

  Box domain(IntVect::Zero, (h - 1)*IntVect::Unit);
  DisjointBoxLayout boxes(domain, h*IntVect::Unit);
  LevelData<FArrayBox> u[3] = { LevelData<FArrayBox>(boxes, 1, 1), LevelData<FArrayBox>(boxes, 1, 1),
                                LevelData<FArrayBox>(boxes, 1, 1) };
  DataIterator dit(boxes);

  // initial pulse sin^6(pi (r + 1/2)) inside r <= 1/2 around the domain centre, un = unm1
  {
    const Real pi = 3.141592653589793;
    FArrayBox& fab = u[0][dit];
    fab.setVal(0.);
    MD_ARRAY_RESTRICT(a, fab);
    MD_BOXLOOP(domain, i)
      {
        const Real x = (i0 + 0.5)*dx, y = (i1 + 0.5)*dx, z = (i2 + 0.5)*dx;
        const Real r = std::sqrt(std::pow(x - 0.5, 2) + std::pow(y - 0.5, 2) + std::pow(z - 0.5, 2));
        a[MD_IX(i, 0)] = (r <= 0.5) ? std::pow(std::sin(pi*(r + 0.5)), 6) : 0.;
      }
    u[2][dit].setVal(0.);
    u[2][dit].copy(domain, fab);
    u[1][dit].setVal(0.);
  }

  StencilOps::ZeroGradientBC bc[3];
  for (int l = 0; l != 3; ++l) bc[l].define(u[l], domain);

  int n = 0, np1 = 1, nm1 = 2;
  // one untimed step to upload and warm up, then restart from the initial state
  bc[n].apply(u[n]);
  StencilOps::waveStep(u[np1], u[n], u[nm1], factor, SGC_FMA | SGC_WAVE_DIRSUM);
  SGC_SAFE_CALL(sgc_sync());

  Stopwatch<> timer;
  sgc_event_t e0, e1;
  SGC_SAFE_CALL(sgc_event_create(&e0));
  SGC_SAFE_CALL(sgc_event_create(&e1));
  timer.start();
  SGC_SAFE_CALL(sgc_event_record(e0));
  {
    // the whole iteration group stays on the device (one BC plan serves the three equal-geometry levels)
    int idx[3] = { n, np1, nm1 };
    StencilOps::waveRun(u, idx, bc[0], factor, numIter, SGC_FMA | SGC_WAVE_DIRSUM);
    n = idx[0]; np1 = idx[1]; nm1 = idx[2];
  }
  SGC_SAFE_CALL(sgc_event_record(e1));
  SGC_SAFE_CALL(sgc_sync());
  timer.stop();
  float ms = 0;
  SGC_SAFE_CALL(sgc_event_elapsed_ms(e0, e1, &ms));

  // FNV-1a over the valid cells of un, x fastest
  unsigned long long hsh = 1469598103934665603ull;
  double sum = 0;
  {
    const FArrayBox& fab = u[n][dit];
    for (int k = 0; k != h; ++k)
      for (int j = 0; j != h; ++j)
        {
          const unsigned char* s = reinterpret_cast<const unsigned char*>(&fab(IntVect(0, j, k), 0));
          for (size_t b = 0; b != sizeof(Real)*h; ++b) { hsh ^= s[b]; hsh *= 1099511628211ull; }
          for (int i = 0; i != h; ++i) sum += fab(IntVect(i, j, k), 0);
        }
  }
  std::cout << std::left << std::setw(40) << "Time for advance (ms): " << ms << "  (host wall " << timer.time() << ")"
            << std::endl;
  std::printf("{\"driver\": \"wave\", \"h\": %d, \"iters\": %d, \"factor\": %.17g, \"gcell_per_s\": %.3f, "
              "\"sum\": %.17g, \"un_valid_fnv1a64\": \"%016llx\"}\n",
              h, numIter, factor, (double)h*h*h*numIter/(ms*1e-3)/1e9, sum, hsh);
  sgc_event_destroy(e0);
  sgc_event_destroy(e1);
  return 0;
}
