// latticeBoltzmann — the reference's multi-box application (application/latticeBoltzmann/
// latticeBoltzmann.cpp, LBLevel.H, LBLevel.cpp, LBPatch.H) written against the facade: a D3Q19
// channel flow on a 19-component level, periodic in x and y, bounce-back walls in z, body force in x.
// The class below keeps the reference's LBLevel interface (initialData, advance, writePlotFile,
// computeTotalMass); every phase of advance() runs on the device:
//   collision -> exchange(PeriodicX|PeriodicY, TrimCorner) -> bounce-back fill -> stream -> macroscopic.
// The copy loops are written as in the reference — one BaseFab::copy per box and component — but
// RECORDED into a CopyBatch once and replayed in one launch per step.
//   usage: latticeBoltzmann [nx ny nz [box [nstep [plotEvery [stepsPerCall]]]]]
//          default 64 32 32 16 4001 200 (the reference's hard-wired run, latticeBoltzmann.cpp:16,36-49)
#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <utility>
#include <vector>

#include "BaseFab.H"
#include "Copier.H"
#include "DisjointBoxLayout.H"
#include "LevelData.H"
#include "StencilOps.H"
#include "Stopwatch.H"

namespace LBParameters
{
constexpr int g_numVelDir = 19;
constexpr int g_numGhost = 1;
constexpr Real g_weight[g_numVelDir] = {
  1./3.,
  1./18., 1./18., 1./18., 1./18., 1./18., 1./18.,
  1./36., 1./36., 1./36., 1./36., 1./36., 1./36.,
  1./36., 1./36., 1./36., 1./36., 1./36., 1./36. };
// -x +x -y +y -z +z, then the twelve edges (LBParameters.H:71-96)
constexpr int g_vel[g_numVelDir][3] = {
  { 0, 0, 0 }, { -1, 0, 0 }, { 1, 0, 0 }, { 0, -1, 0 }, { 0, 1, 0 }, { 0, 0, -1 }, { 0, 0, 1 },
  { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 }, { 1, 1, 0 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 },
  { 1, 0, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 }, { 0, 1, 1 } };
constexpr int g_opposite[g_numVelDir] = { 0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15 };
inline IntVect latticeVelocity(const int a_ei) { return IntVect(g_vel[a_ei][0], g_vel[a_ei][1], g_vel[a_ei][2]); }
inline int oppositeVelDir(const int a_ei) { return g_opposite[a_ei]; }
inline const char* const* stateNames()
{
  static const char* const names[] = { "density", "x-velocity", "y-velocity", "z-velocity" };
  return names;
}
}  // namespace LBParameters

class LBLevel
{
  using LevelSolData = LevelData<BaseFab<Real> >;

public:
  explicit LBLevel(const DisjointBoxLayout& a_dbl)
    : m_dbl(a_dbl),
      m_curr(a_dbl, LBParameters::g_numVelDir, LBParameters::g_numGhost),
      m_prev(a_dbl, LBParameters::g_numVelDir, LBParameters::g_numGhost),
      m_macro_comps(a_dbl, 4, LBParameters::g_numGhost)
  {
    initialData();
    m_copier.defineExchangeLD(m_curr, PeriodicX | PeriodicY, TrimCorner);
    recordBounceBack();
    recordStream();
  }

  void initialData()
  {
    for (int k = 0; k < LBParameters::g_numVelDir; ++k)
      {
        m_curr.setVal(k, LBParameters::g_weight[k]);
        m_prev.setVal(k, LBParameters::g_weight[k]);
      }
    m_macro_comps.setVal(0, 1.0);
    for (int k = 1; k < 4; ++k) m_macro_comps.setVal(k, 0.0);
  }

  /// a_nstep time steps (the reference's advance() is a_nstep = 1)
  void advance(const int a_nstep = 1)
  {
    if (StencilOps::lbAdvance(m_curr, m_prev, m_macro_comps, m_copier, m_bounce, m_stream, a_nstep))
      std::swap(m_curr, m_prev);
  }

  /// the reference's advance() spelled out phase by phase (five launches)
  void advancePhases()
  {
    StencilOps::lbCollision(m_curr, m_macro_comps);
    m_curr.exchange(m_copier);
    m_bounce.run(m_curr, m_curr);
    m_stream.run(m_prev, m_curr);
    std::swap(m_curr, m_prev);
    StencilOps::lbMacroscopic(m_curr, m_macro_comps);
  }

  /// Plot file: CGNS is not available, so the data the reference hands to CGNS (per box the vertex
  /// coordinates are implied by the box; per box and variable the valid cells) goes to a raw file.
  int writePlotFile(const int a_iter) const
  {
    std::vector<Real> data;
    m_macro_comps.gatherValid(data);
    char name[64];
    std::snprintf(name, sizeof name, "plot/solution%04d.sgcplot", a_iter);
    std::FILE* f = std::fopen(name, "wb");
    if (!f) return 1;
    std::fprintf(f, "SGCPLOT 1\nzones %d\nvariables 4", m_dbl.localSize());
    for (int v = 0; v != 4; ++v) std::fprintf(f, " %s", LBParameters::stateNames()[v]);
    std::fprintf(f, "\n");
    for (DataIterator dit(m_dbl); dit.ok(); ++dit)
      {
        const Box& b = m_dbl[dit];
        std::fprintf(f, "Zone_%06d %d %d %d %d %d %d\n", (*dit).globalIndex() + 1, b.loVect(0), b.loVect(1), b.loVect(2),
                     b.hiVect(0), b.hiVect(1), b.hiVect(2));
      }
    std::fprintf(f, "data %zu\n", data.size());
    const size_t n = std::fwrite(data.data(), sizeof(Real), data.size(), f);
    std::fclose(f);
    return n == data.size() ? 0 : 1;
  }

  /// sum of f_i over the valid cells (LBLevel.cpp:135-152), in the reference's order box -> iVel -> cells
  Real computeTotalMass() const
  {
    std::vector<Real> data;
    m_curr.gatherValid(data);
    Real mass = 0.;
    for (const Real v : data) mass += v;
    return mass;
  }

  const LevelSolData& macro() const { return m_macro_comps; }

protected:
  void recordStream()
  {
    // LBPatch::stream (LBPatch.H:49-64)
    for (DataIterator dit(m_dbl); dit.ok(); ++dit)
      for (int k = 0; k < LBParameters::g_numVelDir; ++k)
        {
          const IntVect shift_dir = LBParameters::latticeVelocity(LBParameters::oppositeVelDir(k));
          const Box dst_box = m_dbl[dit];
          Box src_box = dst_box;
          src_box.shift(shift_dir);
          m_stream.record(*dit, dst_box, k, *dit, src_box, k, 1);
        }
  }

  void recordBounceBack()
  {
    // LBLevel.cpp:65-112: no-slip walls at the top and the bottom of the domain
    static const int top[5][2] = { { 6, 5 }, { 13, 12 }, { 14, 11 }, { 17, 16 }, { 18, 15 } };   // from, into
    const Box& domain = m_dbl.problemDomain();
    for (DataIterator dit(m_dbl); dit.ok(); ++dit)
      {
        const Box& box = m_dbl[dit];
        int side = 0;
        if (box.hiVect(2) == domain.hiVect(2)) side = 1;
        else if (box.loVect(2) == domain.loVect(2)) side = -1;
        if (side == 0) continue;
        IntVect lo = box.loVect(), hi = box.hiVect();
        if (side > 0) lo[2] = hi[2]; else hi[2] = lo[2];
        const Box layer(lo, hi);
        for (int m = 0; m != 5; ++m)
          {
            const int from = side > 0 ? top[m][0] : top[m][1], into = side > 0 ? top[m][1] : top[m][0];
            Box ghost = layer;
            ghost.shift(LBParameters::latticeVelocity(from));
            m_bounce.record(*dit, ghost, into, *dit, layer, from, 1);
          }
      }
  }

  DisjointBoxLayout m_dbl;
  LevelSolData m_curr;
  LevelSolData m_prev;
  LevelSolData m_macro_comps;
  Copier m_copier;
  StencilOps::CopyBatch m_bounce;
  StencilOps::CopyBatch m_stream;
};

static unsigned long long fnv1a(const void* p, size_t n)
{
  const unsigned char* s = static_cast<const unsigned char*>(p);
  unsigned long long h = 1469598103934665603ull;
  for (size_t i = 0; i != n; ++i) { h ^= s[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, const char* argv[])
{
  const int nx = argc > 3 ? std::atoi(argv[1]) : 64;
  const int ny = argc > 3 ? std::atoi(argv[2]) : 32;
  const int nz = argc > 3 ? std::atoi(argv[3]) : 32;
  const int bs = argc > 4 ? std::atoi(argv[4]) : 16;
  const int nstep = argc > 5 ? std::atoi(argv[5]) : 4001;
  const int plotEvery = argc > 6 ? std::atoi(argv[6]) : 200;
  const int perCall = argc > 7 ? std::atoi(argv[7]) : 0;       // 0: as many as fit between two plot files

  DisjointBoxLayout::initMPI(argc, argv);
  Box domain(IntVect::Zero, IntVect(nx - 1, ny - 1, nz - 1));
  Stopwatch<std::chrono::steady_clock> stopwatch;
  stopwatch.start();
  DisjointBoxLayout dbl(domain, bs*IntVect::Unit);
  LBLevel lblvl(dbl);
  if (plotEvery > 0) mkdir("plot", 0755);

  int k = 0;
  while (k < nstep)
    {
      if (plotEvery > 0 && k % plotEvery == 0)
        {
          std::cout << "Writing during iteration " << k << std::endl;
          lblvl.writePlotFile(k);
        }
      int chunk = nstep - k;
      if (plotEvery > 0) chunk = std::min(chunk, plotEvery - k % plotEvery);
      if (perCall > 0) chunk = std::min(chunk, perCall);
      lblvl.advance(chunk);
      k += chunk;
    }
  SGC_SAFE_CALL(sgc_sync());
  stopwatch.stop();
  std::cout << stopwatch.time() << std::endl;

  std::vector<Real> macro;
  lblvl.macro().gatherValid(macro);
  const double cells = (double)nx*ny*nz;
  std::printf("{\"driver\": \"latticeBoltzmann\", \"domain\": [%d, %d, %d], \"box\": %d, \"nstep\": %d, \"ms\": %.3f, "
              "\"mcell_updates_per_s\": %.3f, \"mass\": %.17g, \"macro_fnv1a64\": \"%016llx\"}\n",
              nx, ny, nz, bs, nstep, stopwatch.time(), cells*nstep/(stopwatch.time()*1e-3)/1e6, lblvl.computeTotalMass(),
              fnv1a(macro.data(), sizeof(Real)*macro.size()));
  DisjointBoxLayout::finalizeMPI();
  return 0;
}
