/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * ref_lb_shim.cpp: drives the REFERENCE's lattice-Boltzmann application
 * (application/latticeBoltzmann/{LBLevel.H,LBLevel.cpp,LBPatch.H,LBPhysics.H,LBParameters.H},
 * compiled unmodified) — the only shipped multi-box application: 19-component D3Q19 level,
 * collision, exchange(PeriodicX|PeriodicY, TrimCorner), bounce-back ghost fill, shifted-copy
 * stream, macroscopic moments (LBLevel.cpp:7-127) — and its plot writer (LBLevel.H:96-188)
 * through the recording CGNS stand-in oracle/cgns_record/cgnslib.h.
 * Built by oracle/build_oracle.sh into oracle/_ref/libsgcref_lb{,_nofma}.so together with
 * LBLevel.cpp and the BoxFramework library sources (WITHOUT -DNO_CGNS, so that the library's
 * CGNS writers are compiled too).
 */
#include <cstring>
#include <string>
#include <vector>

#include "LBLevel.H"

// defined (non-inline) in LBPatch.H, which only LBLevel.cpp includes
namespace LBPatch
{
using SolFab = BaseFab<Real>;
void macroscopic(DisjointBoxLayout& a_dbl, LevelData<SolFab>& curr, LevelData<SolFab>& macro);
void collision(LevelData<SolFab>& curr, LevelData<SolFab>& macro, DisjointBoxLayout& a_dbl);
void stream(DisjointBoxLayout& a_dbl, LevelData<SolFab>& m_curr, LevelData<SolFab>& m_prev);
}

namespace
{

struct LBProbe : public LBLevel
{
  explicit LBProbe(DisjointBoxLayout& a_dbl) : LBLevel(a_dbl) { }
  LevelData<BaseFab<Real> >& curr() { return m_curr; }
  LevelData<BaseFab<Real> >& prev() { return m_prev; }
  LevelData<BaseFab<Real> >& macro() { return m_macro_comps; }
  DisjointBoxLayout& dbl() { return m_dbl; }
};

void levelIn(LevelData<BaseFab<Real> >& a_lvl, DisjointBoxLayout& a_dbl, const double* a_src)
{
  if (!a_src) return;
  size_t off = 0;
  for (DataIterator dit(a_dbl); dit.ok(); ++dit)
    {
      BaseFab<Real>& fab = a_lvl[dit];
      std::memcpy(fab.dataPtr(), a_src + off, sizeof(Real)*fab.size());
      off += fab.size();
    }
}

void levelOut(LevelData<BaseFab<Real> >& a_lvl, DisjointBoxLayout& a_dbl, double* a_dst)
{
  if (!a_dst) return;
  size_t off = 0;
  for (DataIterator dit(a_dbl); dit.ok(); ++dit)
    {
      BaseFab<Real>& fab = a_lvl[dit];
      std::memcpy(a_dst + off, fab.dataPtr(), sizeof(Real)*fab.size());
      off += fab.size();
    }
}

//--The record the CGNS stand-in appends to

struct Record
{
  std::vector<double> data;     // every value "written", in call order
  std::vector<int> meta;        // per call: kind (1 zone, 2 coord, 3 field), zone, count, name hash
  std::string names;            // names, '\n' separated, in call order
  int nzone = 0;
};
Record g_rec;

void note(int kind, int zone, int count, const char* name)
{
  g_rec.meta.push_back(kind);
  g_rec.meta.push_back(zone);
  g_rec.meta.push_back(count);
  g_rec.names += name;
  g_rec.names += '\n';
}

int g_zoneCells[1 << 16][3];

}  // namespace

extern "C" {

/*--------------------------------------------------------------------*
 * CGNS stand-in (see oracle/cgns_record/cgnslib.h)
 *--------------------------------------------------------------------*/

int cg_open(const char*, int, int* fn) { *fn = 1; return 0; }
int cg_close(int) { return 0; }
int cg_base_write(int, const char*, int, int, int* B) { *B = 1; return 0; }
void cg_error_print(void) { }

int cg_zone_write(int, int, const char* zonename, const cgsize_t* size, enum ZoneType_t, int* Z)
{
  *Z = ++g_rec.nzone;                       // CGNS zone indices start at 1
  // size = { vertices[3], cells[3], boundary vertices[3] }
  for (int d = 0; d != 3; ++d) g_zoneCells[*Z][d] = size[3 + d];
  note(1, *Z, 0, zonename);
  for (int d = 0; d != 9; ++d) g_rec.meta.push_back(size[d]);
  return 0;
}

int cg_coord_write(int, int, int Z, enum DataType_t, const char* coordname, const void* coord_ptr, int* C)
{
  const int n = (g_zoneCells[Z][0] + 1)*(g_zoneCells[Z][1] + 1)*(g_zoneCells[Z][2] + 1);
  const double* p = static_cast<const double*>(coord_ptr);
  g_rec.data.insert(g_rec.data.end(), p, p + n);
  note(2, Z, n, coordname);
  *C = 1;
  return 0;
}

int cg_sol_write(int, int, int, const char*, enum GridLocation_t, int* S) { *S = 1; return 0; }

int cg_field_general_write(int, int, int Z, int, enum DataType_t, const char* fieldname,
                           const cgsize_t* rmin, const cgsize_t* rmax, int m_numdim, const cgsize_t* m_dims,
                           const cgsize_t* m_rmin, const cgsize_t* m_rmax, const void* field_ptr, int* F)
{
  // the file range must be the whole zone and match the memory range in shape
  for (int d = 0; d != m_numdim; ++d)
    if (rmin[d] != 1 || rmax[d] != g_zoneCells[Z][d] || m_rmax[d] - m_rmin[d] != rmax[d] - rmin[d]) return 1;
  const double* p = static_cast<const double*>(field_ptr);
  int n = 0;
  for (int k = m_rmin[2]; k <= m_rmax[2]; ++k)          // 1-based, first index fastest
    for (int j = m_rmin[1]; j <= m_rmax[1]; ++j)
      for (int i = m_rmin[0]; i <= m_rmax[0]; ++i, ++n)
        g_rec.data.push_back(p[(i - 1) + (size_t)m_dims[0]*((j - 1) + (size_t)m_dims[1]*(k - 1))]);
  note(3, Z, n, fieldname);
  *F = 1;
  return 0;
}

/*--------------------------------------------------------------------*
 * LB drivers
 *--------------------------------------------------------------------*/

/* LBLevel on `domain` cut into maxbox boxes (latticeBoltzmann.cpp:16,36-37), optional initial
 * fi / macroscopic state (fabs concatenated in box order, ghosts included; NULL keeps
 * LBLevel::initialData), nstep LBLevel::advance() calls.  Copies fi and the macroscopic state
 * out and returns LBLevel::computeTotalMass().                                              */
double sgcref_lb_run(const int dlo[3], const int dhi[3], const int maxbox[3], int nstep,
                     const double* fi_init, const double* macro_init, double* fi_out, double* macro_out)
{
  DisjointBoxLayout dbl(Box(IntVect(dlo[0], dlo[1], dlo[2]), IntVect(dhi[0], dhi[1], dhi[2])),
                        IntVect(maxbox[0], maxbox[1], maxbox[2]));
  LBProbe lb(dbl);
  levelIn(lb.curr(), lb.dbl(), fi_init);
  levelIn(lb.macro(), lb.dbl(), macro_init);
  for (int s = 0; s != nstep; ++s) lb.advance();
  levelOut(lb.curr(), lb.dbl(), fi_out);
  levelOut(lb.macro(), lb.dbl(), macro_out);
  return lb.computeTotalMass();
}

/* One phase of the step in isolation: 0 = LBPatch::collision (fi in place, macro read),
 * 1 = LBPatch::macroscopic (macro written), 2 = LBPatch::stream (prev <- shifted fi, then the
 * two levels swap: fi receives what the reference calls m_curr afterwards, prev the other).  */
int sgcref_lb_phase(const int dlo[3], const int dhi[3], const int maxbox[3], int phase,
                    double* fi, double* macro, double* prev)
{
  DisjointBoxLayout dbl(Box(IntVect(dlo[0], dlo[1], dlo[2]), IntVect(dhi[0], dhi[1], dhi[2])),
                        IntVect(maxbox[0], maxbox[1], maxbox[2]));
  LBProbe lb(dbl);
  levelIn(lb.curr(), lb.dbl(), fi);
  levelIn(lb.macro(), lb.dbl(), macro);
  levelIn(lb.prev(), lb.dbl(), prev);
  switch (phase)
    {
    case 0: LBPatch::collision(lb.curr(), lb.macro(), lb.dbl()); break;
    case 1: LBPatch::macroscopic(lb.dbl(), lb.curr(), lb.macro()); break;
    case 2: LBPatch::stream(lb.dbl(), lb.curr(), lb.prev()); break;
    default: return -1;
    }
  levelOut(lb.curr(), lb.dbl(), fi);
  levelOut(lb.macro(), lb.dbl(), macro);
  levelOut(lb.prev(), lb.dbl(), prev);
  return 0;
}

/* LBLevel::writePlotFile (LBLevel.H:96-188) of the state after nstep steps, through the recording
 * stand-in.  Two-call protocol: with data == NULL returns the sizes; otherwise fills the buffers.
 *   sizes[0] = number of doubles, sizes[1] = number of meta ints, sizes[2] = bytes of names.  */
int sgcref_lb_plot(const int dlo[3], const int dhi[3], const int maxbox[3], int nstep, const double* macro_init,
                   long long sizes[3], double* data, int* meta, char* names)
{
  DisjointBoxLayout dbl(Box(IntVect(dlo[0], dlo[1], dlo[2]), IntVect(dhi[0], dhi[1], dhi[2])),
                        IntVect(maxbox[0], maxbox[1], maxbox[2]));
  LBProbe lb(dbl);
  levelIn(lb.macro(), lb.dbl(), macro_init);
  for (int s = 0; s != nstep; ++s) lb.advance();
  g_rec = Record();
  const int err = lb.writePlotFile(nstep);
  sizes[0] = (long long)g_rec.data.size();
  sizes[1] = (long long)g_rec.meta.size();
  sizes[2] = (long long)g_rec.names.size();
  if (data) std::memcpy(data, g_rec.data.data(), sizeof(double)*g_rec.data.size());
  if (meta) std::memcpy(meta, g_rec.meta.data(), sizeof(int)*g_rec.meta.size());
  if (names) std::memcpy(names, g_rec.names.data(), g_rec.names.size());
  return err;
}

}  // extern "C"
