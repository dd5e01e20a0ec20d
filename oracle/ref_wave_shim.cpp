/* TEST INFRASTRUCTURE — not part of the product path.
 *
 * ref_wave_shim.cpp: drives the REFERENCE's shipped wave time stepper, WavePatch
 * (application/wave/WavePatch.{H,cpp}, compiled unmodified with its USE_VEX pencil path,
 * WavePatch.cpp:186-223), through its own public interface.  Built by oracle/build_oracle.sh into
 * oracle/_ref/libsgcref_wave.so with an empty cgnslib.h stub on the include path and -DNO_CGNS
 * (SURVEY §8c).  Used by oracle/gen_golden.py to pin the "dirsum" wave form of the oracle.
 */
#include <cstring>

#include "Parameters.H"
#include "Box.H"
#include "BaseFab.H"
#include "WavePatch.H"

extern "C" {

/* h^3 domain (0..h-1), single box, c = 1, dx = 1/h, cfl as given (wave.cpp:110-116).
 * u_init / um1_init: optional initial un / unm1 on grow(domain,1) (NULL: WavePatch::initialData()).
 * After nstep advance() calls copies un and unm1 (fabs on grow(domain,1), reference order) out.
 * Returns factor = (dt*c/dx)^2/3 through *factor.                                           */
int sgcref_wavepatch_run(int h, double cfl, int nstep, const double* u_init, const double* um1_init,
                         double* un_out, double* unm1_out, double* factor)
{
  const Box domain(IntVect::Zero, (h - 1)*IntVect::Unit);
  const Real dx = 1./h, c = 1.;
  WavePatch patch(domain, h*IntVect::Unit, "plot/plot.", c, dx, cfl);
  patch.initialData();
  if (u_init) std::memcpy(patch.un().dataPtr(), u_init, sizeof(Real)*patch.un().size());
  if (um1_init) std::memcpy(patch.unm1().dataPtr(), um1_init, sizeof(Real)*patch.unm1().size());
  for (int s = 0; s != nstep; ++s) patch.advance();
  std::memcpy(un_out, patch.un().dataPtr(), sizeof(Real)*patch.un().size());
  std::memcpy(unm1_out, patch.unm1().dataPtr(), sizeof(Real)*patch.unm1().size());
  const Real dt = dx*cfl/c;
  if (factor) *factor = (dt*c/dx)*(dt*c/dx)/3;
  return patch.un().size();
}

}  // extern "C"
