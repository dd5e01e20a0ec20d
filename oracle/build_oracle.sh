#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the checkers, never the product.
#
#  1. oracle/_build/libsgcoracle.so : the plain-C restatement (oracle/sgc_oracle.c)
#  2. oracle/_ref/libsgcref{,_nofma,_omp,_dbg}.so : the REAL reference library compiled from its
#     own sources where they lie under $REF_DIR (default /root/reference/Structured_gridcalc) with
#     direct g++ calls (the reference's make system needs csh, absent here — SURVEY.md §8c) plus
#     oracle/ref_shim.cpp.  Skipped (exit 0) when $REF_DIR is absent, e.g. on the GPU box, where
#     the prebuilt files from this container are used.
#  3. oracle/_ref/test*: the reference's 7 serial tests, built in debug mode and run (must pass).
#
# Flags follow the reference's release build (configure.ac:123-124,138: -O3 -funroll-loops
# -march=native -DRELEASE) except -march=x86-64-v3 (AVX2+FMA) instead of native, because the .so
# is built in this container and executed on the GPU box's host CPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_DIR="${REF_DIR:-/root/reference/Structured_gridcalc}"
MARCH="${SGC_ORACLE_MARCH:-x86-64-v3}"
CXX="${SGC_CXX:-g++}"   # not $CXX: this image exports a wrapper without libgomp.spec
CC="${SGC_CC:-gcc}"

mkdir -p "$HERE/_build"
# -ffp-contract=off: the restatement writes its fused multiply-adds explicitly (fma()).
CFL="-O3 -funroll-loops -march=$MARCH -std=gnu11 -ffp-contract=off -fPIC -shared -Wall -Wextra"
$CC $CFL -o "$HERE/_build/libsgcoracle.so" "$HERE/sgc_oracle.c" -lm
echo "[oracle] built _build/libsgcoracle.so"

if [ ! -d "$REF_DIR/lib/src/BoxFramework" ]; then
  echo "[oracle] $REF_DIR not present: keeping prebuilt oracle/_ref (if any)"
  exit 0
fi

S="$REF_DIR/lib/src/BoxFramework"
OUT="$HERE/_ref"
mkdir -p "$OUT"
LIBSRC="$S/IntVect.cpp $S/Box.cpp $S/BaseFab.cpp $S/DisjointBoxLayout.cpp $S/LevelData.cpp $S/LinuxSupport.cpp"
COMMON="-std=c++14 -Wno-vla -Wno-unknown-pragmas -DNO_CGNS -I$S -fPIC"
REL="-O3 -funroll-loops -march=$MARCH -DRELEASE"

$CXX $COMMON $REL                   -shared -o "$OUT/libsgcref.so"       $LIBSRC "$HERE/ref_shim.cpp"
$CXX $COMMON $REL -ffp-contract=off -shared -o "$OUT/libsgcref_nofma.so" $LIBSRC "$HERE/ref_shim.cpp"
$CXX $COMMON $REL -fopenmp          -shared -o "$OUT/libsgcref_omp.so"   $LIBSRC "$HERE/ref_shim.cpp"
$CXX $COMMON -O1 -g                 -shared -o "$OUT/libsgcref_dbg.so"   $LIBSRC "$HERE/ref_shim.cpp"
# the reference's single-precision build (Real = float, Parameters.H:17-23,197): same shim, float buffers
$CXX $COMMON $REL -DUSE_SINGLE_PRECISION                   -shared -o "$OUT/libsgcref_sp.so"       $LIBSRC "$HERE/ref_shim.cpp"
$CXX $COMMON $REL -DUSE_SINGLE_PRECISION -ffp-contract=off -shared -o "$OUT/libsgcref_sp_nofma.so" $LIBSRC "$HERE/ref_shim.cpp"
# the shipped wave time stepper (WavePatch, USE_VEX pencil path); cgnslib.h is included
# unconditionally by WavePatch.cpp:15 -> an empty stub on the include path
mkdir -p "$OUT/stub"
: > "$OUT/stub/cgnslib.h"
W="$REF_DIR/application/wave"
$CXX $COMMON $REL -I"$OUT/stub" -I"$W" -shared -o "$OUT/libsgcref_wave.so" $LIBSRC "$W/WavePatch.cpp" "$HERE/ref_wave_shim.cpp"
# the shipped lattice-Boltzmann application (the only multi-box one) + the library's CGNS plot
# writers, against the recording CGNS stand-in oracle/cgns_record/cgnslib.h (no -DNO_CGNS here)
LB="$REF_DIR/application/latticeBoltzmann"
LBFL="-std=c++14 -Wno-vla -Wno-unknown-pragmas -w -I$S -fPIC -I$HERE/cgns_record -I$LB"
$CXX $LBFL $REL                   -shared -o "$OUT/libsgcref_lb.so"       $LIBSRC "$LB/LBLevel.cpp" "$HERE/ref_lb_shim.cpp"
$CXX $LBFL $REL -ffp-contract=off -shared -o "$OUT/libsgcref_lb_nofma.so" $LIBSRC "$LB/LBLevel.cpp" "$HERE/ref_lb_shim.cpp"
$CXX $LBFL $REL -fopenmp          -shared -o "$OUT/libsgcref_lb_omp.so"   $LIBSRC "$LB/LBLevel.cpp" "$HERE/ref_lb_shim.cpp"
echo "[oracle] built _ref/libsgcref{,_nofma,_omp,_dbg,_sp,_sp_nofma,_wave,_lb,_lb_nofma,_lb_omp}.so from $REF_DIR"

# the reference's own serial tests, asserts live (no -DRELEASE)
T="$REF_DIR/lib/test/BoxFramework"
fail=0
for t in testIntVect testBox testBoxIterator testBaseFab testDisjointBoxLayout testLayoutIterator testLevelData; do
  $CXX $COMMON -O1 -o "$OUT/$t" "$T/$t.cpp" $LIBSRC
  if ( cd "$OUT" && "./$t" ) | tee "$OUT/$t.log" | grep -q passed; then :; else fail=1; echo "[oracle] reference test $t FAILED"; fi
done
# the shipped PR1 driver, unmodified (n=64): asserts B == 0
$CXX $COMMON -O3 -funroll-loops -march=$MARCH -o "$OUT/laplacian" "$REF_DIR/application/laplacian/laplacian.cpp" $LIBSRC
( cd "$OUT" && ./laplacian ) > "$OUT/laplacian.log"
echo "[oracle] reference serial tests: $([ $fail = 0 ] && echo all passed || echo FAILURES)"
exit $fail
