#!/usr/bin/env bash
# Builds the C++ programs that sit ABOVE the C ABI (include/BoxFramework facade + libsgc_b200.so):
#   apps/_build/laplacian, apps/_build/levelheat      — this repo's drivers (apps/*.cpp)
#   apps/_build/ref_test*                               — the REFERENCE's own serial test programs,
#       compiled UNMODIFIED from /root/reference against this repo's facade headers (drop-in proof);
#       only when /root/reference is present (binaries travel to the GPU box with the snapshot).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
PKG="$ROOT/structuredgridcalc-boxframework_b200"
REF_DIR="${REF_DIR:-/root/reference/Structured_gridcalc}"
CXX="${SGC_CXX:-g++}"
OUT="$HERE/_build"
mkdir -p "$OUT"
FL="-std=c++17 -O2 -Wall -Wno-vla -Wno-unknown-pragmas -I$ROOT/include -I$ROOT/include/BoxFramework"
LD="-L$PKG -lsgc_b200 -Wl,-rpath,\$ORIGIN/../../structuredgridcalc-boxframework_b200"
for app in laplacian levelheat wave latticeBoltzmann; do
  [ -f "$HERE/$app.cpp" ] && $CXX $FL -o "$OUT/$app" "$HERE/$app.cpp" $LD
done
if [ -d "$REF_DIR/lib/test/BoxFramework" ]; then
  for t in testIntVect testBox testBoxIterator testBaseFab testDisjointBoxLayout testLayoutIterator testLevelData; do
    $CXX $FL -o "$OUT/ref_$t" "$REF_DIR/lib/test/BoxFramework/$t.cpp" $LD
  done
  # the reference's VLA / MD_ARRAY address-identity program (application/vlaloops/vlaloops.cpp:30-95), unmodified
  $CXX $FL -o "$OUT/ref_vlaloops" "$REF_DIR/application/vlaloops/vlaloops.cpp" $LD
  echo "[apps] built the reference's 7 serial tests and vlaloops against the B200 facade"
fi
echo "[apps] done"
