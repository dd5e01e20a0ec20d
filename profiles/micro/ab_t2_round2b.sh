#!/bin/bash
# A/B of two-step-sweep variants (variants/*.so) on ONE GPU: parity of the variants named in $PARITY against the
# oracle (profiles/micro/t2_probe.py parity), then the 100-sweep loop timed twice per variant, interleaved, then
# the physical-boundary instance of those named in $BND.
L=structuredgridcalc-boxframework_b200/libsgc_b200.so
cp $L /tmp/orig.so
line() { timeout 120 python profiles/micro/t2_probe.py perf $1 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.4f ms two-step, %.1f Gcell/s overall' % (j['two_step_ms'], j['gcell_per_s']))"; }
for v in $PARITY; do
  cp variants/$v.so $L; touch $L
  echo -n "$v parity: "
  timeout 300 python profiles/micro/t2_probe.py parity 2>&1 | grep -E "MISMATCH|PARITY|Error|error" | tr '\n' ' '
  echo
done
for rep in 1 2; do
  for f in variants/*.so; do
    cp $f $L; touch $L
    echo -n "$f: "; line 7
  done
done
for v in $BND; do
  cp variants/$v.so $L; touch $L
  echo -n "$v BND: "; line 0
done
cp /tmp/orig.so $L
