#!/bin/bash
# A/B of library variants (variants/*.so) on ONE GPU with the bench's own step loop, interleaved, three rounds.
L=structuredgridcalc-boxframework_b200/libsgc_b200.so
cp $L /tmp/orig.so
for rep in 1 2 3; do
  for f in variants/*.so; do
    cp $f $L; touch $L
    echo -n "$f: "
    python bench.py --no-configs --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read()); r=j['roofline']; print('%.3f ms/step  %.1f Gcell/s  t2 %.4f ms x %d' % (j['ms_per_step'], j['value'], r['avg_launch_ms'], r['launches_timed']))"
  done
done
cp /tmp/orig.so $L
