L=structuredgridcalc-boxframework_b200/libsgc_b200.so
cp $L /tmp/orig.so
for rep in 1 2; do for f in variants/*.so; do cp $f $L; touch $L; echo -n "$f: "; timeout 120 python profiles/micro/t2_probe.py perf 7 32 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('%.4f ms two-step, %.1f Gcell/s overall' % (j['two_step_ms'], j['gcell_per_s']))"; done; done
cp /tmp/orig.so $L
