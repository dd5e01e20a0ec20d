L=structuredgridcalc-boxframework_b200/libsgc_b200.so
cp $L /tmp/orig.so
for f in variants/*.so; do cp $f $L; touch $L; echo "$f"; timeout 200 python profiles/micro/smallbox_probe.py 2>&1 | tail -2; done
cp /tmp/orig.so $L
