#!/bin/bash
# two-GPU timing of the step loop next to the one-GPU timing on the SAME box (GPU-to-GPU variance between boxes
# is ~6 %, more than the cost of the halo): run 0 = peer-memory two-step passes, run 1 = SGC_NO_T2=1 (one-step)
for o in 0 1; do
  if [ $o = 1 ]; then export SGC_NO_T2=1; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$o \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > /tmp/n2_$o.json
  python - <<PY
import json
j = json.load(open("/tmp/n2_$o.json"))
print("no_t2=$o value %.1f  two-step launch %.4f ms  conserved %s" % (j["value"], j["roofline"]["avg_launch_ms"], j["sum_conserved"]))
PY
done
unset SGC_NO_T2
timeout 100 python profiles/micro/t2_probe.py perf | tail -1 | cut -c1-250
