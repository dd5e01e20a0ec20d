#!/bin/bash
# All N back to back on ONE 8xB200 box (B200s of the pool differ by a few per cent: never compare across boxes).
# Usage (on the GPU box): bash profiles/micro/scaling_run.sh [tag]   -> gpurun_out/<tag>_n{1,2,4,8}.json
tag=${1:-scale}
python bench.py --no-configs > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
port=29600
for n in 2 4 8; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
done
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus 8 --config C4 --e2e-steps 0 > gpurun_out/${tag}_c4.json 2> gpurun_out/${tag}_c4.err
for f in gpurun_out/${tag}_n*.json gpurun_out/${tag}_c4.json; do python - "$f" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1]))
    e = j.get("e2e") or {}
    print(sys.argv[1], "N=%d value=%.1f ms/step=%.3f e2e=%s ceiling=%s bw=%s parity=%s" % (
        j["n_gpus"], j["value"], j["ms_per_step"], e.get("value"), e.get("ceiling_value"), e.get("host_bw_ceiling_gbs"),
        (j.get("parity_nranks") or j.get("parity") or {}).get("ok")))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
