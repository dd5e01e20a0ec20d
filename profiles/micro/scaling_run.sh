#!/bin/bash
# weak scaling on ONE 8-GPU box (all N on the same GPUs), then the C4 shape (1024^3 over 8 GPUs)
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/scale_n$n.log 2>&1
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/scale_n$n.log 2>&1
  fi
  tail -1 gpurun_out/scale_n$n.log > gpurun_out/scale_n$n.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 \
  bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 --domain 1024 --domain-z 128 > gpurun_out/scale_c4.log 2>&1
tail -1 gpurun_out/scale_c4.log > gpurun_out/scale_c4.json
for f in gpurun_out/scale_n1.json gpurun_out/scale_n2.json gpurun_out/scale_n4.json gpurun_out/scale_n8.json gpurun_out/scale_c4.json; do cut -c1-130 $f; done
