#!/bin/bash
# round 2, 8-GPU call: cfg-2 and cfg-4 at 8 x B200 (weak scaling, in-library NCCL all-reduce)
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
for cfg in cfg2 cfg4; do
  echo "== bench $cfg N=8"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config $cfg 2> gpurun_out/r02_bench_${cfg}_n8.err | tail -1 > gpurun_out/r02_bench_${cfg}_n8.json
  tail -c 400 gpurun_out/r02_bench_${cfg}_n8.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_${cfg}_n8.json'))
print('$cfg N=8: value %.1f %s  %.2f ms/step  e2e %.1f  err-rate %.2f ms  allreduce %.3f ms' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['roofline']['families']['allreduce']['ms_per_step']))
PY
done
