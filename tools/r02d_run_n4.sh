#!/bin/bash
# round 2 (fourth session), 4-GPU call: bench line of the final build at N=4 (in-library NCCL, early read under data parallel)
N=${1:-4}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02d_bench_n$N.err | tail -1 > gpurun_out/r02d_bench_cfg2_n$N.json; tail -c 300 gpurun_out/r02d_bench_n$N.err
python - <<PY | tee gpurun_out/r02d_n$N.txt
import json
try:
    d = json.load(open('gpurun_out/r02d_bench_cfg2_n$N.json'))
    print('N=$N: value %.1f utt/s  %.2f ms/step  e2e %.1f (%.2f ms/step)  allreduce %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['families']['allreduce']['ms_per_step']))
except Exception as e:
    print('bench failed', e)
PY
