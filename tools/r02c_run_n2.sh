#!/bin/bash
# round 2 (third session), 2-GPU call: the in-library NCCL path, lockstep data parallel, bench at N=2 on the final build
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== mgpu lockstep"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_lockstep.py 2>&1 | tail -4 | tee gpurun_out/r02c_mgpu_lockstep.log
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r02c_bench_n2.err | tail -1 > gpurun_out/r02c_bench_cfg2_n2.json; tail -c 600 gpurun_out/r02c_bench_n2.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c_bench_cfg2_n2.json'))
print('N=2: value %.1f utt/s  %.2f ms/step  e2e %.1f  err-rate %.2f ms  allreduce %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['with_error_rate']['ms_per_step'], d['roofline']['families']['allreduce']['ms_per_step']))
PY
