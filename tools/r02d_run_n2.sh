#!/bin/bash
# round 2 (fourth session), 2-GPU call: end_batch's early read under data parallel (accumulator all-reduce + copy on a side
# stream behind the CTC kernel; the gradient all-reduce stays behind the backward pass), lockstep test, bench at N=2
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
echo "== mgpu lockstep"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_lockstep.py 2>&1 | tail -4 | tee gpurun_out/r02d_mgpu_lockstep.log
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/r02d_bench_n2.err | tail -1 > gpurun_out/last_n2.json; tail -c 300 gpurun_out/r02d_bench_n2.err; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last_n2.json'))
    print('   N=2: value %.1f utt/s  %.2f ms/step  e2e %.1f (%.2f ms/step)  allreduce %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['families']['allreduce']['ms_per_step']))
except Exception as e:
    print('   bench failed', e)
PY
}
{
echo "RS_EARLY_READ_DP=0"; RS_EARLY_READ_DP=0 run 29512
echo "RS_EARLY_READ_DP=1"; run 29513
cp gpurun_out/last_n2.json gpurun_out/r02d_bench_cfg2_n2.json
} | tee gpurun_out/r02d_sweep_n2.log
