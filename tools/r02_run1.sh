#!/bin/bash
# round 2, GPU call 1: gradient parity at the cfg-2 shape (bf16x3 and plain-bf16 dh recurrence), model tests, schedule sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_run1_gpu.txt 2>&1
echo "== parity x3" ; timeout 600 python -m pytest tests/test_gpu_model.py -x -q -s -k "cfg2_shape_backward" 2>&1 | tail -60 | tee gpurun_out/r02_parity_x3.log
for t in nodrop dropout; do cp gpurun_out/r02_grad_parity_$t.json gpurun_out/r02_grad_parity_x3_$t.json 2>/dev/null; done
echo "== parity plain bf16" ; RS_BWD_X3=0 timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k "cfg2_shape_backward" 2>&1 | tail -60 | tee gpurun_out/r02_parity_x1.log
for t in nodrop dropout; do cp gpurun_out/r02_grad_parity_$t.json gpurun_out/r02_grad_parity_x1_$t.json 2>/dev/null; done
echo "== model tests" ; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -x -q 2>&1 | tail -15 | tee gpurun_out/r02_model_tests.log
run() { echo "== bench $*"; env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.readline()
try:
    d = json.loads(l); r = d['roofline']
    print('   value %.1f utt/s  %.2f ms/step  e2e %.1f  launches %d  fwd %s bwd %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], ['%.2f' % x for x in r['launch_ms']['fwd']], ['%.2f' % x for x in r['launch_ms']['bwd']]))
except Exception as e:
    print('   FAILED', l[:300])
"; }
run RS_BWD_X3=1 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_BWD_X3=0 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_TC_WINDOW=3 RS_TC_SIDE_TPC=1 RS_TC_DX_TPC=1 RS_TC_GX_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_TC_WINDOW=3 RS_TC_SIDE_TPC=2 RS_TC_DX_TPC=1 RS_TC_GX_TPC=1 RS_TC_CHUNK=64 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_TC_WINDOW=3 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_TC_WINDOW=2 RS_TC_SIDE_TPC=1 RS_TC_DX_TPC=1 RS_TC_GX_TPC=1 2>&1 | tee -a gpurun_out/r02_sweep1.log
run RS_TC_CHUNK=64 2>&1 | tee -a gpurun_out/r02_sweep1.log
