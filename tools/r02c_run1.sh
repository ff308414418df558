#!/bin/bash
# round 2 (third session), GPU call 1: one-state-per-thread CTC lattice, layer-0 gx GEMMs hoisted into the wavefront's fill,
# validated-exchange forward kernel for batches of at most 16 rows (cfg-4)
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02c_gpu_tests_run1.log
echo "== ctc"; for v in 0 1; do echo "RS_CTC_LATTICE=$v"; RS_CTC_LATTICE=$v timeout 300 python tests/gpu_diag.py ctc 2>&1 | grep -v "^ *worst\|rowsum" ; done | tee gpurun_out/r02c_ctc_diag_run1.log
bench() { timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline']['families']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f ; fwd %.2f ctc %.2f bwd %.2f ms' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          [v for k, v in f.items() if k.startswith('lstm_stack_forward')][0]['ms_per_step'], [v for k, v in f.items() if k.startswith('ctc')][0]['ms_per_step'],
          [v for k, v in f.items() if k.startswith('lstm_stack_backward')][0]['ms_per_step']))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for h in 0 40 26 20; do echo "RS_TC_HOIST=$h"; RS_TC_HOIST=$h bench; done
echo "RS_CTC_LATTICE=0 (hoist default)"; RS_CTC_LATTICE=0 bench
for x in 0 1; do echo "cfg4 RS_TS_XCHG16=$x"; RS_TS_XCHG16=$x bench --config cfg4; done
} | tee gpurun_out/r02c_sweep1.log
timeout 300 python tests/gpu_diag.py trace > gpurun_out/r02c_trace_run1.txt 2>&1; head -5 gpurun_out/r02c_trace_run1.txt
