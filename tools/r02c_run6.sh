#!/bin/bash
# round 2 (third session), GPU call 6: CTC lattice in base 2 with the shift from three rows back; batch tiles of 32 by default
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02c_gpu_tests_run6.log
echo "== ctc"; timeout 300 python tests/gpu_diag.py ctc 2>&1 | grep -v "rowsum" | cut -c1-400 | tee gpurun_out/r02c_ctc_diag_run6.log
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   value %.1f %s  %.2f ms/step  e2e %.1f (%s); fwd %.2f ctc %.2f bwd %.2f ms; with_error_rate %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          d['e2e'].get('latency_ms', {}).get('p50'), g('lstm_stack_forward'), g('ctc'), g('lstm_stack_backward'), d.get('with_error_rate', {}).get('ms_per_step')))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
echo "cfg2 default"; bench
echo "cfg5 default (tiles of 32)"; bench --config cfg5
} | tee gpurun_out/r02c_sweep6.log
echo "== ncu ctc"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:ctc_lattice1 -s 18 -c 1 -f -o gpurun_out/r02c_ctc_lattice1_v3 python tests/gpu_diag.py ctc > gpurun_out/r02c_ncu_ctc.log 2>&1; tail -2 gpurun_out/r02c_ncu_ctc.log
