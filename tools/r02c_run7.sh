#!/bin/bash
# round 2 (third session), GPU call 7: cfg-4 (5x1024, batch 16: 64 CTAs per recurrent launch) -- where the backward time goes
mkdir -p gpurun_out
for w in 2 1; do echo "== RS_TC_WINDOW=$w"; RS_TRACE_CFG=4 RS_TC_WINDOW=$w timeout 300 python tests/gpu_diag.py trace 2>&1 | cut -c1-330; done | tee gpurun_out/r02c_trace_cfg4_run7.txt
bench() { timeout 400 python bench.py --no-cpu-baseline "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
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
echo "cfg4 RS_TC_WINDOW=1"; RS_TC_WINDOW=1 bench --config cfg4 --steps 16 --warmup 8
echo "cfg4 RS_TC_WINDOW=2 RS_TC_CHUNK=64 RS_TC_CHUNK_FWD=96"; RS_TC_CHUNK=64 RS_TC_CHUNK_FWD=96 bench --config cfg4 --steps 16 --warmup 8
echo "cfg4 RS_TC_WINDOW=2 RS_TC_CHUNK=256 RS_TC_CHUNK_FWD=96"; RS_TC_CHUNK=256 RS_TC_CHUNK_FWD=96 bench --config cfg4 --steps 16 --warmup 8
} | tee gpurun_out/r02c_sweep7.log
