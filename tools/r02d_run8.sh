#!/bin/bash
# round 2 (fourth session), GPU call 8: back-off between a fetch and the first poll of the hint counter (RS_TS_POLLB[_BWD] cycles)
mkdir -p gpurun_out
bench() { timeout 400 python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>gpurun_out/last.err | tail -1 > gpurun_out/last.json; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/last.json'))
    f = d['roofline'].get('families', {})
    g = lambda p: ([v for k, v in f.items() if k.startswith(p)] or [{'ms_per_step': float('nan')}])[0]['ms_per_step']
    print('   %.2f ms/step  e2e %.2f; fwd %.2f bwd %.2f ms; launch ms fwd %s bwd %s' % (d['ms_per_step'], d['e2e']['ms_per_step'],
          g('lstm_stack_forward'), g('lstm_stack_backward'), ['%.2f' % x for x in d['roofline']['launch_ms']['fwd']], ['%.2f' % x for x in d['roofline']['launch_ms']['bwd']]))
except Exception as e:
    print('   bench failed', e); print(open('gpurun_out/last.err').read()[-1500:])
PY
}
{
for b in 0 1500 3000 4000; do
  echo "RS_TS_POLLB=$b RS_TS_POLLB_BWD=$((b * 3 / 2))"; RS_TS_POLLB=$b RS_TS_POLLB_BWD=$((b * 3 / 2)) bench
done
echo "RS_TS_POLLB=2000 RS_TS_POLLB_BWD=0"; RS_TS_POLLB=2000 RS_TS_POLLB_BWD=0 bench
echo "RS_TS_POLLB=0 RS_TS_POLLB_BWD=3000"; RS_TS_POLLB=0 RS_TS_POLLB_BWD=3000 bench
} | tee gpurun_out/r02d_sweep8.log
