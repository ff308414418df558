#!/bin/bash
# round 2 (fourth session), GPU call 7: hint counter polled with K loads in flight (RS_TS_POLLQ, RS_TS_POLLD cycles apart)
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
for q in 1 2 4 3; do for d in 100 250; do
  if [ $q = 1 ] && [ $d = 250 ]; then continue; fi
  echo "RS_TS_POLLQ=$q RS_TS_POLLD=$d"; RS_TS_POLLQ=$q RS_TS_POLLD=$d bench
done; done
echo "RS_TS_POLLQ=4 RS_TS_POLLD=500"; RS_TS_POLLQ=4 RS_TS_POLLD=500 bench
echo "RS_TS_POLLQ=1"; RS_TS_POLLQ=1 bench
} | tee gpurun_out/r02d_sweep7.log
echo "== model tests with RS_TS_POLLQ=4"; RS_TS_POLLQ=4 timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/r02d_tests_run7.log
