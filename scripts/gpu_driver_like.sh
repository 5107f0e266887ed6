#!/bin/bash
# What the driver runs at round end on one GPU: smoke, GPU tests, the reference arm, then our arm with default settings.
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps ${STEPS:-20} --warmup 5 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps ${STEPS:-20} --warmup 5 ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/bench_ref.err; tail -6 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
r = json.loads(open('gpurun_out/bench_ref.log').read().strip().splitlines()[-1])
print('value %.4g  ms %.4f  e2e %.4g (frac %.2f of probe)  ref %.4g (%s, %d cores)  clocks %s' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['roofline']['frac'] or 0, r['value'], r['cpu_baseline']['kind'],
    r['cpu_baseline']['cores'], d['clocks']))
print('roofline', {k: d['roofline'][k] for k in ('bound', 'achieved', 'peak', 'frac', 'traffic')})
print('sustained', d.get('sustained'))
for k, v in (d.get('configs') or {}).items():
    if 'error' in v:
        print(k, 'ERROR', v['error']); continue
    if k == 'c3_full':
        print(k, 'R,T,A %.4g pts/s (%.1f ms)' % (v['R_T_A_to_hbm']['points_per_s'], v['R_T_A_to_hbm']['ms']), 'fused %.4g pts/s (%.1f ms)' % (v['T_plus_fused_peak_search']['points_per_s'], v['T_plus_fused_peak_search']['ms']), 'violations', v['property_violations']); continue
    rf = v.get('roofline') or {}
    print(k, '%.4g pts/s' % v.get('value', v.get('points_per_s_compute', 0)), 'ms', v.get('ms_per_step', v.get('compute_ms')), rf.get('bound'), rf.get('frac'))
print('ablation', {k: v['value'] for k, v in (d.get('ablation') or {}).items() if isinstance(v, dict)})
PY
