import sys, json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d['roofline']; r.setdefault('achieved', 0); r.setdefault('peak', 0); r.setdefault('hbm_store', {}); r['hbm_store'].setdefault('achieved', r['hbm_store'].get('achieved_gbs', 0))
        print(f"{f[17:-4]:10s} value {d['value']:.4g} ms {d['ms_per_step']:.4g} lsteps/s {d['config']['layer_steps_per_s']:.4g} "
              f"roof {r['achieved']:.2f}/{r['peak']:.2f} hbm {r['hbm_store']['achieved']:.0f} e2e {d['e2e']['value']:.4g} same={d['e2e']['identical_to_device_path']}")
    except Exception as e:
        print(f, 'ERR', e, open(f).read()[-300:])
