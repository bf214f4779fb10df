import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f"{f}: n={d['n_gpus']} value={d['value']:.4g} e2e={d['e2e']['value']:.4g} ms/step={d['ms_per_step']:.3f} sweep_us={d['us_per_dipole_iteration']:.1f} "
          f"stage={ {k: round(v,3) for k,v in d['stage_ms'].items()} } launches={d['gpu_launches']} frac={d['roofline']['frac']:.3f}")
