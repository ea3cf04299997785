import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, {k:d.get(k) for k in ("value","n_gpus","ms_per_frame","gpu_launches","multi_gpu_frame_matches_single_gpu","rank_render_ms_per_step")}, "e2e", round(d["e2e"]["value"],1), d["config"]["workload"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-300:])
