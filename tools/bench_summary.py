import json, sys
for p in sys.argv[1:]:
    try:
        d=json.loads(open(p).read().strip().splitlines()[-1])
        e=d.get("e2e") or {"value":0}
        print(p, "value %.0f img/s (%.1f ms) e2e %.0f | fwd %.0f GB/s (%.3f) %.2f ms | bwd %.0f GB/s (%.3f) %.2f ms | launches %d clocks %s"%(d["value"], d["ms_per_step"], e["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["launch_ms"], d["roofline_backward"]["achieved"], d["roofline_backward"]["frac"], d["roofline_backward"]["launch_ms"], d["gpu_launches"], d["clocks"]))
    except Exception as ex:
        print(p, "failed", ex); print(open(p).read()[-1200:])
