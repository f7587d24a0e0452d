import json, sys
for f in sys.argv[1:]:
    try:
        for l in open(f):
            if l.startswith("{"):
                d = json.loads(l); k = d["kernels"]
                print(f, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None, d.get("exchange"), d["checksum_abs_mean_shift_z"])
                print("  ", {n: (v["ms_per_step"], v["launches_per_step"]) for n, v in k.items() if v["ms_per_step"] > 0.2})
    except Exception as e:
        print(f, "ERR", e)
