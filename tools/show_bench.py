import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("primary", round(d["value"]), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), "parity", {k:v for k,v in (d.get("parity") or {}).items() if k not in("grad_relerr","checker","tolerance")})
print("cpu", d["cpu_baseline"])
for k,v in (d.get("secondary") or {}).items():
    print(f"{k:18s} {round(v['value']):9d} {v['unit']:10s} ms {v['ms_per_step']:8.3f} e2e {round(v['e2e']['value']):9d} frac {v['roofline']['whole_step']['frac']:.4f} cpu {v['cpu_baseline'] and round(v['cpu_baseline']['value'],2)}")
