"""Per-source-line instruction counts and stall samples from `ncu --page source --csv --print-source cuda,sass`.
   python tools/ncu_lines.py src.csv [N] [sort: inst|samples]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
key = sys.argv[3] if len(sys.argv) > 3 else "samples"
f = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].strip().isdigit(): continue
    d = dict(zip(hdr[4:], r[4:]))
    def num(k):
        try: return int(d.get(k, "0"))
        except ValueError: return 0
    stalls = {k[6:]: num(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
    out.append((f, int(r[0]), r[1].strip(), num("Warp Stall Sampling (All Samples)"), num("Instructions Executed"), stalls))
ts = sum(o[3] for o in out); ti = sum(o[4] for o in out)
print(f"total samples {ts}  total warp-instructions {ti}")
for f, ln, src, s, ins, st in sorted(out, key=lambda o: -(o[4] if key == "inst" else o[3]))[:N]:
    top = " ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100*s/ts:5.1f}%smp {100*ins/ti:5.1f}%ins {f}:{ln:4d} {src[:64]:64s} | {top}")
tt = collections.Counter()
for o in out:
    for k, v in o[5].items(): tt[k] += v
print({k: round(100 * v / max(1, sum(tt.values())), 1) for k, v in tt.most_common(10)})
