"""Per-kernel SASS opcode counts of libfumi_b200.so (evidence of which kernels are tcgen05 / TMA and which are mma.sync):
   python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fumi_b200", "lib", "libfumi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA.16816", "HMMA.1688", "FFMA", "LDSM", "LDGSTS", "MUFU", "BAR.SYNC", "RED.E", "ATOM"]
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("fumi_epi::", "").replace("void ", "")
        name = re.sub(r"\((?!.*\().*$", "", name) if name.endswith(")") else name
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[name][k] += 1
print("SASS opcode counts per kernel (cuobjdump -sass fumi_b200/lib/libfumi_b200.so, sm_100a).")
print("UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load, HMMA = mma.sync, LDGSTS = cp.async\n")
hdr = f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{k:>10s}" for k in KEYS)
print(hdr)
for n, c in counts.items():
    print(f"{n[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[k]:10d}" for k in KEYS))
