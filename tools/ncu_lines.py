"""Attribute ncu per-instruction counts to CUDA source lines using nvdisasm -g line info.
   python tools/ncu_lines.py rep.ncu-rep obj.o mangled_kernel_name"""
import csv, io, re, subprocess, sys, os, collections
rep, obj, kname = sys.argv[1:4]
tmp = "/tmp/ncu_lines"; os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {os.path.abspath(obj)} > /dev/null", shell=True, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# find kernel section
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." + kname in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].strip().startswith(".section")), len(dis))
cur = ("?", 0); lines = []   # per instruction: (file, line)
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
print("sass instr in disasm", len(lines), "in ncu", len(data))
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for i, r in enumerate(data):
    key = lines[i] if i < len(lines) else ("?", 0)
    a = agg[key]; a[0] += f(r, "Instructions Executed"); a[1] += f(r, "# Samples"); a[2] += f(r, "Thread Instructions Executed")
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
srcs = {}
def srcline(fn, ln):
    if fn not in srcs:
        for d in ("raytracerwin_b200/csrc", "include"):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, fn)
            if os.path.exists(p): srcs[fn] = open(p).read().splitlines(); break
        else: srcs[fn] = []
    L = srcs[fn]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
print("%inst %samp thr/inst  file:line  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 60]:
    print(f"{100*a[0]/ti:5.1f} {100*a[1]/ts:5.1f} {a[2]/max(a[0],1):5.1f}  {key[0]}:{key[1]}  {srcline(*key)}")
