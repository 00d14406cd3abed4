#!/usr/bin/env python
"""Developer tool: attribute the SASS of the k_sweep main loop to source lines.
   python scripts/sass_lines.py build/fv2d_sweep_s1.o [ILi256ELb1ELi1ELi0ELb0ELi0E]"""
import collections, os, re, subprocess, sys, tempfile
obj = sys.argv[1] if len(sys.argv) > 1 else "build/fv2d_sweep_s1.o"
frag = sys.argv[2] if len(sys.argv) > 2 else "ILi256ELb1ELi1ELi0ELb0ELi0E"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], cwd=d, capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.strip().startswith(".section") and ".text._ZN4fv2d7k_sweep" + frag in l][0]
end = next((i for i in range(start + 1, len(txt)) if txt[i].strip().startswith(".section")), len(txt))
cur, ins, labels = None, [], {}
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2)); continue
    m = re.match(r'^(\.L_x_\d+):', l)
    if m:
        labels[m.group(1)] = len(ins); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
# main loop = smallest backward-branch span that holds a BAR.SYNC and >= 100 DFMA
best = None
for k, (a, t, _) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+.*`\((\.L_x_\d+)\)', t)
    if m and m.group(1) in labels and labels[m.group(1)] < k:
        lo = labels[m.group(1)]
        body = ins[lo:k + 1]
        if sum('BAR.SYNC' in x[1] for x in body) >= 1 and sum(x[1].startswith('DFMA') for x in body) >= 100:
            if best is None or k - lo < best[1] - best[0]:
                best = (lo, k)
lo, hi = best
body = ins[lo:hi + 1]
nb = sum('BAR.SYNC' in x[1] for x in body)
print(f"main loop: {len(body)} instructions for {nb} row(s)")
def op(t):
    toks = t.split()
    return (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
tot = collections.Counter(op(t) for _, t, _ in body)
fp = sum(tot[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print(f"per row: {len(body)/nb:.0f} total, {fp/nb:.0f} fp64;", " ".join(f"{k}={v/nb:.0f}" for k, v in tot.most_common(22)))
src = open("fv2d_b200/csrc/fv2d_sweep.cu").read().split("\n")
by = collections.defaultdict(collections.Counter)
for _, t, ln in body:
    by[ln][op(t)] += 1
print("\nline  n/row  source | mix")
for ln in sorted(by, key=lambda x: (x is None, x)):
    c = by[ln]; n = sum(c.values())
    s = src[ln - 1].strip()[:70] if ln and ln <= len(src) else "?"
    print(f"{ln!s:>5} {n/nb:5.1f}  {s:70s} | " + " ".join(f"{k}={v}" for k, v in c.most_common(6)))
