#!/bin/bash
# Developer tool: opcode histogram of the main loop of a k_sweep instantiation.
#   scripts/sass_loop_hist.sh build/fv2d_sweep_s1.o [mangled-name-fragment]
obj=${1:-build/fv2d_sweep_s1.o}; frag=${2:-ILi256ELb1ELi1ELi0ELb0ELi0E}
fun=$(cuobjdump -elf $obj 2>/dev/null | grep -o "_ZN4fv2d7k_sweep${frag}[A-Za-z0-9_]*" | sort -u | head -1)
cuobjdump -sass -fun "$fun" $obj > /tmp/_loop.sass
# loop = from the target of the last backward BRA.U to that branch
python3 - <<'PY'
import re
L=[l for l in open('/tmp/_loop.sass') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/',l)]
ins=[]
for l in L:
    m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
    ins.append((int(m.group(1),16),m.group(2).strip()))
# find largest backward branch span
best=None
for a,t in ins:
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s+)?(0x[0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a:
            n=sum(1 for b,u in ins if tgt<=b<=a and u.split()[0]=='DFMA')
            nb=sum(1 for b,u in ins if tgt<=b<=a and 'BAR.SYNC' in u)
            if n<100 or nb<1: continue
            key=(-(a-tgt),)
            if best is None or key>best[0]: best=(key,tgt,a)
_,lo,hi=best
body=[t for a,t in ins if lo<=a<=hi]
import collections
c=collections.Counter()
for t in body:
    toks=t.split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    c[op.split('.')[0]]+=1
print(f"loop {lo:#x}..{hi:#x}: {len(body)} instructions")
fp=sum(c[k] for k in ('DFMA','DMUL','DADD','DSETP'))
print("fp64:",fp, " ".join(f"{k}={v}" for k,v in c.most_common(24)))
PY
