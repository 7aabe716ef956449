"""List the loops (backward branches) of a kernel in a cuobjdump -sass dump with their
instruction mix.  Usage: python tools/sass_loops.py file.sass [min_len]"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
minlen = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []
for ln in lines:
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_ix = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.\w+)*\s+(?:`\(\.L_x_\d+\)|0x([0-9a-f]+))', t)
    if m and m.group(1):
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_ix: loops.append((addr_ix[tgt], i))
for s, e in sorted(loops, key=lambda x: x[1] - x[0]):
    n = e - s + 1
    if n < minlen: continue
    ops = collections.Counter()
    for a, t in ins[s:e + 1]:
        t = re.sub(r'^@!?U?P\d+\s+', '', t)
        op = t.split()[0].split('.')[0]
        ops[op] += 1
    print(f"loop {ins[s][0]:#06x}-{ins[e][0]:#06x} len {n}: " + " ".join(f"{k}:{v}" for k, v in ops.most_common(18)))
