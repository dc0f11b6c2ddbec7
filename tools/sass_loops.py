"""Opcode histogram of the largest backward-branch loop of every kernel in a cuobjdump -sass dump (stdin or file)."""
import collections
import re
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
only = sys.argv[2] if len(sys.argv) > 2 else None
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if only and only not in name:
        continue
    ins = []
    for l in f.split("\n"):
        mm = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if mm:
            ins.append((int(mm.group(1), 16), mm.group(2).strip()))
    loops = []
    for addr, t in ins:
        mb = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if mb:
            tgt = int(mb.group(1), 16)
            if tgt <= addr:
                loops.append((tgt, addr))
    if not loops:
        print(name[:60], "no loop")
        continue
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 1]:
        body = [t for a, t in ins if lo <= a <= hi]
        c = collections.Counter()
        for t in body:
            p = t.split()
            op = p[1] if p[0].startswith("@") else p[0]
            c[op.split(".")[0] + ("(pred)" if p[0].startswith("@") and "BRA" not in t else "")] += 1
        print(f"{name[:50]:50s} [{lo:#x},{hi:#x}] {len(body):4d} ", dict(c.most_common(12)))
