"""Opcode histogram of the innermost loop that holds the packed-FP work of a kernel.
usage: sass_hot.py dump.sass <name-substring> [min-FFMA2-count]"""
import collections
import re
import sys

txt = open(sys.argv[1]).read()
need = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if sys.argv[2] not in name:
        continue
    ins = []
    for l in f.split("\n"):
        mm = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if mm:
            ins.append((int(mm.group(1), 16), mm.group(2).strip()))
    loops = []
    for addr, t in ins:
        mb = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
        if mb and int(mb.group(1), 16) <= addr:
            loops.append((int(mb.group(1), 16), addr))
    best = None
    for lo, hi in sorted(loops, key=lambda x: x[1] - x[0]):
        body = [t for a, t in ins if lo <= a <= hi]
        nfp = sum(1 for t in body if re.search(r"\b(FFMA2?|FMUL2?|FADD2?)\b", t))
        if nfp >= need:
            best = (lo, hi, body)
            break
    if not best:
        print(name[:70], "no loop")
        continue
    lo, hi, body = best
    c = collections.Counter()
    for t in body:
        p = t.split()
        op = p[1] if p[0].startswith("@") else p[0]
        c[op.split(".")[0]] += 1
    print(f"{name[:70]} [{lo:#x},{hi:#x}] {len(body)} instr")
    print("   ", dict(c.most_common(30)))
