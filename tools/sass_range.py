"""print the SASS of one kernel between two addresses: sass_range.py dump.sass <name-substring> <lo-hex> <hi-hex>"""
import re, sys
txt = open(sys.argv[1]).read()
lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    if sys.argv[2] not in f.split("\n")[0]:
        continue
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m and lo <= int(m.group(1), 16) <= hi:
            print(m.group(1), m.group(2).strip())
    break
