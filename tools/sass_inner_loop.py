#!/usr/bin/env python
"""Dumps the hot loop of a kernel from libtsqb200.so (cuobjdump -sass): the innermost loop with the most
instructions, its instruction histogram and the listing -- the evidence behind DESIGN.md's instructions per packed
cell.   usage: python tools/sass_inner_loop.py <mangled-name-substring> <cells-per-iteration> [block] > profiles/sass_<...>.txt
("block": the straight-line block with the most DPX instructions inside that loop, for loops with alternative bodies)"""
import re
import subprocess
import sys
from collections import Counter

so = "tweakseq_b200/libtsqb200.so"
want, cells = sys.argv[1], float(sys.argv[2])
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, lines, name = False, [], None
for l in txt.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = want in m.group(1)
        if fn:
            name = m.group(1)
        continue
    if fn:
        lines.append(l)
addr = {}
for l in lines:
    m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        addr[int(m.group(1), 16)] = m.group(2).strip()
loops = []
for a, t in addr.items():
    m = re.search(r"BRA\s+(?:U[!A-Z0-9]*,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
# innermost loops: no other loop strictly inside
inner = [(lo, hi) for lo, hi in loops if not any(l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops)]
lo, hi = max(inner, key=lambda r: sum(1 for x in addr if r[0] <= x <= r[1]))
body = [(x, addr[x]) for x in sorted(addr) if lo <= x <= hi]
if len(sys.argv) > 3 and sys.argv[3] == "block":
    # the loop holds several alternative bodies (common case + tails): take the straight-line block with the most DPX
    targets = set()
    for _, t in body:
        m = re.search(r"(?:BRA|BSSY\S*)\s+(?:[UB]\S*,\s*)?0x([0-9a-f]+)", t)
        if m:
            targets.add(int(m.group(1), 16))
    blocks, cur = [], []
    for x, t in body:
        if x in targets and cur:
            blocks.append(cur)
            cur = []
        cur.append((x, t))
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        if op.split(".")[0] in ("BRA", "EXIT", "BSYNC", "WARPSYNC", "ENDCOLLECTIVE"):
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    body = max(blocks, key=lambda b: sum(1 for _, t in b if "VIMNMX3" in t))
    lo, hi = body[0][0], body[-1][0]
ops = Counter()
for _, t in body:
    p = t.split()
    ops[p[1] if p[0].startswith("@") else p[0]] += 1
print(f"kernel   : {name}")
print(f"hot loop : 0x{lo:x} .. 0x{hi:x}, {len(body)} instructions, {cells:g} packed cells per iteration -> {len(body) / cells:.2f} instructions per packed cell")
print("histogram:")
for k, v in ops.most_common():
    print(f"  {v:5d}  {k}")
print("listing  :")
for x, t in body:
    print(f"  /*{x:04x}*/ {t}")
