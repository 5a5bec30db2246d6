"""Attribute the warp-state samples of an `ncu --set full --import-source on` capture of the update kernel to
source lines of csrc/spo_update.cu and to the phases between PHASE_MARKs.

  ncu -i REPORT.ncu-rep --page source --csv > src.csv
  cd /tmp && cuobjdump -xelf all .../obj/spo_update.o && nvdisasm -gi -c *spo_update*.cubin > lines.txt
  python tools/ncu_lines.py src.csv lines.txt safe-policy-optimization_b200/csrc/spo_update.cu STEPS [ILi1E]
"""
import collections
import csv
import re
import sys

src_csv, lines_txt, cu, steps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
inst = sys.argv[5] if len(sys.argv) > 5 else "ILi1E"
inner = len(sys.argv) > 6 and sys.argv[6] == "inner"   # attribute to the innermost line of the inline chain (default: outermost)
src = open(cu).read().split("\n")
marks = []
for i, l in enumerate(src):
    m = re.search(r"PHASE_MARK\((\d+)\);\s*(//\s*(.*))?", l)
    if m and "define" not in l:
        marks.append((i + 1, int(m.group(1)), (m.group(3) or "").strip()))
loop0 = next(i for i, l in enumerate(src) if "for (int64_t qt = 0" in l) + 1


def phase_of(ln):
    if ln < loop0:
        return "(before the loop / hoisted)"
    for mline, idx, txt in marks:
        if ln <= mline:
            return f"{idx:2d} {txt}"
    return "(after the loop)"


lines = open(lines_txt).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and "spo_update_kernel" + inst in l][0]
offs, chain, last_file = {}, [], False
for l in lines[start + 1:]:
    if l.startswith(".text.") and inst not in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if not last_file:
            chain = []
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        last_file = True
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        offs[int(m.group(1), 16)] = list(chain)
        last_file = False
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ci, ie = h.index("# Samples"), h.index("Instructions Executed")
base = int(rows[2][0], 16)
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
by_phase, by_line = collections.OrderedDict(), collections.Counter()
line_ex = collections.Counter()
tot = 0
for r in rows[2:]:
    ch = offs.get(int(r[0], 16) - base, [])
    outer = [c for c in ch if c[0] == "spo_update.cu"]
    ln = (outer[0][1] if inner else outer[-1][1]) if outer else 0
    ph = phase_of(outer[-1][1]) if outer else "(other files)"
    a = by_phase.setdefault(ph, [0, 0, collections.Counter()])
    n = int(r[ci])
    a[0] += n
    a[1] += int(r[ie])
    by_line[ln] += n
    line_ex[ln] += int(r[ie])
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][h[i][6:]] += v
    tot += n
print(f"total samples {tot}\n")
print("| phase (ends at PHASE_MARK) | samples | % | warp-instr per CTA-step | top stall reasons |\n|---|---:|---:|---:|---|")
for ph, (n, ex, st) in sorted(by_phase.items(), key=lambda kv: kv[0]):
    print(f"| {ph} | {n} | {100 * n / tot:.1f} | {ex / steps / 12:.0f} | " + ", ".join(f"{k} {v}" for k, v in st.most_common(4)) + " |")
print("\ntop source lines:")
for ln, n in by_line.most_common(40):
    print(f"{n:6d} {100 * n / tot:5.1f}%  ex/step/cta {line_ex[ln] / steps / 12:7.0f}  L{ln}: {src[ln - 1].strip()[:120] if ln else ''}")
