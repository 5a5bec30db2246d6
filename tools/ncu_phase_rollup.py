"""Roll the warp-state samples of an `ncu --set full --import-source on` capture of spo_update_kernel<1>
up to the phases of the minibatch step.

  ncu -i REPORT.ncu-rep --page source --csv > src.csv          (SASS-level samples)
  cuobjdump -xelf all spo_update.o; nvdisasm -gi -c *.cubin > lines.txt   (SASS -> source line, with inlining)
  python tools/ncu_phase_rollup.py src.csv lines.txt safe-policy-optimization_b200/csrc/spo_update.cu STEPS

Every SASS instruction is attributed to the outermost spo_update.cu line of its inline chain; a line belongs
to the phase that ends at the next PHASE_MARK / TRACE_MARK after it.  Samples of the barrier-only 4th CTA
(all of them on the cluster-barrier wait) are removed from the percentages."""
import collections
import csv
import re
import sys

src_csv, lines_txt, cu, steps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
NAMES = {22: "top: stage-in", 15: "top: barrier", 0: "top: log-std constants", 11: "hidden GEMM (x2)", 1: "forward layer 1: GEMM + tanh epilogue + barrier", 2: "forward layer 2: GEMM + tanh epilogue + barrier",
         17: "output layer", 18: "Adam scalars + loss rows", 3: "loss barrier + sums", 19: "small grads (dW3, db3, dlog_std)", 4: "dz2 + barrier", 12: "dW2 GEMM",
         13: "db2 column sums", 14: "dh1 GEMM", 5: "dh1 epilogue + barrier", 6: "dW1 GEMM + db1", 7: "cross-GPU exchange (off)", 20: "regulariser + sum of squares",
         8: "block reduce", 16: "cluster arrive + request next tile", 9: "cluster wait", 21: "DSMEM read + clip + Adam (W1, W2)", 10: "Adam (small params)"}
src = open(cu).read().split("\n")
loop0 = next(i for i, l in enumerate(src) if "for (int64_t q = 0" in l) + 1
marks = []
for i, l in enumerate(src):
    m = re.search(r"(PHASE|TRACE)_MARK\((\d+)\)", l)
    if m and i + 1 > loop0 and "define" not in l:
        marks.append((i + 1, int(m.group(2))))
hidden_line = next(i for i, l in enumerate(src) if "PHASE_MARK(11)" in l) + 1
loop_end = marks[-1][0]


def phase_of(ln):
    if ln < loop0 and not (hidden_line - 12 <= ln <= hidden_line + 12):
        return "(declarations: hoisted address math etc.)"
    if hidden_line - 12 <= ln <= hidden_line:
        return NAMES[11]
    if hidden_line < ln <= hidden_line + 12:
        return "L1/L2 epilogue (tanh)"
    for mline, idx in marks:
        if ln <= mline:
            return NAMES.get(idx, str(idx))
    return "(after the loop)"


lines = open(lines_txt).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and "spo_update_kernelILi1E" in l][0]
offs, chain, last_file = {}, [], False
for l in lines[start + 1:]:
    if l.startswith(".text.") and "ILi1E" not in l:
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
agg = collections.OrderedDict()
tot = idle = 0
for r in rows[2:]:
    ch = offs.get(int(r[0], 16) - base, [])
    outer = [c for c in ch if c[0] == "spo_update.cu"]
    ph = phase_of(outer[-1][1]) if outer else "(other files)"
    a = agg.setdefault(ph, [0, 0, collections.Counter()])
    n = int(r[ci])
    a[0] += n
    a[1] += int(r[ie])
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[2][h[i][6:]] += v
    tot += n
wait = agg.get(NAMES[9])
idle = min(wait[0], tot // 4) if wait else 0      # the 4th CTA spends the whole launch there
act = tot - idle
print(f"total samples {tot}, of which barrier-only CTA ~{idle}; percentages are of the remaining {act}\n")
print("| phase | samples | % of active | warp-instr per CTA-step | top stall reasons |\n|---|---:|---:|---:|---|")
for ph, (n, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    n2 = n - idle if ph == NAMES[9] else n
    print(f"| {ph} | {n2} | {100 * n2 / act:.1f} | {ex / steps / 3:.0f} | " + ", ".join(f"{k} {v}" for k, v in st.most_common(3)) + " |")
