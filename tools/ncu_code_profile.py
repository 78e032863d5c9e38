"""Executed-instruction profile of the first kernel in an ncu report (needs --import-source on, -lineinfo build):
stall reasons, opcode mix, code footprint in 128-byte lines, top source lines by executed warp instructions.
usage: python tools/ncu_code_profile.py report.ncu-rep [top_lines]"""
import collections
import csv
import subprocess
import sys


def page(rep, what):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what],
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
sass = page(rep, "sass")
hdr = next(r for r in sass if r and r[0] == "Address")
col = {name: i for i, name in enumerate(hdr)}
rows = [r for r in sass[sass.index(hdr) + 1:] if len(r) > 10]
ins = [int(r[col["Instructions Executed"]]) for r in rows]
smp = [int(r[col["# Samples"]]) for r in rows]
tot, tots = sum(ins) or 1, sum(smp) or 1
print(f"{len(rows)} SASS instructions, {tot} executed warp instructions, {tots} stall samples")
reasons = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[col[n]]) for r in rows) for n in reasons}
print("stall reasons:", ", ".join(f"{n[6:]} {100 * v / tots:.1f}%" for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
ops = collections.Counter()
for r, i in zip(rows, ins):
    w = r[col["Source"]].split()
    ops[(w[1] if w[0].startswith("@") else w[0]).split(".")[0]] += i
print("opcode mix:", ", ".join(f"{o} {100 * v / tot:.1f}%" for o, v in ops.most_common(14)))
lines = [sum(ins[i:i + 8]) for i in range(0, len(ins), 8)]
for thr in (0.0, 1e-5, 1e-4, 1e-3):
    hot = [v for v in lines if v / tot > thr]
    print(f"code lines (128 B) with > {thr:g} of the executed instructions: {len(hot)} of {len(lines)} = "
          f"{len(hot) * 128 / 1024:.1f} KB, covering {100 * sum(hot) / tot:.1f}%")
addr2line, src, cur = {}, {}, (None, None)
cur_file = None
for r in page(rep, "cuda,sass"):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0].isdigit():
        cur = (cur_file, int(r[0]))
        src[cur] = r[1].strip()
    elif r[0] == "" and len(r) > 2 and r[2].startswith("0x"):
        addr2line[r[2]] = cur
per = collections.defaultdict(lambda: [0, 0, 0])
for r, i, s in zip(rows, ins, smp):
    a = per[addr2line.get(r[col["Address"]], ("?", 0))]
    a[0] += i
    a[1] += s
    a[2] += 1
print("top source lines by executed instructions (share of instructions, share of samples, SASS sites):")
for fl, (i, s, n) in sorted(per.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100 * i / tot:5.2f}% ins {100 * s / tots:5.2f}% smp {n:4d}  {fl[0]}:{fl[1]:<4d} {src.get(fl, '')[:90]}")
