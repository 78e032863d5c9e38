"""Summarise an ncu report per CUDA source line: samples, instructions executed, shared-memory conflicts.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = next(r for r in rows if r and r[0] == "Line No")
idx = {name: i for i, name in enumerate(hdr)}
si, ii, ci = idx["# Samples"], idx["Instructions Executed"], idx["L1 Wavefronts Shared Excessive"]
lines = []
for r in rows:
    if len(r) > si and r[0].isdigit():
        try:
            lines.append((int(r[si]), int(r[ii]), int(r[ci]) if r[ci].isdigit() else 0, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print(f"total samples {tot}, total warp instructions {toti}")
for s, i, c, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins  confl {c:>10d}  L{ln:<4d} {src}")
