"""Aggregate ncu stall samples / instructions by source-line regions of planes.cuh.  usage: ncu_regions.py rep"""
import csv, subprocess, sys, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# multiple files appear; track current file
cur = None; data = {}
hdr = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr and r and r[0].isdigit():
        try:
            data.setdefault(cur, []).append((int(r[0]), int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]), r[1]))
        except (ValueError, IndexError):
            pass
for f, lines in data.items():
    tot = sum(l[1] for l in lines); toti = sum(l[2] for l in lines)
    print(f, "samples", tot, "instr", toti)
src = open("/root/repo/sdim_b200/csrc/planes.cuh").read().splitlines()
def find(pat):
    for i, l in enumerate(src):
        if pat in l: return i + 1
    return None
marks = [("helpers/Geo", 1), ("gates", find("// ---- gates:")), ("noise_event", find("p_noise_event(const KParams")),
         ("measure:prologue+pivot", find("__device__ uint32_t p_measure(")), ("measure:random gather", find("// ---- random branch")),
         ("measure:update", find("// col_i += f_i * col_p")), ("measure:phase", find("// phase_i += f_i*ps")),
         ("measure:fixup", find("// destabilizer p <- old pivot")), ("measure:det", find("// ---- deterministic branch")),
         ("kernel:setup", find("interp_planes_kernel(const")), ("kernel:fetch", find("for (int64_t i0 = 0; i0 < p.n_ops")),
         ("kernel:dispatch", find("while (todo) {")), ("kernel:end", find("if (p.flags & SDIMB_WRITEBACK) {      // fold"))]
marks = [(n, l) for n, l in marks if l]
for f, lines in data.items():
    if not f.endswith("planes.cuh"): continue
    tot = sum(l[1] for l in lines) or 1; toti = sum(l[2] for l in lines) or 1
    agg = {}
    for ln, s, i, _ in lines:
        name = [n for n, l in marks if l <= ln][-1]
        a = agg.setdefault(name, [0, 0]); a[0] += s; a[1] += i
    for n, _ in marks:
        if n in agg: print(f"{n:28s} samples {100*agg[n][0]/tot:5.1f}%  instr {100*agg[n][1]/toti:5.1f}%")
