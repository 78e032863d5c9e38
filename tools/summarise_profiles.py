"""Turn gpurun_out/<round>_* captures into the committed summaries under profiles/ (run on the CPU box)."""
import csv, json, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout

# 1. launch list: keep our kernels + a per-kernel share table
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_launches.csv"), errors="ignore")) if len(r) > 14 and r[0].isdigit()]
total = sum(float(r[14]) for r in rows)
with open(os.path.join(P, f"{R}_launches.csv"), "w") as fh:
    fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --shots 16384 --no-cpu\n")
    fh.write("id,kernel,block,grid,duration_ns,share_of_all_launches\n")
    for r in rows:
        name = r[4].split("(")[0][-60:]
        fh.write(f"{r[0]},{name},{r[7]},{r[8]},{r[14]},{float(r[14]) / total:.4f}\n")

# 2. full captures -> text summaries + traffic.json
want = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
traffic = {}
for tag, shots in (("headline_planes", 16384), ("headline_front", 16384), ("headline_tail", 16384), ("headline_lanes_global", 4096),
                   ("gates_stream", 9472), ("config5_cluster", 1), ("config3_tile", 0), ("config4_tile", 0)):
    rep = os.path.join(G, f"{R}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = list(csv.reader(run("ncu", "-i", rep, "--page", "raw", "--csv").splitlines()))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = run(sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25")
    with open(os.path.join(P, f"{R}_ncu_{tag}.txt"), "w") as fh:
        what = {"gates_stream": "the 2000 gates of the headline circuit (no noise, no measurement), uint8 HBM store",
                "headline_front": "the headline workload, first kernel of the step (gates + noise: everything in front of the trailing measurement run)",
                "headline_tail": "the headline workload, second kernel of the step (the 256 trailing measurements on the generator-major image)",
                "config5_cluster": "config 5 (random Clifford d = 5, n = 4096, 8192 gates + 4096 measurements, one 64 MiB tableau) on a 16-CTA cluster"
                }.get(tag, "the headline workload")
        fh.write(f"# ncu --set full --clock-control none --import-source on, one launch of {what}, {shots} shots\n")
        fh.write(f"kernel: {m.get('Kernel Name', ('?',))[0]}\n")
        for k in want:
            if k in m:
                fh.write(f"{k:68s} {m[k][0]} {m[k][1]}\n")
        fh.write("\n# hottest source lines (share of stall samples, share of executed warp instructions)\n" + lines)
    def to_bytes(key):
        v, u = m[key]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        return float(v) * scale
    def num(key):
        return float(m[key][0].replace(",", "")) if key in m else None
    dur = float(m["gpu__time_duration.sum"][0])
    dur_ms = dur * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(m["gpu__time_duration.sum"][1], 1.0)
    traffic[tag] = {"shots": shots, "kernel": m.get("Kernel Name", ("?",))[0].split("(")[0].split("::")[-1],
                    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                    "duration_ms": dur_ms, "inst_executed_per_launch": num("smsp__inst_executed.sum"),
                    "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "l2_hit_pct": num("lts__t_sector_hit_rate.pct"), "l1_hit_pct": num("l1tex__t_sector_hit_rate.pct"),
                    "registers": num("launch__registers_per_thread"), "capture": R}
# the headline step is two kernels since round 2 (interpreter in front, tail run behind): the dominant one leads
two = [traffic[k] for k in ("headline_front", "headline_tail") if k in traffic]
if two:
    old = {}
    if os.path.exists(os.path.join(P, "traffic.json")):
        old = json.load(open(os.path.join(P, "traffic.json"))).get("all", {})
    old.update(traffic)
    dom = max(two, key=lambda t: t["duration_ms"])
    t = dict(dom); t["all"] = old
    t["step_kernels"] = {k: traffic[k] for k in ("headline_front", "headline_tail") if k in traffic}
    json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
elif "headline_planes" in traffic:
    t = dict(traffic["headline_planes"]); t["all"] = traffic
    json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for name in (f"{R}_sanitizer.txt", f"{R}_bench_n1.json", f"{R}_bench_reference.json", f"{R}_bench_n2.json",
             f"{R}_configs.json", f"{R}_config5_breakdown.txt"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, name))
print(os.listdir(P))
