"""Summarise ncu captures from gpurun_out/ into small text files under profiles/ (committed)."""
import csv, json, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"
OUT.mkdir(exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]

def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        res.append(d)
    return res

traffic = {}
for rep in sorted((ROOT / "gpurun_out").glob("*.ncu-rep")):
    rows = raw(rep)
    with open(OUT / f"{tag}_ncu_{rep.stem.replace('prof_', '')}.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --cache-control none  ({rep.name}); one block per captured launch\n")
        for d in rows:
            f.write("\n" + d["kernel"][:150] + "\n")
            for k in KEYS:
                if k in d:
                    f.write(f"  {k:72s} {d[k]}\n")
    if rep.stem in ("prof_gram8192", "prof_gram16k") and rows:
        rows = sorted(rows, key=lambda d: -float(d["gpu__time_duration.sum"].split()[0].replace(",", "")))   # the top-level one
        def num(s):
            v, u = s.split()[0], s.split()[1] if len(s.split()) > 1 else ""
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            return float(v.replace(",", "")) * mult
        traffic["tc_gemm_gram_top_bytes"] = num(rows[0]["dram__bytes_read.sum"]) + num(rows[0]["dram__bytes_write.sum"])
        traffic["tc_gemm_gram_top_tensor_pipe_pct"] = rows[0]["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
if traffic:
    import datetime
    traffic["source"] = f"profiles/{tag}_ncu_gram16k.txt (ncu --set full, {datetime.date.today().isoformat()})"
    (OUT / "traffic.json").write_text(json.dumps(traffic, indent=1))
for name in ("launches_16k", "launches_1m", "launches_256k"):
    src = ROOT / "gpurun_out" / f"{name}.csv"
    if src.exists():
        agg = subprocess.run([sys.executable, str(ROOT / "scripts" / "agg_launches.py"), str(src), "2"], capture_output=True, text=True).stdout
        (OUT / f"{tag}_{name}_summary.txt").write_text(
            f"# ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none, python scripts/gpu_profile_run.py (2 runs, per-run numbers)\n" + agg)
print("wrote", sorted(p.name for p in OUT.iterdir()))
