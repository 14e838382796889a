"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).  Usage: python profiles/summarize.py <tag>"""
import collections, csv, json, subprocess, sys, os
tag = sys.argv[1]
out = {}
lp = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0].isdigit()]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        tot[name] += float(r[-1]); cnt[name] += 1
    T = sum(tot.values())
    out["launch_list"] = {"command": "ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)", "launches": len(rows), "total_us": round(T / 1e3, 1),
                          "kernels": {k: {"us": round(v / 1e3, 1), "n": cnt[k], "avg_us": round(v / cnt[k] / 1e3, 2), "share": round(v / T, 4)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}}
    os.makedirs("profiles", exist_ok=True)
    open(f"profiles/launches_{tag}.csv", "w").write(open(lp).read())
rp = f"gpurun_out/prof_{tag}.ncu-rep"
if os.path.exists(rp):
    raw = subprocess.check_output(["ncu", "-i", rp, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    units = {hdr[i]: rows[1][i] for i in idx}
    recs = [{hdr[i]: r[i] for i in idx} for r in rows[2:]]
    out["full_capture"] = {"command": "ncu --set full --clock-control none --import-source on", "units": units, "launches": recs}
json.dump(out, open(f"profiles/summary_{tag}.json", "w"), indent=1)
print("wrote", f"profiles/summary_{tag}.json")
