#!/usr/bin/env python
"""Regenerate the round-2 table of profiles/bench_history.md from the raw bench.py lines kept in profiles/bench_lines/.
    python tools/make_history.py C100k=bench_r02i C1=bench_r02f_C1 ...        (name=file pairs; prints markdown)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(f):
    return json.loads(open(os.path.join(ROOT, "profiles", "bench_lines", f + ".json")).read().strip().splitlines()[-1])


def main():
    pairs = [a.split("=") for a in sys.argv[1:]]
    print("| config | raw / down-sampled pts | GPU value scans/s (ms/step) | e2e | e2e_raw | LIO / mesh ms (blocking) | CPU all-cores pipelined (serial) | CPU 4-thread pipelined (serial) | value ÷ CPU pipelined | dominant kernel: frac of HBM peak | whole step frac |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    for name, f in pairs:
        d = load(f)
        c = d.get("cpu_baseline") or {}
        a, r = c.get("all_cores", {}), c.get("reference_4_threads", {})
        raw = (d.get("e2e_raw") or {}).get("value")
        cpu = a.get("pipelined_scans_s")
        print(f"| {name} (`{f}`) | {d['config']['points_per_scan_raw']} / {d['config']['points_per_scan_downsampled']} | **{d['value']:.0f}** ({d['ms_per_step']:.3f}) | {d['e2e']['value']:.0f} | "
              f"{raw if raw is None else round(raw)} | {d['stage_ms']['lio_total']:.3f} / {d['stage_ms']['mesh_total']:.3f} | {cpu} ({a.get('serial_scans_s')}) | "
              f"{r.get('pipelined_scans_s')} ({r.get('serial_scans_s')}) | {'' if not cpu else round(d['value'] / cpu)}× | {d['roofline']['kernel']}: {d['roofline']['frac']:.4f} | {d['roofline_whole_step']['frac']:.4f} |")


if __name__ == "__main__":
    main()
