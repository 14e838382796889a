#!/usr/bin/env python
"""Print the essentials of bench.py JSON lines: python tools/benchline.py file.json [...]"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "ERR", e); continue
    if d.get("value") is None:
        print(f, "NO VALUE", d.get("error"), d.get("parity_gate")); continue
    print(f"{f}: value {d['value']} e2e {d['e2e']['value']} ms/step {d['ms_per_step']} n_gpus {d['n_gpus']}")
    if "timing_detail" in d:
        td = d["timing_detail"]
        print("   host_enqueue_ms", td["host_enqueue_ms_per_scan"], "serial_blocking_ms", td["serial_ms_per_scan_blocking"], "graphs", td["cuda_graphs"])
    print("   stage", d.get("stage_ms"))
    print("   gate", (d.get("parity_gate") or {}).get("ok"), "roof", d.get("roofline"))
    print("   whole", d.get("roofline_whole_step"))
    k = d.get("kernel_ms_per_scan") or {}
    print("   kern", {a: b for a, b in list(k.items())[:14]})
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print("   cpu", c["value"], "all", c.get("all_cores"), "ref4", c.get("reference_4_threads"))
    for key in ("e2e_raw", "multi_stream", "sharded_single_stream"):
        if d.get(key):
            print("  ", key, d[key])
