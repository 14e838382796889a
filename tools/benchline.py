import json,sys
d=json.loads(sys.stdin.read()); print(sys.argv[1], d["config"].get("host_enqueue_ms_per_scan"), d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms"]["lio_total"], d["stage_ms"]["mesh_total"])
