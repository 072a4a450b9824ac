"""stdin: one bench.py JSON line -> one short line (value, ms/step, per-step min/median/max, e2e, aggregations, clocks)."""
import json, sys

line = sys.stdin.read().strip()
if not line:
    print("no JSON line")
    sys.exit(0)
d = json.loads(line)
ps = d.get("config", {}).get("per_step_ms") or {}
e = d.get("e2e") or {}
print("%s N=%s: %.0f %s, %.3f ms/step (per step min %s median %s max %s at %s), aggregations %s, e2e %s (%s ms/step, max %s), clocks %s" % (
    d.get("config", {}).get("algo"), d.get("n_gpus"), d["value"], d.get("unit"), d["ms_per_step"], ps.get("min"), ps.get("median"), ps.get("max"),
    ps.get("argmax_step"), d.get("config", {}).get("aggregations_in_window"), ("%.0f" % e["value"]) if e.get("value") else None,
    e.get("ms_per_step"), (e.get("per_step_ms") or {}).get("max"), d.get("clocks")))
