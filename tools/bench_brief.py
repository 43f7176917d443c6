"""Print the headline numbers of a bench.py JSON line read from stdin."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
c = d.get("config", {})
print("value %.1f  ms/step %.3f | e2e %.1f  ms/step %.3f | plan_ms %.3f | settle %s  step_ms min/med/max %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], (d.get("roofline") or {}).get("plan_ms", 0.0),
    c.get("settle_steps_untimed"), c.get("step_ms_min_median_max")), "| e2e steps", c.get("e2e_step_ms_min_median_max"))
