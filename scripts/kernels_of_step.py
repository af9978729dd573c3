#!/usr/bin/env python
"""Print every kernel of one bench step with its time (reads a bench.py JSON line from stdin)."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print("value %.0f %s  step %.3f ms  families %s" % (d["value"], d["unit"], d["ms_per_step"],
      {k: round(v, 3) for k, v in (d.get("kernel_family_ms_per_step") or {}).items()}))
for k in d["roofline"]["kernels"]:
    print("  %-28s x%d %8.3f ms  %5.1f %%" % (k["kernel"], k["launches_per_step"], k["ms_per_step"], 100 * k["share_of_kernel_time"]))
print("  kernel time / step time = %.3f" % d["roofline"]["kernel_time_over_step_time"])
if d.get("secondary"):
    print("  secondary:", {k: round(v["value"]) for k, v in d["secondary"].items()})
