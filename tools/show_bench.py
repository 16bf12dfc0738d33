"""Print the headline fields of bench.py JSON lines.  usage: python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    lines = [l for l in open(path).read().strip().split("\n") if l.startswith("{")]
    if not lines:
        print(path, "no JSON line")
        continue
    j = json.loads(lines[-1])
    pr = (j.get("per_rank") or [{}])[0]
    print(path, "windows/s", round(j["value"], 2), "ms/step", round(j["ms_per_step"], 1), "e2e", round(j["e2e"]["value"], 2),
          "clocks", j.get("clocks"), "settle", pr.get("settle"), "step_ms", pr.get("step_ms"))
