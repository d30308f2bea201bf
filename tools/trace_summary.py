"""Print the head of a tools/trace_step.py JSON (kernel counts, span, top kernels by end-to-end delta and by duration)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("kernels", "span_us", "busy_us", "idle_us", "median_gap_us", "p90_gap_us")})
print("end-to-end deltas (critical path view)")
for k, v in list(d["end_to_end_delta"].items())[:8]:
    print("  ", k, v)
print("per kernel durations")
for k, v in list(d["per_kernel"].items())[:8]:
    print("  ", k, v)
