"""Per-stream critical-path summary of a tools/timeline.py dump: for every kernel name, the mean time between the END of the previous
kernel on the same stream and its own END (= what the kernel adds to that chain's critical path when PDL overlaps prologues)."""
import collections, json, sys
for path in sys.argv[1:]:
    d = json.load(open(path))
    ev = d["events"]
    print("==", path, "kernels", d["total_kernels"], "span_us", round(d["span_us"], 1))
    bys = collections.defaultdict(list)
    for e in ev:
        bys[e[1]].append(e)
    for s, l in sorted(bys.items()):
        l.sort(key=lambda e: e[3])
        cp = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for p, c in zip(l[:-1], l[1:]):
            cp[c[0]][0] += 1
            cp[c[0]][1] += c[3] - p[3]
            cp[c[0]][2] += c[3] - c[2]
        tot = sum(v[1] for v in cp.values())
        print(" stream", s, "window span", round(l[-1][3] - l[0][3], 1))
        for k, v in sorted(cp.items(), key=lambda kv: -kv[1][1]):
            print(f"   {k:44s} n={v[0]:4d} end_delta={v[1] / v[0]:6.2f} duration={v[2] / v[0]:6.2f} share={v[1] / tot:.3f}")
