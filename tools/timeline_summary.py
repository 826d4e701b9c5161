#!/usr/bin/env python
"""Where does the replayed step spend its wall time?  Reads the chrome trace written by `bench.py --timeline` (CUPTI
kernel records of the CUDA graph: start, duration, stream) and prints, for the LAST step in the trace:
  * wall time, summed kernel time, average number of kernels in flight;
  * the time with exactly 0 / 1 / 2 / >= 3 kernels running;
  * the kernels that run ALONE (nothing else on the GPU): that is the step's critical path, grouped by name;
  * the longest gaps with nothing running.
Usage: python tools/timeline_summary.py gpurun_out/timeline.json [--csv out.csv]"""
import collections
import json
import re
import sys


def main():
    path = sys.argv[1]
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    step = ev[len(ev) // 2:]          # the trace holds two replays of the same graph: keep the second
    t0 = step[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in step)
    wall = t1 - t0
    busy = sum(e["dur"] for e in step)
    print(f"{len(step)} kernels, wall {wall:.1f} us, kernel time {busy:.1f} us, mean concurrency {busy / wall:.2f}, "
          f"{len(set(e['args'].get('stream') for e in step))} streams")
    pts = []
    for i, e in enumerate(step):
        pts.append((e["ts"], 1, i))
        pts.append((e["ts"] + e["dur"], -1, i))
    pts.sort(key=lambda p: (p[0], p[1]))
    level = collections.Counter()
    alone = collections.Counter()
    alone_n = collections.Counter()
    active = set()
    last = t0
    idle = []
    short = lambda n: re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "").replace("void ", ""))[:60]
    for t, d, i in pts:
        dt = t - last
        if dt > 0:
            level[min(len(active), 3)] += dt
            if len(active) == 1:
                k = short(step[next(iter(active))]["name"])
                alone[k] += dt
            if not active:
                idle.append((dt, last - t0))
        if d > 0:
            active.add(i)
        else:
            active.discard(i)
        last = t
    for e in step:
        alone_n[short(e["name"])] += 1
    print("time with n kernels running: " + ", ".join(f"{n if n < 3 else '>=3'}: {level[n]:.0f} us ({100 * level[n] / wall:.0f} %)"
                                                      for n in range(4)))
    print("running alone (critical path), top 25:")
    for k, v in alone.most_common(25):
        print(f"  {v:8.1f} us  ({alone_n[k]:3d} launches)  {k}")
    print("longest idle gaps (us, at offset):", [(round(a, 1), round(b)) for a, b in sorted(idle, reverse=True)[:8]])
    if "--csv" in sys.argv:
        out = sys.argv[sys.argv.index("--csv") + 1]
        with open(out, "w") as f:
            f.write("start_us,dur_us,stream,kernel\n")
            for e in step:
                f.write(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{e['args'].get('stream')},\"{short(e['name'])}\"\n")


if __name__ == "__main__":
    main()
