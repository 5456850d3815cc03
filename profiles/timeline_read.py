#!/usr/bin/env python
"""Summarise a torch.profiler chrome trace: per-stream busy time and per-kernel durations."""
import collections
import gzip
import json
import sys

tr = json.load(gzip.open(sys.argv[1]))
ev = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and e.get("ph") == "X"]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print("events %d, span %.3f ms" % (len(ev), (t1 - t0) / 1e3))
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream")].append(e)
for sid, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in es)
    names = collections.Counter(e["name"].split("<")[0].split("(")[0].replace("void ", "").replace("gccb::", "") for e in es)
    print("stream %s: %d kernels, busy %.3f ms, top: %s" % (sid, len(es), busy / 1e3, names.most_common(3)))
by = collections.defaultdict(list)
for e in ev:
    by[e["name"].split("(")[0].replace("void ", "").replace("gccb::", "")[:60]].append(e["dur"])
print("%-62s %5s %9s %9s %9s" % ("kernel", "n", "mean_us", "max_us", "total_ms"))
for k, v in sorted(by.items(), key=lambda kv: -sum(kv[1]))[:34]:
    print("%-62s %5d %9.1f %9.1f %9.3f" % (k, len(v), sum(v) / len(v), max(v), sum(v) / 1e3))

if len(sys.argv) > 2:
    # sequence view of one step on the busiest GIN stream: start offset, duration, gap to previous
    main_sid = max(streams, key=lambda k: sum(1 for e in streams[k] if "gin_" in e["name"] or "infonce" in e["name"]))
    es = streams[main_sid]
    # one step = between consecutive adam_ema kernels (any stream)
    adam = sorted(e["ts"] for e in ev if "adam_ema" in e["name"])
    lo, hi = adam[1], adam[2]
    print("step window %.3f ms; kernels of ALL streams inside it:" % ((hi - lo) / 1e3))
    inside = [e for e in ev if lo < e["ts"] <= hi and ("gin_" in e["name"] or "infonce" in e["name"] or "adam" in e["name"]
                                                        or "moco" in e["name"] or "gradnorm" in e["name"])]
    prev_end = {}
    for e in inside:
        sid = e["args"].get("stream")
        gap = e["ts"] - prev_end.get(sid, e["ts"])
        prev_end[sid] = e["ts"] + e["dur"]
        print("%8.1f us  s%-4s dur %7.1f gap %7.1f  %s" % (e["ts"] - lo, sid, e["dur"], gap,
              e["name"].split("(")[0].replace("void ", "").replace("gccb::", "")[:50]))
