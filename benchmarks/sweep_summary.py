#!/usr/bin/env python
"""One line per bench.py output file: tick ms, value, e2e, stage times, host enqueue."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d["stage_ms"]
        print("%-40s tick %.4f  %.3f M  e2e %.3f M (%.4f)  gal %.4f match %.4f  host %.4f thr %.4f" % (
            f.split("/")[-1], d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"],
            st["cosine"], st["match"], d["host_enqueue_ms_per_step"], d.get("host_throttled_ms_per_step", 0)))
    except Exception as e:
        print(f, "ERR", e)
