"""Condense bench.py's JSON line (stdin) to one readable line; errors pass through."""
import json, sys
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        r = d["roofline"]
        msg = "  value %.2f G/s  ms/step %.3f  frac %.4f  avg_launch %.3f  share %.3f  launches %d" % (
            d["value"] / 1e9, d["ms_per_step"], r["frac"], r["avg_launch_ms"], r["share_of_step"], d["gpu_launches"])
        if d.get("e2e"):
            e = d["e2e"]
            msg += "  e2e %.2f G/s (%.3f ms/step, h2d %d, d2h %d)" % (e["value"] / 1e9, e["ms_per_step"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"])
        print(msg)
    elif "rror" in line or "Traceback" in line:
        print(line.rstrip())
