#!/bin/sh
# A/B of alll_config.flags values on one B200: stand-alone sweep time and four solves per value.
# usage: tools/flags_ab.sh <workload> <tag> <flags values...>
W=$1; TAG=$2; shift 2
O=gpurun_out
mkdir -p $O
for F in "$@"; do
    python tools/prof_sweep.py --workload $W --reps 20 --solves 4 --flags $F > $O/${TAG}_${W}_f$F.json 2> $O/${TAG}_${W}_f$F.err
    python - $O/${TAG}_${W}_f$F.json $F <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
s = [d[k] for k in d if k.startswith("solve")]
print("flags", sys.argv[2], "sweep_ms %.4f" % d["sweep_ms"], "GB/s %.0f" % d["achieved_GBps"], "solve", ["%.3f" % x["ms"] for x in s],
      "sweeps in solve", ["%.3f" % x["sweep_ms"] for x in s], "between", ["%.3f" % x["between_ms"] for x in s], [x["iters"] for x in s])
PY
done
