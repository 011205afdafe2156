#!/bin/sh
# A/B of the ALLL_TUNE measurement knobs (alll_device.cuh: TUNE_*) on one B200: per-phase traces and solve times.
# usage: tools/tune_ab.sh <workload> <tag> <tune values...>
W=$1; TAG=$2; shift 2
O=gpurun_out
mkdir -p $O
for T in "$@"; do
    env ALLL_TUNE=$T ${TRACE:+ALLL_TRACE=1} python tools/prof_sweep.py --workload $W --reps 10 --solves 4 ${MAXR:+--max-rounds $MAXR} > $O/${TAG}_${W}_t$T.json 2> $O/${TAG}_${W}_t$T.txt
    python - $O/${TAG}_${W}_t$T.json $T <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
s = [d[k] for k in d if k.startswith("solve")]
print("tune", sys.argv[2], "sweep_ms %.4f" % d["sweep_ms"], "solve_ms", ["%.3f" % x["ms"] for x in s], "between", ["%.3f" % x["between_ms"] for x in s],
      "sweep_in_solve", ["%.3f" % x["sweep_ms"] for x in s], "iters", [x["iters"] for x in s])
PY
done
