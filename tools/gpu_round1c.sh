#!/bin/bash
#  Second (last) gpurun call of the round: pick the fastest sweep kernel shape and measure with it.
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round1c.log) 2>&1
date
echo "== a. stage tests, all sweep shapes";  date +%T
: > $OUT/shapes_ok.txt
for V in 1:sweep_v1 2:sweep_128x8 3:sweep_256x6 4:sweep_256x7; do
  if timeout 100 python -m pytest tests/test_gpu_stages.py -x -q -k "sort and ${V#*:}" > $OUT/stage_${V%%:*}.log 2>&1; then echo ${V%%:*} >> $OUT/shapes_ok.txt; fi
  echo "shape ${V}: $(tail -1 $OUT/stage_${V%%:*}.log)"
done
echo "== b. A/B of the sweep shapes, 8 Gbp in 5 passes";  date +%T
timeout 150 python tools/ab_bench.py 8 5 sortonly 2>&1 | tee $OUT/ab_8gbp_shapes.jsonl | cut -c1-420
WIN=$(python - <<'PY'
import json
ok = set(int(x) for x in open("gpurun_out/shapes_ok.txt").read().split())
best = None
for l in open("gpurun_out/ab_8gbp_shapes.jsonl"):
    try: d = json.loads(l)
    except Exception: continue
    if d["sort"] not in ok: continue
    t = d["ms"]["msSort"] + d["ms"]["msSortHist"]
    if best is None or t < best[0]: best = (t, d["sort"])
print(best[1] if best else 1)
PY
)
echo "fastest sweep shape: $WIN"
echo "== c. whole-path tests with MB_SORT=$WIN";  date +%T
MB_SORT=$WIN timeout 150 python -m pytest tests/test_gpu_count.py -x -q 2>&1 | tail -3
echo "== d. bench 100 Gbp with MB_SORT=$WIN";  date +%T
MB_SORT=$WIN timeout 300 python bench.py --steps 2 --warmup 1 > $OUT/bench_100gbp_win.json 2> $OUT/bench_100gbp_win.err; cut -c1-1800 $OUT/bench_100gbp_win.json; tail -3 $OUT/bench_100gbp_win.err
echo "== e. union-sum next to the reference";  date +%T
timeout 150 python tools/union_bench.py 0.25 2>&1 | tee $OUT/union_bench.json | cut -c1-900
date
