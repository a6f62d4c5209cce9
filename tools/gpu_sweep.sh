#!/bin/bash
# usage: tools/gpu_sweep.sh <tag> "<batch sizes>" [kernel-regex-for-full-capture]
tag=$1; kre=$3
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
for b in $2; do
  python bench.py --steps 10 --warmup 3 --batch $b --no-cpu-baseline > gpurun_out/bench_${tag}_b$b.json 2>> gpurun_out/bench_$tag.err; echo "batch $b rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_${tag}_b$b.json"))
print("batch",$b,"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms/step",round(d["ms_per_step"],3),d["config"].get("pair_stats"))
print({k:round(v["ms_total"]/d["steps"],3) for k,v in d["stages"].items()})
PY
done
if [ -n "$kre" ]; then
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$kre" -s 54 -c 18 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_f_$tag.log 2>&1
fi
