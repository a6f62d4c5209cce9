#!/bin/bash
# One GPU-box round: parity tests, headline bench (both arms), ncu launch list + full capture of every kernel.
# usage: tools/gpu_round.sh <tag> [nofull]
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_$tag.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2>> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$tag.err
CMD="python bench.py --steps 2 --warmup 3 --batch 50 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_l_$tag.log 2>&1
if [ -z "$2" ]; then
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'^k_' -s 120 -c 26 -f -o gpurun_out/prof_$tag $CMD > gpurun_out/ncu_f_$tag.log 2>&1
fi
python tools/time_configs.py > gpurun_out/other_configs_$tag.json 2>> gpurun_out/bench_$tag.err; echo "configs rc=$?"
for ing in grey bgr; do python bench.py --steps 10 --warmup 3 --ingest $ing > gpurun_out/bench_ingest_${ing}_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ingest $ing rc=$?"; done
cat gpurun_out/bench_$tag.json
