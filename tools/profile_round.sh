#!/bin/bash
# Refreshes one piece of ncu evidence under gpurun_out/ (run through gpurun on ONE GPU; one ncu pass per call, only
# after the same command passed without ncu):
#   tools/profile_round.sh <tag> launches        launch list of a short bench run (gpu__time_duration per launch)
#   tools/profile_round.sh <tag> cfg1|cfg3|mid1  one `--set full` capture of the fused decode kernel at that shape
TAG=$1; WHAT=$2
if [ "$WHAT" = launches ]; then
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.log 2>&1 || exit 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
else
  timeout 200 python tools/prof_attn.py $WHAT 1 > gpurun_out/${TAG}_prof_plain_$WHAT.log 2>&1 || exit 1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sparse_decode_attn -s $([ "$WHAT" = cfg1 ] && echo 4 || echo 3) -c 1 -f -o gpurun_out/${TAG}_attn_$WHAT \
      python tools/prof_attn.py $WHAT 1 > gpurun_out/${TAG}_ncu_$WHAT.log 2>&1
fi
ls -la gpurun_out/${TAG}_* | tail -5
