#!/bin/bash
# Collects the round's ncu evidence in ONE gpurun call on ONE GPU: tools/profile_all.sh <tag>
# Every ncu pass runs a command that has just exited 0 without ncu.  Outputs under gpurun_out/<tag>_*.
TAG=$1
set -x
export MFB200_BENCH_NO_MODEL=1
# 1. launch list of a short bench run (per-launch durations: cold-cache and serialised under ncu; shares matter)
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:compress|sparse_decode|window_append|prune|formulation|peer_wait|lengths_add" -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
# 2. --set full captures of the decode kernel at the BASELINE shapes
for c in cfg5 cfg3 cfg1 mid1 cfg4s; do
  timeout 300 python tools/prof_attn.py $c 1 > gpurun_out/${TAG}_prof_plain_$c.log 2>&1 || exit 1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sparse_decode_attn -s $([ "$c" = cfg1 ] && echo 4 || echo 3) -c 1 -f \
      -o gpurun_out/${TAG}_attn_$c python tools/prof_attn.py $c 1 > gpurun_out/${TAG}_ncu_$c.log 2>&1
done
# 3. the compression kernels
timeout 300 python tools/prof_compress.py > gpurun_out/${TAG}_prof_plain_compress.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_prefill -s 2 -c 1 -f -o gpurun_out/${TAG}_compress_prefill \
    python tools/prof_compress.py > gpurun_out/${TAG}_ncu_compress_prefill.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_append_chunk -c 1 -f -o gpurun_out/${TAG}_compress_append \
    python tools/prof_compress.py > gpurun_out/${TAG}_ncu_compress_append.log 2>&1
ls -la gpurun_out/${TAG}_*
