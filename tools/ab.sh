#!/bin/bash
# A/B two builds of the library inside ONE gpurun call (same box, same clocks): tools/ab.sh <libA.so> <libB.so>
for rep in 1 2; do
  for lib in "$1" "$2"; do
    export MFB200_LIB=$PWD/$lib
    echo "== $lib (rep $rep)"
    for c in cfg1 cfg3 cfg5s cfg4s mid1 mid2; do python tools/prof_attn.py $c 3 | sed 's/n_split.*cold//'; done
    python bench.py --steps 32 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench us/launch', round(d['roofline']['us_per_launch'],2), 'us/layer-step', round(d['us_per_layer_step'],2))"
  done
done
