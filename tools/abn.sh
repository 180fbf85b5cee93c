#!/bin/bash
# A/B several builds of the library inside ONE gpurun call: tools/abn.sh <which: gqa|mha|perf> <lib1.so> <lib2.so> ...
WHICH=$1; shift
for rep in 1 2; do
  for lib in "$@"; do
    export MFB200_LIB=$PWD/mustafar_b200/$lib
    echo "== $lib (rep $rep)"
    python tools/quick_attn.py $WHICH 2>&1 | sed 's/max .* repeat_equal=True//'
  done
done
