#!/bin/bash
# Builds an A/B variant of the library: tools/build_variant.sh <name> [extra nvcc flags...]  ->  mustafar_b200/libmustafar_b200_<name>.so
set -e
NAME=$1; shift
cd "$(dirname "$0")/../mustafar_b200/csrc"
mkdir -p build
OBJS=""
for f in runtime prune_compress spmv_compat decode_attn; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr "$@" \
      -c $f.cu -o build/${NAME}_$f.o 2> build/${NAME}_$f.ptxas.log || { cat build/${NAME}_$f.ptxas.log; exit 1; } &
  OBJS="$OBJS build/${NAME}_$f.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmustafar_b200_${NAME}.so $OBJS -lcudart
grep -A3 "sparse_decode_attn_kernelILi4ELi1" build/${NAME}_decode_attn.ptxas.log | grep -E "Used|spill" | sed 's/ptxas info    ://'
