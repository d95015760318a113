#!/bin/bash
# Ablation builds of libspmvb200.so (one -D each) into tools/variants/; select with SPMVB200_LIB=...
#   tools/build_variants.sh g4=-DSPMV_GATHER_MODE=4 g5=-DSPMV_GATHER_MODE=5
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    make -s -C "$root/spmv_samples_b200/csrc" -j8 EXTRA="$flags" OBJDIR=/tmp/spmv_variants/$name OUT="$root/tools/variants/$name.so" > /dev/null
    echo "built tools/variants/$name.so ($flags)"
done
