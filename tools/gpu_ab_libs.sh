#!/bin/bash
# The in-tree library, then every build_variants/lib*.so, on the quick workload set.  Usage (under gpurun): bash tools/gpu_ab_libs.sh [quick]
echo "#### in-tree"; bash tools/gpu_ab_one.sh ${1:-quick}
for so in build_variants/lib*.so; do
  [ -f "$so" ] || continue
  echo "#### $so"; ADDER_B200_SO=$PWD/$so bash tools/gpu_ab_one.sh ${1:-quick}
done
