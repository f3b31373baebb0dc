echo "#### default (tile pf + level pf), offset form"; bash tools/gpu_ab_one.sh quick
echo "#### tilepf only, offset form"; ADDER_B200_SO=$PWD/build_variants/lib_tilepf.so bash tools/gpu_ab_one.sh quick
echo "#### no pf, offset form"; ADDER_B200_SO=$PWD/build_variants/lib_nopf.so bash tools/gpu_ab_one.sh quick
echo "#### default lib, eager form (tile pf on)"; ADDER_B200_OFFSET=0 bash tools/gpu_ab_one.sh quick
