timeout 300 python -m pytest tests/test_gpu_multiframe.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for T in 1 8; do echo "#### TBLOCK=$T offset"; ADDER_B200_TBLOCK=$T bash tools/gpu_ab_one.sh quick; done
echo "#### TBLOCK=8 eager"; ADDER_B200_OFFSET=0 ADDER_B200_TBLOCK=8 bash tools/gpu_ab_one.sh quick
