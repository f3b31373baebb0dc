export ADDER_B200_OFFSET=0
run_t() { timeout 90 python tools/profile_run.py --reps 3 --batch --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.5 --warm-frames 608 2>&1 | grep -E "rep 2|rror" | sed -e 's/^/   /'; }
run_d() { timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:integrate_frame -s 19 -c 1 python tools/profile_run.py --w 7680 --h 4320 --c 1 --kind 3 --crf 3 --ref 256 --dtm 1048576 --frames 32 --cap 0.25 --warm-frames 608 --batch --reps 1 2>&1 | grep -E "dram__|gpu__time" | sed -e 's/^/   /'; }
echo "#### eager form, in-tree"; run_t; run_d
export ADDER_B200_SO=$PWD/build_variants/lib_nltc.so
echo "#### eager form, deep-level loads with L2::64B"; run_t; run_d
