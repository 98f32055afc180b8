# Round-2 evidence captures (one B200).  Every ncu pass runs only after the same command exited 0 without ncu.
set -x
python bench.py --steps 3 --warmup 3 --no-cpu --quick > gpurun_out/bench_ll_plain.json 2> gpurun_out/bench_ll_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --quick > gpurun_out/ncu_ll.log 2>&1
python tools/prof_targets.py 8 > gpurun_out/prof_targets_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_traffic.csv python tools/prof_targets.py 8 > gpurun_out/ncu_lt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sift_descriptor_kernel|sift_refine_kernel|sift_orient_kernel|sift_extrema|sift_blur_tma_kernel|sift_base_stream|sift_small_oct|sift_rank_bucket" -c 33 -o gpurun_out/r2_sift -f python tools/prof_targets.py 8 > gpurun_out/ncu_sift.log 2>&1
python tools/prof_match.py 32768 match > gpurun_out/prof_match_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:match_topk_u8 -s 2 -c 1 -o gpurun_out/r2_match_u8 -f python tools/prof_match.py 32768 match > gpurun_out/ncu_match.log 2>&1
ls -la gpurun_out/
python tools/prof_float.py 32768 > gpurun_out/prof_float_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"match_topk_kernel" -s 1 -c 1 -o gpurun_out/r2_match_float -f python tools/prof_float.py 32768 > gpurun_out/ncu_float.log 2>&1
