for v in 0 1; do
VO_EXT_TMA=$v ncu --set full --clock-control none --import-source on -k regex:sift_extrema -c 4 -o gpurun_out/r2_ext_tma$v -f python tools/prof_targets.py 8 2000 > gpurun_out/ncu_ext$v.log 2>&1
echo rc=$? ; tail -2 gpurun_out/ncu_ext$v.log
done
