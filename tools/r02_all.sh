# round-2 final captures: everything under gpurun_out/r02_* in one GPU call (then: python tools/make_profiles.py r02)
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_launches_bench.log 2>&1
USRT_NO_GRAPH=1 timeout 600 $NCU --set full --import-source on -k regex:"k_trace_primary|k_construct_bvh|k_morton|k_histogram|k_distribute_keys|k_construct_tree|k_onesweep|k_scan_histogram" --launch-skip 22 --launch-count 11 -f -o gpurun_out/r02_prof_build_trace python tools/ncu_step.py > gpurun_out/r02_prof_step.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_onesweep -s 28 -c 4 -f -o gpurun_out/r02_prof_sort ./tools/micro/lab_a > gpurun_out/r02_prof_sort.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_trace_primary --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_prof_trace_c1 python tools/ncu_trace.py c1 > gpurun_out/r02_prof_c1.log 2>&1
timeout 900 $NCU --set full --import-source on -k regex:k_trace_rays --launch-skip 2 --launch-count 1 -f -o gpurun_out/r02_prof_trace_c4 python tools/ncu_trace.py c4 > gpurun_out/r02_prof_c4.log 2>&1
./tools/micro/cub_sort_calib > gpurun_out/r02_cub.txt 2>&1
python tools/sort_bench.py --n 20 22 24 26 28 30 > gpurun_out/r02_sort_sweep.txt 2>&1
python tools/config_bench.py > gpurun_out/r02_configs.txt 2>&1
ls -la gpurun_out/r02_*
tail -n 3 gpurun_out/r02_sort_sweep.txt; tail -n 3 gpurun_out/r02_configs.txt
