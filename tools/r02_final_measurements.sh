set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_onesweep -s 28 -c 4 -f -o gpurun_out/r02_prof_sort ./tools/micro/lab_a > gpurun_out/r02_prof_sort.log 2>&1
./tools/micro/cub_sort_calib > gpurun_out/r02_cub.txt 2>&1
python tools/sort_bench.py --n 20 22 24 26 28 30 > gpurun_out/r02_sort_sweep.txt 2>&1
python tools/config_bench.py > gpurun_out/r02_configs.txt 2>&1
tail -3 gpurun_out/r02_sort_sweep.txt gpurun_out/r02_configs.txt
