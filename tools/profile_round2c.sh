#!/bin/bash
# The commands behind profiles/r2c_* (second session of round 2, final code; run on a B200 box from the repository root).
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2c_gputests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_ref.err; echo "ref rc=$?"
python tools/bench_small.py 4096 8192 16384 32768 65536 > gpurun_out/r2c_small_batches.txt 2>&1
python tools/bench_configs.py > gpurun_out/r2c_configs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2c_launch_list_ncu.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ff_tma|k_linesearch|k_kpass" -c 4 -o gpurun_out/r2c_c5 \
    python tools/run_car_small.py 65536 > gpurun_out/r2c_ncu_c5.log 2>&1
