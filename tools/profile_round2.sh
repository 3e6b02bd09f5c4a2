#!/bin/bash
# The commands behind profiles/r2_* (run on a B200 box from the repository root, e.g. through gpurun).
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke
python -m pytest tests -m gpu -q
python bench.py --steps 20 --warmup 5                        > gpurun_out/r2_bench_n1.json
python bench.py --early-exit --steps 5 --warmup 3 --no-cpu-baseline      > gpurun_out/r2_bench_n1_early_exit.json
python bench.py --notebook-budget --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_notebook_budget.json
python bench.py --impl reference --steps 2 --warmup 1        > gpurun_out/r2_bench_reference_arm.json
python tools/bench_configs.py                                > gpurun_out/r2_configs.txt
python tools/bench_small.py 4096 8192 16384 32768 65536      > gpurun_out/r2_small_batches.txt
python tools/probe_overlap.py                                > gpurun_out/r2_probe_overlap.txt
python tools/variant_diff.py                                 > gpurun_out/r2_variant_diff.txt
# launch list of the timed step (shares must agree with the CUDA-event shares of the bench line)
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launch_list_ncu.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline
# full captures: C5 kernels, small-batch kernels, SLS kernels  (condense with tools/ncu_select.py / tools/ncu_source_top.py)
ncu --set full --clock-control none --import-source on -k regex:"k_ff_tma|k_linesearch|k_kpass" -c 4 -o gpurun_out/r2_c5 \
    python tools/run_car_small.py 65536
ncu --set full --clock-control none --import-source on -k regex:"k_ff_tma|k_linesearch|k_kpass|k_outer_end" -c 8 \
    -o gpurun_out/r2_car8192 python tools/run_car_small.py 8192
ncu --set full --clock-control none --import-source on -k regex:"k_sls_admm|k_sls_ctrl|k_dgemm" -c 12 -o gpurun_out/r2_sls \
    python tools/bench_sls.py
# multi-GPU (gpurun --gpus N): strong + weak scaling in one line per N
# python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus N
