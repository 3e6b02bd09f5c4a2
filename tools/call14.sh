set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/c14_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c14_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/c14_bench_n1.json 2> gpurun_out/c14_bench_n1.err
python bench.py --early-exit --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_early.json 2>> gpurun_out/c14_bench_n1.err
python bench.py --notebook-budget --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench_nb.json 2>> gpurun_out/c14_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c14_bench_ref.json 2>> gpurun_out/c14_bench_n1.err
python tools/bench_configs.py > gpurun_out/c14_configs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/c14_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c14_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_ff_tma|k_linesearch|k_kpass" -c 4 -o gpurun_out/c14_c5 python tools/run_car_small.py 65536 > gpurun_out/c14_ncu_full.log 2>&1
