set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernel_variants.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c2_variants.log
python tools/bench_small.py 4096 8192 16384 32768 > gpurun_out/c2_small_tma.log 2>&1
ISLS_FF_JC=0 python tools/bench_small.py 4096 8192 16384 > gpurun_out/c2_small_tma_nojc.log 2>&1
ISLS_FF_MODE=0 python tools/bench_small.py 8192 > gpurun_out/c2_small_staged.log 2>&1
ISLS_FF_MODE=2 python tools/bench_small.py 65536 > gpurun_out/c2_tma_65536.log 2>&1
ISLS_FF_MODE=2 ISLS_FF_JC=0 python tools/bench_small.py 65536 > gpurun_out/c2_tma_nojc_65536.log 2>&1
BENCH_MODEL=arm python tools/bench_small.py 2048 16384 > gpurun_out/c2_small_arm.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/c2_pytest.log
ncu --set full --clock-control none --import-source on -k regex:"k_ff_tma|k_kpass" -c 3 -o gpurun_out/c2_car8192 python tools/run_car_small.py 8192 > gpurun_out/c2_ncu.log 2>&1
