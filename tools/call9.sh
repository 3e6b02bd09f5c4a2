set -x
mkdir -p gpurun_out
BENCH_MODEL=arm python tools/bench_small.py 2048 16384 > gpurun_out/c9_arm.log 2>&1
ISLS_FF_MODE=0 BENCH_MODEL=arm python tools/bench_small.py 16384 > gpurun_out/c9_arm_staged.log 2>&1
python tools/variant_diff.py > gpurun_out/c9_variant_diff.log 2>&1
python tools/bench_configs.py --quick > gpurun_out/c9_configs_quick.log 2>&1
