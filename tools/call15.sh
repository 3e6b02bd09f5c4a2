set -x
mkdir -p gpurun_out
python tools/variant_diff.py > gpurun_out/c15_variant_diff.log 2>&1
BENCH_MODEL=arm python tools/bench_small.py 2048 16384 > gpurun_out/c15_arm_ws.log 2>&1
ISLS_FF_WS=0 BENCH_MODEL=arm python tools/bench_small.py 16384 > gpurun_out/c15_arm_nows.log 2>&1
python tools/bench_configs.py --quick > gpurun_out/c15_configs_quick.log 2>&1
