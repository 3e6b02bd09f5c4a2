set -x
mkdir -p gpurun_out
python tools/bench_small.py 4096 8192 16384 65536 > gpurun_out/c12_unrolled.log 2>&1
python tools/variant_diff.py > gpurun_out/c12_variant_diff.log 2>&1
BENCH_MODEL=arm python tools/bench_small.py 16384 > gpurun_out/c12_arm.log 2>&1
