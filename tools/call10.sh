set -x
mkdir -p gpurun_out
python tools/bench_small.py 8192 65536 > gpurun_out/c10_pf1.log 2>&1
ISLS_LS_PREFETCH=0 python tools/bench_small.py 8192 65536 > gpurun_out/c10_pf0.log 2>&1
python tools/variant_diff.py > gpurun_out/c10_variant_diff.log 2>&1
