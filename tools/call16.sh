set -x
mkdir -p gpurun_out
for s in 0 10000 20000 30000; do ISLS_LS_STAGGER_NS=$s python tools/bench_small.py 65536 > gpurun_out/c16_stag$s.log 2>&1; done
ISLS_LS_STAGGER_NS=20000 python tools/variant_diff.py > gpurun_out/c16_variant_diff.log 2>&1
