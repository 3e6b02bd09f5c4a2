#!/bin/bash
mkdir -p gpurun_out
{
python tools/bench_small.py 65536 || exit 1
python tools/bench_small.py 8192 32768
} > gpurun_out/xpart_bench.txt 2>&1
python tools/variant_diff.py > gpurun_out/xpart_variants.txt 2>&1
