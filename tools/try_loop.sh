#!/bin/bash
mkdir -p gpurun_out
python tools/bench_small.py 4096 8192 16384 65536 > gpurun_out/epi_bench.txt 2>&1
python tools/variant_diff.py > gpurun_out/epi_variants.txt 2>&1
