#!/bin/bash
mkdir -p gpurun_out
python tools/bench_small.py 8192 65536 > gpurun_out/epi_bench.txt 2>&1
