#!/bin/bash
mkdir -p gpurun_out
{
echo "== default"; python tools/bench_small.py 14336 16384 18944 24576 32768 || exit 1
echo "== ISLS_LS_CPT=54"; ISLS_LS_CPT=54 python tools/bench_small.py 8192 14336 16384 18944 24576
} > gpurun_out/ls54.txt 2>&1
