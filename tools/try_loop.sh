#!/bin/bash
mkdir -p gpurun_out
{
for pd in 1 2 3 4; do echo "== ISLS_LQT_PD=$pd"; ISLS_LQT_PD=$pd python tools/probe_c1.py 2>&1 | grep -v "^B 1024.*75\." | awk 'NR%3==0'; done
} > gpurun_out/lqt_pd.txt 2>&1
python tools/variant_diff.py > gpurun_out/lqt_pd_variants.txt 2>&1
