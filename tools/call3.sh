set -x
mkdir -p gpurun_out
python tools/variant_diff.py > gpurun_out/c3_variant_diff.log 2>&1
python tools/bench_small.py 8192 65536 > gpurun_out/c3_default.log 2>&1
ISLS_OVERLAP=0 python tools/bench_small.py 65536 > gpurun_out/c3_noovl.log 2>&1
for c in 0 1 3; do ISLS_OVL_LS_CTAS=$c python tools/bench_small.py 65536 > gpurun_out/c3_ovl_ls$c.log 2>&1; done
for f in 0 3 5 6; do ISLS_OVL_FF_DEPTH=$f python tools/bench_small.py 65536 > gpurun_out/c3_ovl_ff$f.log 2>&1; done
ISLS_OVERLAP=1 python tools/bench_small.py 16384 32768 > gpurun_out/c3_ovl_mid.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c3_pytest.log
