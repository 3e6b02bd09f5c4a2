set -x
mkdir -p gpurun_out
python tools/bench_sls.py > gpurun_out/c8_sls.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/c8_pytest.log
python tools/bench_configs.py > gpurun_out/c8_configs.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_sls_admm|k_sls_ctrl" -c 4 -o gpurun_out/c8_sls python tools/bench_sls.py > gpurun_out/c8_ncu.log 2>&1
