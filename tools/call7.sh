set -x
mkdir -p gpurun_out
python tools/bench_sls.py > gpurun_out/c7_sls.log 2>&1
ISLS_SLS_CTRL_DENSE=1 python tools/bench_sls.py > gpurun_out/c7_sls_dense.log 2>&1
python tools/variant_diff.py > gpurun_out/c7_variant_diff.log 2>&1
python -m pytest tests/test_gpu_projections.py tests/test_gpu_sls.py tests/test_gpu_parity.py::test_host_admm_driver_vs_reference_golden tests/test_gpu_parity.py::test_di_lqt_admm_batch_form_and_stage_api -m gpu -q 2>&1 | tail -30 > gpurun_out/c7_pytest.log
ncu --set full --clock-control none --import-source on -k regex:"k_sls_admm|k_sls_ctrl|k_dgemm" -c 8 -o gpurun_out/c7_sls python tools/bench_sls.py > gpurun_out/c7_ncu.log 2>&1
