set -x
mkdir -p gpurun_out
python tools/probe_overlap.py > gpurun_out/c4_probe.log 2>&1
python tools/variant_diff.py > gpurun_out/c4_variant_diff.log 2>&1
python -m pytest tests/test_gpu_parity.py::test_ragged_horizon_and_candidate_counts tests/test_gpu_tutorial.py::test_tutorial_fullsize_vs_reference_golden -m gpu -q 2>&1 | tail -60 > gpurun_out/c4_fail.log
ISLS_FF_MODE=0 python -m pytest tests/test_gpu_parity.py::test_ragged_horizon_and_candidate_counts tests/test_gpu_tutorial.py::test_tutorial_fullsize_vs_reference_golden -m gpu -q 2>&1 | tail -5 > gpurun_out/c4_fail_staged.log
