set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py::test_non_default_device_matches_device_0 -m gpu -q 2>&1 | tail -30 > gpurun_out/c13_dev1.log
