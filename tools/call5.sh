set -x
mkdir -p gpurun_out
for c in 22 13 12 0; do ISLS_FF_BIG=$c python tools/bench_small.py 65536 > gpurun_out/c5_ffbig$c.log 2>&1; done
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/c5_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/c5_bench_ref.json 2> gpurun_out/c5_bench_ref.err
