set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c6_bench_n2.json 2> gpurun_out/c6_bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-seconds 5 > gpurun_out/c6_ref_n2.json 2> gpurun_out/c6_ref_n2.err
tail -3 gpurun_out/c6_bench_n2.err
