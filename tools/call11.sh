set -x
mkdir -p gpurun_out
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/c11_bench_n$n.json 2> gpurun_out/c11_bench_n$n.err
done
tail -3 gpurun_out/c11_bench_n8.err
