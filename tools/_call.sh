set -x
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c5_n${N}_b.json 2> gpurun_out/r02_bench_c5_n${N}_b.err; tail -3 gpurun_out/r02_bench_c5_n${N}_b.err
done
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | grep -E "NVLS|P2P|Channel 00" | head -8 > gpurun_out/r02_nccl_info_n8.log
