# 2-GPU box: bit-identity test, scaling bench N=1 and N=2 back to back (same box), NCCL topology line
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r2_b_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2_b_multi.log
timeout 400 python bench.py --gpus 1 --steps 8 --warmup 3 --cpu-passes 0 --no-parity-check > gpurun_out/r2_b_bench_n1.json 2> gpurun_out/r2_b_bench_n1.err
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_b_bench_n2.json 2> gpurun_out/r2_b_bench_n2.err
for c in 1 2 8; do timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$c bench.py --gpus 2 --steps 8 --warmup 3 --gather-chunks $c > gpurun_out/r2_b_bench_n2_chunks$c.json 2>/dev/null; done
grep -E "NVLS|P2P|via" gpurun_out/r2_b_bench_n2.err | head -20 > gpurun_out/r2_b_nccl_topology.txt
echo done
