O=gpurun_out; mkdir -p $O
for c in 1 2; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2964$c bench.py --gpus 4 --steps 8 --warmup 3 --gather-chunks $c --train-steps 3 > $O/r02_scale_n4_chunks$c.json 2>/dev/null; done
echo done
