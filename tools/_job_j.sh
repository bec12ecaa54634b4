timeout 400 python tools/time_generate.py 1 32 --modes 1,0 --prof --iters 3 > gpurun_out/r2_j_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_j_time.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_synth.py -q -x > gpurun_out/r2_j_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_j_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_j_bench.json 2> gpurun_out/r2_j_bench.err; echo "rc=$?" >> gpurun_out/r2_j_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o gpurun_out/r2_j_decode_b32 \
    python tools/profile_step.py --batch 32 --steps 40 --no-ssrn > gpurun_out/r2_j_ncu_b32.log 2>&1
echo done
