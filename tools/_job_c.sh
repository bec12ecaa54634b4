timeout 300 python tools/time_generate.py 1 4 32 --modes 1,2,0 --prof --iters 3 > gpurun_out/r2_d_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_d_time.log
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_synth.py -q -x > gpurun_out/r2_d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_d_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_d_bench.json 2> gpurun_out/r2_d_bench.err; echo "rc=$?" >> gpurun_out/r2_d_bench.err
echo done
