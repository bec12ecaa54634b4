O=gpurun_out; mkdir -p $O
timeout 300 python tools/time_generate.py 1 32 --modes 1 --prof --iters 3 > $O/r2_b_time.log 2>&1; echo "rc=$?" >> $O/r2_b_time.log
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_synth.py tests/test_reference_shim.py -q -x -m gpu > $O/r2_b_pytest.log 2>&1; echo "rc=$?" >> $O/r2_b_pytest.log
