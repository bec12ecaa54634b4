timeout 300 python tools/time_generate.py 1 4 32 --modes 1,2,0 --prof --iters 3 > gpurun_out/r2_a_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_a_time.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_a_bench.json 2> gpurun_out/r2_a_bench.err; echo "rc=$?" >> gpurun_out/r2_a_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_a_ref.json 2> gpurun_out/r2_a_ref.err
timeout 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_trainer_run.py > gpurun_out/r2_a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_a_pytest.log
timeout 900 python -m pytest tests/test_gpu_trainer_run.py -q -x > gpurun_out/r2_a_trainer.log 2>&1; echo "rc=$?" >> gpurun_out/r2_a_trainer.log
echo done
