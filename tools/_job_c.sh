O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_train.py tests/test_gpu_trainer_run.py -q -m gpu > $O/r2_c_pytest.log 2>&1; echo "rc=$?" >> $O/r2_c_pytest.log
timeout 300 python tools/bench_train.py --steps 10 --warmup 3 > $O/r2_c_train_tc1.json 2> $O/r2_c_train_tc1.err; echo "rc=$?" >> $O/r2_c_train_tc1.err
timeout 300 python tools/bench_train.py --net 2 --steps 5 --warmup 2 > $O/r2_c_train_ssrn.json 2>> $O/r2_c_train_tc1.err
