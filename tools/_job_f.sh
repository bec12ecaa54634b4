O=gpurun_out; mkdir -p $O
timeout 300 python tools/bench_train.py --net 2 --steps 5 --warmup 2 > $O/r02_bench_train_ssrn_n1.json 2> $O/r2_f_ssrn.err; echo "rc=$?" >> $O/r2_f_ssrn.err
timeout 300 python tools/bench_train.py --net 2 --steps 5 --warmup 2 --train-tc 0 > $O/r02_bench_train_ssrn_n1_fp32.json 2>> $O/r2_f_ssrn.err; echo "rc=$?" >> $O/r2_f_ssrn.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r02_launches_train_ssrn_b32.csv python tools/bench_train.py --net 2 --steps 1 --warmup 1 > $O/r2_f_ncu0.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/r02_launches_train_ssrn_b32_probe.csv python tools/bench_train.py --net 2 --steps 1 --warmup 1 --probe 1 > $O/r2_f_ncu1.log 2>&1
timeout 300 python tools/bench_train.py --steps 10 --warmup 3 > $O/r2_f_train_t2m.json 2>> $O/r2_f_ssrn.err
