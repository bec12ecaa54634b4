O=gpurun_out; mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 640 --csv --log-file $O/r02_launches_train_step_b32_tc.csv python tools/bench_train.py --steps 2 --warmup 4 > $O/r2_d_ncu.log 2>&1
bash tools/r02_sanitize.sh
