O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_gputest.log 2>&1; echo "rc=$?" >> $O/r02_gputest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "rc=$?" >> $O/r02_bench_n1.err
