# re-entry job: GPU test suite, bench line, ncu launch list of one full pass, full capture of the decode kernel
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2_a_pytest.log 2>&1; echo "rc=$?" >> $O/r2_a_pytest.log
timeout 500 python bench.py --steps 8 --warmup 3 > $O/r2_a_bench.json 2> $O/r2_a_bench.err; echo "rc=$?" >> $O/r2_a_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches_b32_pass.csv \
    python tools/profile_step.py --batch 32 --steps 210 > $O/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o $O/r02_decode_cluster_b32 \
    python tools/profile_step.py --batch 32 --steps 40 --no-ssrn > $O/r02_ncu_decode.log 2>&1
echo done
