#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): GPU test suite, smoke, bench line, ncu launch list of the bench workload, full
# captures of the dominant kernels.  Numbers printed under ncu are never bench values.  Outputs under gpurun_out/ (copied to profiles/).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02_gputest.log 2>&1; echo "rc=$?" >> $O/r02_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "rc=$?" >> $O/r02_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "rc=$?" >> $O/r02_bench_n1.err
# every launch of one whole pass (TextEnc + persistent decode + SSRN), cold cache, serialised: compare SHARES
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches_b32_pass.csv \
    python tools/profile_step.py --batch 32 --steps 210 > $O/r02_launches.log 2>&1
# the persistent decode kernel, one full capture of the real 210-frame launch at B = 32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o $O/r02_decode_cluster_b32 \
    python tools/profile_step.py --batch 32 --steps 210 --no-ssrn > $O/r02_ncu_decode.log 2>&1
# the training GEMM kernel: the first (2, 64) forward launch of a Text2Mel step at B = 32
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 1 -o $O/r02_gemm_tc_b32 \
    python tools/bench_train.py --steps 1 --warmup 1 > $O/r02_ncu_gemm.log 2>&1
echo done
