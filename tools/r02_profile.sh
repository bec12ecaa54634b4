#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): bench line, ncu launch list of the bench workload, full captures of the
# dominant kernels.  Numbers printed under ncu are never bench values.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/r02_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err
kill $SMI
# every launch of one whole pass (TextEnc + persistent decode + SSRN), cold cache, serialised: compare SHARES
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches_b32_pass.csv \
    python tools/profile_step.py --batch 32 --steps 210 > $O/r02_launches.log 2>&1
# the persistent decode kernel, one full capture (B = 32, 60 frames keep the ~40 replays short)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o $O/r02_decode_cluster_b32 \
    python tools/profile_step.py --batch 32 --steps 60 --no-ssrn > $O/r02_ncu_decode.log 2>&1
# the tcgen05 block kernel at the benchmark shape
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ln_tc -s 1 -c 1 -o $O/r02_conv_ln_tc_hc11_b32 \
    python tools/profile_block.py SSRN/HC_11 32 840 1 > $O/r02_ncu_hc11.log 2>&1
echo done
