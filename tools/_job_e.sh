# ncu source-level capture of the persistent decode kernel (stall reasons per SASS line) + a timing sanity check
timeout 200 python tools/time_generate.py 1 32 --modes 1,2 --iters 3 > gpurun_out/r2_e_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_e_time.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o gpurun_out/r2_e_decode_b1 \
    python tools/profile_step.py --batch 1 --steps 40 --no-ssrn > gpurun_out/r2_e_ncu_b1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 1 -c 1 -o gpurun_out/r2_e_decode_b32 \
    python tools/profile_step.py --batch 32 --steps 40 --no-ssrn > gpurun_out/r2_e_ncu_b32.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
