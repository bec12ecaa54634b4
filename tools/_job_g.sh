timeout 400 python tools/time_generate.py 4 6 7 8 12 14 16 24 28 32 --modes 2 --iters 3 > gpurun_out/r2_g_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_g_time.log
echo done
