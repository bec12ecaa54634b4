#!/bin/bash
# compute-sanitizer over the mbarrier / cluster / DSMEM / tcgen05 code.  One GPU.
# racecheck does not model cp.async.bulk / tcgen05 (async proxy) accesses, so its report covers the generic-proxy code only.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py decode graph attention > $O/r02_sanitizer_memcheck_decode.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_memcheck_decode.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py blocks train > $O/r02_sanitizer_memcheck_blocks_train.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_memcheck_blocks_train.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_run.py decode attention train > $O/r02_sanitizer_synccheck.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_synccheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py decode > $O/r02_sanitizer_racecheck_decode.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_racecheck_decode.log
echo done
