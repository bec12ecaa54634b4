#!/bin/bash
# compute-sanitizer over the mbarrier / cluster / DSMEM / tcgen05 code.  One GPU.
# racecheck does not model cp.async.bulk / tcgen05 (async proxy) accesses, so its report covers the generic-proxy code only.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for tool in memcheck synccheck; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py decode graph attention > $O/r02_sanitizer_${tool}_decode.log 2>&1
    echo "rc=$?" >> $O/r02_sanitizer_${tool}_decode.log
done
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_run.py blocks > $O/r02_sanitizer_memcheck_blocks.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_memcheck_blocks.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_run.py decode > $O/r02_sanitizer_racecheck_decode.log 2>&1
echo "rc=$?" >> $O/r02_sanitizer_racecheck_decode.log
echo done
