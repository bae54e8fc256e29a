#!/bin/bash
mkdir -p gpurun_out
for part in wide unrolled; do
timeout 170 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_round2b.py $part > gpurun_out/r02bh_sanitizer_racecheck_$part.txt 2>&1
echo "racecheck $part exit $?"; grep -v "warning\|Remark\|constexpr\|\^\|detected during\|^$" gpurun_out/r02bh_sanitizer_racecheck_$part.txt | grep -v "Host Frame\|^=========     " | tail -8
done
