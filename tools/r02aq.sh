#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_round2b.py > gpurun_out/r02aq_sanitizer_${tool}.txt 2>&1
  echo "$tool exit $?"; grep -c "Error\|error" gpurun_out/r02aq_sanitizer_${tool}.txt; tail -3 gpurun_out/r02aq_sanitizer_${tool}.txt
done
# kernel evidence for the wide adjoint at the reference's batch (256 filters, 128 images): launch list + full captures
NS=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wide_conv|k_gn_backward_ex|k_wide_raw_op" -c 8 -f -o gpurun_out/r02aq_wide_vjp python tools/wide_vjp_debug.py > gpurun_out/r02aq_ncu.log 2>&1; echo "ncu exit $?"
