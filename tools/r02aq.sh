#!/bin/bash
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
run() { # tool part timeout
  timeout $3 compute-sanitizer --tool $1 --print-limit 20 python tools/sanitize_round2b.py $2 > gpurun_out/r02aq_sanitizer_$1_$2.txt 2>&1
  echo "$1 $2 exit $?"; grep -v "warning\|Remark\|constexpr\|\^\|detected during\|^$" gpurun_out/r02aq_sanitizer_$1_$2.txt | tail -6
}
run memcheck wide 200
run memcheck unrolled 150
run memcheck adjoint 150
run synccheck wide 200
NS=128 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_wide_conv|k_gn_backward_ex|k_wide_raw_op" -c 8 -f -o gpurun_out/r02aq_wide_vjp python tools/wide_vjp_debug.py > gpurun_out/r02aq_ncu.log 2>&1; echo "ncu exit $?"
