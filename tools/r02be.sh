#!/bin/bash
for v in "NODE_B200_ADJOINT_OVERLAP=0" "NODE_B200_ADJOINT_SOLVE=0"; do
  env $v timeout 150 compute-sanitizer --tool racecheck --print-limit 3 python tools/sanitize_round2b.py adjoint > /tmp/rc.txt 2>&1
  echo "$v exit $?"; grep -v "warning\|Remark\|constexpr\|\^\|detected during\|^$" /tmp/rc.txt | grep -v "^=========     \|Host Frame" | tail -8
done
