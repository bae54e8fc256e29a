#!/bin/bash
mkdir -p gpurun_out
python -c "import sys; sys.path.insert(0,'.'); import __graft_entry__ as g; print('stale', g._stale())"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02fin3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02fin3_pytest.log
tail -3 gpurun_out/r02fin3_pytest.log
for tool in memcheck racecheck; do
timeout 200 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_round2b.py adjoint > gpurun_out/r02bd_sanitizer_${tool}_adjoint_overlap.txt 2>&1
echo "$tool exit $?"; grep -v "warning\|Remark\|constexpr\|\^\|detected during\|^$" gpurun_out/r02bd_sanitizer_${tool}_adjoint_overlap.txt | tail -4
done
python __graft_entry__.py smoke 2>&1 | tail -1
