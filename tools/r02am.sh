#!/bin/bash
NS=100,128,300 timeout 300 python tools/wide_vjp_debug.py 2>&1 | grep "^n " | tail -12
NS=128 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/wide_vjp_debug.py 2>&1 | grep -v "Warn\|detach\|return '" | head -20
timeout 600 python tools/wide_train_bench.py 2>&1 | tail -8
