#!/bin/bash
timeout 600 python tools/wide_train_bench.py 2>&1 | tail -8
