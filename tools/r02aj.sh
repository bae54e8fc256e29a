#!/bin/bash
timeout 600 python tools/wide_vjp_debug.py 2>&1 | tail -8
