#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_reference_suite.py -q -k "gradcheck or gradient" 2>&1 | grep -v "^E   *\[\|^E   *[0-9-]" | head -120
