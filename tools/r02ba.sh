#!/bin/bash
s=$(date +%s)
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> /tmp/ref.err | cut -c1-330
tail -2 /tmp/ref.err
echo "reference arm wall $(( $(date +%s) - s )) s"
