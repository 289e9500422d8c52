#!/bin/bash
for c in 1 2 3; do echo "== cold $c"; COLD=$c timeout 120 python scripts/fused_timeline.py citeseer 2>/dev/null; done
