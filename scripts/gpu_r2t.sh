#!/bin/bash
echo "== warm"; timeout 120 python scripts/fused_timeline.py citeseer 2>/dev/null
