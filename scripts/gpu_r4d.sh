#!/bin/bash
python scripts/fused_timeline.py citeseer 2>/dev/null | tail -28
