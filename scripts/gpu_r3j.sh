#!/bin/bash
for i in 1 2 3; do
  echo -n "pdl    "; timeout 600 python scripts/time_bilevel_block.py citeseer 30 2>/dev/null | head -1
  echo -n "no_pdl "; LDS_NO_PDL=1 timeout 600 python scripts/time_bilevel_block.py citeseer 30 2>/dev/null | head -1
done
