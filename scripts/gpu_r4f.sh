#!/bin/bash
cp lds-gnn_b200/lib/liblds_b200.so /tmp/new.so
for v in old new old new; do
  if [ $v = old ]; then cp ab/liblds_old.so lds-gnn_b200/lib/liblds_b200.so; else cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so; fi
  echo "== $v"; python scripts/fused_timeline.py citeseer 2>/dev/null | grep -E "sampled|barrier1|feat_done|barrier2|epi._done|after.|end |tiles_done|cta  60"
done
cp /tmp/new.so lds-gnn_b200/lib/liblds_b200.so
