#!/bin/bash
cp lds-gnn_b200/lib/liblds_b200.so /tmp/cur.so
for v in liblds_old lib_v1 lib_v3 liblds_old lib_v1; do
  cp ab/$v.so lds-gnn_b200/lib/liblds_b200.so
  echo "== $v"; python scripts/fused_timeline.py citeseer 2>/dev/null | grep -E "sampled|barrier2|end "
done
cp /tmp/cur.so lds-gnn_b200/lib/liblds_b200.so
