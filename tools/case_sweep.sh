#!/bin/bash
# each case (K parts M dtype per line, e.g. tools/cases.txt) in its own process with a short timeout, so one hang costs seconds
#   bash tools/case_sweep.sh < tools/cases.txt
while read -r K parts M dt; do
  [ -z "$K" ] && continue
  out=$(timeout -s KILL ${CASE_TIMEOUT:-25} python tools/case_check.py $K $parts $M $dt 2>&1 | tail -1)
  rc=$?
  echo "[$K $parts $M $dt] ${out:-HANG/KILLED rc=$rc}"
done
