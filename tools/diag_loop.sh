#!/bin/bash
# usage: tools/diag_loop.sh <lib.so> <n>  — n fresh processes of tools/diag_bad.py; prints per-run verdicts
lib=$1; n=$2
for i in $(seq 1 $n); do
  TRTOOLS_B200_LIB=$lib python tools/diag_bad.py 100000 2>&1 | grep -E "rep 0 (ok|FAILED)|rep 0 vs last|rep 3 ok" | cut -c1-70 | tr '\n' ' '; echo
done
