#!/bin/bash
# CTC loss schedules at T=2048, C=7375: default (relative split for B <= 24), 4 = split with normalised tables, 5 = relative split, 2 = rows
mkdir -p gpurun_out
for rep in 1 2; do
for b in 16 2 8; do
for m in "" 4 5 2; do echo -n "B=$b HCTR_CTC_OVERLAP=$m  "; HCTR_CTC_OVERLAP=$m timeout 300 python scripts/ctc_bench.py $b 2>&1 | grep -o '"ms": [0-9.]*\|"loss": [0-9.]*' | tr '\n' ' '; echo; done
done
done
