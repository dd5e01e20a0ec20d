#!/bin/bash
# A/B of the batched-copy kernel on the C2 exchange (one GPU, one process each): SGC_COPY_VARIANT 0 = per-unit
# decode only, 1 = single-item chunks with batched loads (default), 2 = single-item chunks, not batched;
# SGC_COPY_NO_SORT=1 keeps the caller's item order (no small-first / padding).  Round-2 results on one B200
# (strip-sourced table / plain table, ms): sorted v0 0.0634/0.1008, v1 0.0610/0.0971, v2 0.0627/0.1280;
# unsorted v0 0.0675/0.0983, v1 0.0864/0.1092, v2 0.0758/0.1166.
for ns in "" 1; do for v in 0 1 2; do
    echo "no_sort=$ns variant=$v"
    SGC_COPY_NO_SORT=$ns SGC_COPY_VARIANT=$v python profiles/micro/r02_probe.py exchange 2>&1 | grep -E "exchange_.*_ms" | head -2
done; done
