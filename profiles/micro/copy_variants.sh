#!/bin/bash
# Historical: the A/B that decided the batched-copy kernel's shape (profiles/r02_copy_variants.log).  The variants
# (SGC_COPY_VARIANT 1 = single-item chunks with batched loads, 3 = multi-item chunks batched too) were removed after it;
# what can still be switched: SGC_COPY_NO_WIDE=1 (no 16-byte class), SGC_COPY_NO_SORT=1 (caller's item order).
for sw in "" "SGC_COPY_NO_WIDE=1" "SGC_COPY_NO_SORT=1"; do
  echo "switch: ${sw:-none}"
  env $sw python profiles/micro/r02_probe.py exchange 2>&1 | grep -E "level|exchange_.*_ms"
done
