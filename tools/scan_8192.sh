#!/bin/bash
# A/B scan of the panel-CTA count and the sweep-ticket lag at 8192^2 (both engines); prints one line per configuration.
cd "$(dirname "$0")/.."
for np in "" 4 8 12 16 24; do
  for lag in "" 4 8 32; do
    QRB200_NP=$np QRB200_LAG=$lag python tools/tc3_time.py 8192 3 2>&1 | tail -1 | sed "s/$/ lag=${lag:--}/"
  done
done
for np in "" 8 12 16 24 32; do
  QRB200_NP=$np python tools/tc3_time.py 8192 0 2>&1 | tail -1
done
