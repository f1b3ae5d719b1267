#!/bin/bash
# A/B scan at 8192^2: panel-CTA count (QRB200_NP) x sweep-ticket lag (QRB200_LAG) for the tensor-core engine, panel-CTA
# count for the SIMT engine.  One line per configuration ("-" = the plan's default).  ~5 s of process start-up per line.
cd "$(dirname "$0")/.."
run() { env "$@" python tools/tc3_time.py 8192 "$TC" 2>&1 | tail -1 | sed "s/\$/  [$*]/"; }
TC=3
for np in - 4 16; do
  for lag in - 0 4 8 32; do
    args=()
    [ "$np" != - ] && args+=("QRB200_NP=$np")
    [ "$lag" != - ] && args+=("QRB200_LAG=$lag")
    run "${args[@]}" X=1
  done
done
TC=0
for np in - 8 16 32; do
  args=(); [ "$np" != - ] && args+=("QRB200_NP=$np")
  run "${args[@]}" X=1
done
