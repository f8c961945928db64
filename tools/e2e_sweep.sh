#!/bin/bash
# host-memory path: result hand-off mode x chunk size
for mode in ${MODES:-zerocopy}; do for chunk in ${CHUNKS:-98304 131072 163840 196608 245760 262144 278528 524288}; do
  echo "mode=$mode chunk=$chunk"
  T2FIT_HOST_OUT=$mode T2FIT_HOST_CHUNK=$chunk python tools/e2e_probe.py 2>&1 | grep "^call" | sed -n 3,6p | awk '{printf "   %s", $0} END {print ""}'
done; done
