#!/bin/bash
# each experiment in its own process: a faulting one must not take the rest down
# args: cw ch mma mn lbo sbo mode [C H W]
for args in "0 0 1 1 4096 512 1" "0 0 1 1 512 4096 1" "0 0 1 1 4096 1024 2" "0 -1 1 1 4096 1024 2" "0 0 1 1 4096 1024 2 16 8 32"; do
  echo "=== $args"; timeout 30 ./tools/tc_debug $args 2>&1 | grep -v encode
done
