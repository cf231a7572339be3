#!/bin/bash
# usage: retry.sh <outfile> <timeout> <command>
out=$1; to=$2; shift 2
for i in $(seq 1 15); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $out 2>&1
  if ! grep -q "status=transient" $out; then exit 0; fi
  sleep 100
done
