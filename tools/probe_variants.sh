#!/bin/bash
# usage: tools/probe_variants.sh "<variant libs (basename without lib/.so) or '-' for default>" <perf_probe cases...>
variants=$1; shift
for v in $variants; do
  echo "== variant $v"
  if [ "$v" != "-" ]; then export SVX_LIB=$PWD/shocovox_b200/lib$v.so; else unset SVX_LIB; fi
  timeout 600 python tools/perf_probe.py "$@" 2>&1 | grep -v "^+" | tail -12
done
