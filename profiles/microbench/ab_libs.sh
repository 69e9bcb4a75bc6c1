#!/bin/bash
# A/B of several builds of libdbbcuda.so on the same box: ab_libs.sh "<bench args>" name=path ...
args="$1"; shift
for rep in 1 2; do
for nv in "$@"; do
  name="${nv%%=*}"; path="${nv#*=}"
  if [ "$path" = "default" ]; then unset DBB_LIB_PATH; else export DBB_LIB_PATH="$path"; fi
  python bench.py $args --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '$args', round(d['value2']/1e9,2), 'Gpairs/s', round(d['stages_ms']['knn'],3), 'ms')"
done
done
