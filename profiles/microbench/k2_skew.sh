#!/bin/bash
# skew experiment: k2_skew.sh <lib> <skew cycles...>
lib=$1; shift
for sk in "$@"; do
  DBB_LIB_PATH=$lib DBB_K2_SKEW=$sk python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$(basename $lib) skew $sk', round(d['value2']/1e9,2), 'Gpairs/s', round(d['stages_ms']['knn'],3), 'ms')"
done
