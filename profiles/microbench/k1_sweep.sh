for cfg in "256 128" "128 128" "128 64" "256 64" "512 128" "128 256"; do set -- $cfg; 
DBB_K1_THR3=$1 DBB_K1_THR2=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', d['stages_ms'], d['roofline']['frac'])"
done
for t in "44 512 4300" "32 512 4300" "44 768 4300" "44 384 4300" "44 256 4300" "64 512 4300" "44 512 2048"; do set -- $t;
DBB_K1_T1=$1 DBB_K1_T2=$2 DBB_K1_T3=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('T $t', d['stages_ms']['signatures'], d['roofline']['frac'])"
done
