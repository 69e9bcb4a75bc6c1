import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get('e2e') or {}
        p = d.get('pruning') or {}
        print('%s: step %.2f ms | sig %.3f gather %s knn %s | %.0f Gbp/s  %s Gpairs/s | hbm frac %.3f | verified %s %s | second %.2f ms | e2e %s Gbp/s | count imbalance %.3f'
              % (f.split('/')[-1], d['ms_per_step'], d['stages_ms']['signatures'], d['stages_ms']['allgather'], d['stages_ms']['knn'], d['value'],
                 None if d['value2'] is None else round(d['value2'] / 1e9, 1), d['roofline']['frac'], d['verified_vs_oracle'],
                 (d.get('oracle_check_stage2') or {}).get('columns'), p.get('ms_second_sweep_and_merge', 0) or 0, e.get('value'), d['shard']['count_time_imbalance']))
    except Exception as ex:
        print(f, 'ERR', ex)
