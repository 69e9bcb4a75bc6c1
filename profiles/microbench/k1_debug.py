import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dbb_b200 import device, synthetic
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
if cfg == 'LONG':
    meta = synthetic.make_metagenome(seed=9, n_scaffolds=64, min_len=200000, max_len=400000, n_genomes=8, extra_lengths=(5000000,) * int(200 * scale))
else:
    meta = synthetic.make_config(cfg, scale)
d_ascii, seq_off = device.synth_ascii(meta)
packed = device.pack_ascii(d_ascii, seq_off)
del d_ascii
plan = device.CountPlan(packed)
n = meta.n_scaffolds
counts = torch.zeros((n, 136), dtype=torch.int32, device='cuda'); freq = torch.zeros((n, 140), dtype=torch.float32, device='cuda')
for _ in range(3): plan.run(counts, freq)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.run(counts, freq)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
bytes_ = packed.algorithmic_bytes()
print('%s x%.2f: %d scaffolds %.3f Gbp  %.3f ms  %.1f Gbp/s  %.1f GB/s (%.1f%% of 6541)  skip_flush=%s' % (
    cfg, scale, n, meta.total_bp / 1e9, ms, meta.total_bp / ms / 1e6, bytes_ / ms / 1e6, 100 * bytes_ / ms / 1e6 / 6541.1,
    os.environ.get('DBB_K1_DEBUG_SKIP_FLUSH', '0')), flush=True)
