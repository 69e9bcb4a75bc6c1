"""Stage-1 throughput of one kernel class on scaffolds of ONE length (windows / clk / SM at 1965 MHz, 148 SMs).
usage: [DBB_K1_T1=.. DBB_K1_T2=.. DBB_K1_T3=.. DBB_K1_DEBUG_SKIP_FLUSH=1] python k1_classes.py TAG L [TOTAL_MBP]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from dbb_b200 import device, synthetic
tag, L = sys.argv[1], int(sys.argv[2])
total = float(sys.argv[3]) * 1e6 if len(sys.argv) > 3 else 400e6
n = max(1, int(total // L))
meta = synthetic.make_metagenome(seed=7, n_scaffolds=n, min_len=L, max_len=L, n_genomes=50, law='uniform', frac_n=0.0, frac_lower=0.0)
d_ascii, seq_off = device.synth_ascii(meta)
packed = device.pack_ascii(d_ascii, seq_off)
del d_ascii
plan = device.CountPlan(packed)
counts = torch.zeros((n, 136), dtype=torch.int32, device='cuda'); freq = torch.zeros((n, 140), dtype=torch.float32, device='cuda')
for _ in range(3): plan.run(counts, freq)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.run(counts, freq)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
bp = float(meta.total_bp)
print('%-14s L=%7d n=%7d  %.4f ms  %7.1f Gbp/s  %5.2f win/clk/SM  %.3f of HBM  skip_flush=%s' % (
    tag, L, n, ms, bp / ms / 1e6, bp / (ms * 1e-3) / (148 * 1.965e9), packed.algorithmic_bytes() / ms / 1e6 / 6541.1,
    os.environ.get('DBB_K1_DEBUG_SKIP_FLUSH', '0')), flush=True)
