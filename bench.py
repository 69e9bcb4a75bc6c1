#!/usr/bin/env python
"""Benchmark of DBB's hot path on B200: tetranucleotide signatures (stage 1) + all-pairs TD kNN (stage 2).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

One STEP = one pass of the whole hot path over one batch: count + normalise every scaffold of the batch,
(N > 1: all-gather the signature matrix), k=20 nearest neighbours + TD neighbourhood size of every row.
Workload at N GPUs: BASELINE.json configs[1] per GPU ("C2": 50 000 scaffolds of 1-100 kb, ~1.07 Gbp, per
rank; the all-pairs problem is over all N*50 000 scaffolds), sequence synthesised on the device by the
seeded generator shared with the tests.  Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'Gbp/s tetramer signatures; scaffold-pairs/s TD kNN'
K_NN = 20
TD_PER = 80


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), float(d.get('sm_max_mhz', 1965.0)), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 1965.0, 'fallback (B200_PROFILING.md)'


def load_traffic():
    """dram bytes per launch of the counting stage from the committed ncu capture, if any."""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def workload(world, scale, name='C2'):
    """C2 (default): weak scaling, every rank brings 50 000 scaffolds.  C4: strong scaling, 1 000 000 scaffolds
    in total (BASELINE.json configs[3]) sharded over the ranks."""
    from dbb_b200 import synthetic
    kw = dict(synthetic.CONFIGS[name])
    if name == 'C2':
        kw['n_scaffolds'] = max(64, int(kw['n_scaffolds'] * scale)) * world
        kw['n_genomes'] = max(4, int(kw['n_genomes'] * min(1.0, scale * world)))
    else:
        kw['n_scaffolds'] = max(64 * world, int(kw['n_scaffolds'] * scale))
    return synthetic.make_metagenome(**kw), kw


def sub_meta(meta, ids, seed_shift):
    from dbb_b200.synthetic import Metagenome
    m = Metagenome()
    m.seed = meta.seed + seed_shift
    for name in ('genome', 'length', 'coverage', 'gc_obs', 'n_start', 'n_len', 'lc_start', 'lc_len'):
        setattr(m, name, getattr(meta, name)[ids])
    m.genome_gc, m.cum = meta.genome_gc, meta.cum
    return m


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port: literal restatement, multiprocessing fork-join over
# all host cores in the reference's own shape: one scaffold / one row per work item)
# ------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_count_worker(i):
    from oracle import dbb_oracle as O
    s = _CPU['seqs'][i]
    return sum(O.seq_counts_literal(s)), len(s)


def _cpu_row_worker(i):
    from oracle import dbb_oracle as O
    row = O.td_row_literal(_CPU['contigs'][i], _CPU['contigs'], _CPU['dist'], None)
    return int(row.sum())


def cpu_reference_sample(meta_small, budget_s, cores):
    """Times the oracle port on a bounded sample of the workload.  Returns dict with Gbp/s and pairs/s."""
    import multiprocessing as mp
    from oracle import dbb_oracle as O
    from dbb_b200 import synthetic
    from dbb_b200.data import load_td_dist
    # --- calibrate single-core rates on tiny pieces
    ids = np.arange(meta_small.n_scaffolds)
    seqs = [s.tobytes().decode('ascii') for s in synthetic.generate_sequences(meta_small, ids)]
    t = time.perf_counter(); O.seq_counts_literal(seqs[0][:100000]); r_bp = min(len(seqs[0]), 100000) / (time.perf_counter() - t)
    want_bp = r_bp * cores * budget_s
    take, acc = [], 0
    for i, s in enumerate(seqs):
        take.append(i); acc += len(s)
        if acc >= want_bp:
            break
    _CPU['seqs'] = seqs
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        t = time.perf_counter()
        out = pool.map(_cpu_count_worker, take, chunksize=1)
        t_count = time.perf_counter() - t
    bp = sum(o[1] for o in out)
    # --- TD neighbour rows (tdNeighbours.py:41-51) over all N columns of the sample set
    n = meta_small.n_scaffolds
    with np.errstate(invalid='ignore', divide='ignore'):
        sigs = [O.seq_signature_np(s) for s in seqs]
    contigs = [O.Contig('c%d' % i, int(meta_small.length[i]), 0.5, 1.0, sigs[i]) for i in range(n)]
    dist = O.Distributions(None, None, load_td_dist(), TD_PER, None, None)
    t = time.perf_counter(); O.td_row_literal(contigs[0], contigs[:200], dist, None); r_pair = 200 / (time.perf_counter() - t)
    n_rows = int(max(cores, min(n, r_pair * cores * budget_s / n)))
    _CPU['contigs'], _CPU['dist'] = contigs, dist
    with ctx.Pool(cores) as pool:
        t = time.perf_counter()
        pool.map(_cpu_row_worker, list(range(n_rows)), chunksize=1)
        t_rows = time.perf_counter() - t
    return {'gbp_s': bp / t_count / 1e9, 'pairs_s': n_rows * n / t_rows, 'sample_bp': bp, 'sample_rows': n_rows,
            'sample_cols': n, 't_count_s': t_count, 't_rows_s': t_rows}


def _cpu_np_count_worker(i):
    from oracle import dbb_oracle as O
    s = _CPU['seqs_b'][i]
    return int(O.seq_counts_np(s).sum()), len(s)


def _cpu_cdist_worker(lohi):
    from scipy.spatial.distance import cdist
    lo, hi = lohi
    d = cdist(_CPU['sig'][lo:hi], _CPU['sig'], 'cityblock')
    np.argpartition(d, K_NN, axis=1)[:, :K_NN]
    return hi - lo


def cpu_best_effort_sample(meta_small, cores):
    """vectorised CPU comparator (NOT the reference's algorithm): numpy rolling-code bincount for the
    counts, scipy cdist('cityblock') + argpartition for k nearest, fork-join over all cores, same bounded sample."""
    import multiprocessing as mp
    from oracle import dbb_oracle as O
    from dbb_b200 import synthetic
    seqs = synthetic.generate_sequences(meta_small)
    _CPU['seqs_b'] = seqs
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        t = time.perf_counter()
        out = pool.map(_cpu_np_count_worker, range(len(seqs)), chunksize=8)
        t_count = time.perf_counter() - t
    with np.errstate(invalid='ignore', divide='ignore'):
        _CPU['sig'] = np.array([O.seq_signature_np(s) for s in seqs])
    n = len(seqs)
    step = max(1, n // (cores * 4))
    with ctx.Pool(cores) as pool:
        t = time.perf_counter()
        pool.map(_cpu_cdist_worker, [(lo, min(n, lo + step)) for lo in range(0, n, step)], chunksize=1)
        t_rows = time.perf_counter() - t
    return {'gbp_s': sum(o[1] for o in out) / t_count / 1e9, 'pairs_s': n * n / t_rows, 'sample_bp': sum(o[1] for o in out), 'sample_n': n}


def cpu_sample_meta(scale):
    """The bounded CPU sample: the first 4 000 scaffolds' worth of the C2 law (same generator, same
    length law; all-pairs columns = that set)."""
    from dbb_b200 import synthetic
    kw = dict(synthetic.CONFIGS['C2'])
    kw['n_scaffolds'] = max(64, int(4000 * min(1.0, scale * 4)))
    kw['n_genomes'] = 20
    return synthetic.make_metagenome(**kw)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    meta_small = cpu_sample_meta(args.scale)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_sample(meta_small, 1.0, cores)
    vals = [cpu_reference_sample(meta_small, args.cpu_seconds, cores) for _ in range(max(1, min(args.steps, 3)))]
    gbp = float(np.mean([v['gbp_s'] for v in vals])); pairs = float(np.mean([v['pairs_s'] for v in vals]))
    sample = ('oracle port (literal restatement of genomicSignatures.py:134-150 and tdNeighbours.py:41-51), fork-join over %d '
              'processes; per step %.1f Mbp counted and %d rows x %d columns of TD pairs' %
              (cores, vals[-1]['sample_bp'] / 1e6, vals[-1]['sample_rows'], vals[-1]['sample_cols']))
    line = {'impl': 'reference', 'metric': METRIC, 'value': gbp, 'unit': 'Gbp/s', 'value2': pairs, 'unit2': 'scaffold-pairs/s',
            'n_gpus': args.gpus, 'steps': len(vals), 'warmup': min(args.warmup, 1),
            'ms_per_step': 1e3 * float(np.mean([v['t_count_s'] + v['t_rows_s'] for v in vals])),
            'higher_is_better': True, 'scaling': 'weak' if args.workload == 'C2' else 'strong', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args, max(int(args.gpus), 1)),       # the GPU arm's config at this N
            'cpu_baseline': {'value': gbp, 'unit': 'Gbp/s', 'value2': pairs, 'unit2': 'scaffold-pairs/s', 'cores': cores,
                             'kind': 'port', 'sample': sample},
            'e2e': {'value': gbp, 'unit': 'Gbp/s', 'value2': pairs, 'unit2': 'scaffold-pairs/s',
                    'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def workload_config(args, world):
    if getattr(args, 'workload', 'C2') == 'C5':
        return {'workload': 'C5: %d contigs of 1-2 kb + 8 scaffolds of 5 Mb (~15 Gbp x %.2f) in total over %d GPU(s) (strong scaling), '
                            'tetramer signatures only (skewed-length sharding / load-balance stress)' % (int(10000000 * args.scale), args.scale, world),
                'scale': args.scale, 'l2_policy': 'inputs larger than L2', 'parallelism': 'scaffolds LPT-sharded over the GPUs, no collective'}
    if getattr(args, 'workload', 'C2') != 'C2':
        return {'workload': '%s: %d scaffolds in total over %d GPU(s) (strong scaling), tetramer signatures + k=%d TD kNN + %s '
                            'neighbourhood (per=%d)' % (args.workload, int((1000000 if args.workload == 'C4' else 250000) * args.scale), world, K_NN,
                                                        'TD+GC+coverage' if args.workload == 'C3' else 'TD', TD_PER),
                'k': K_NN, 'td_percentile': TD_PER, 'scale': args.scale,
                'l2_policy': 'inputs larger than L2', 'parallelism': 'scaffolds LPT-sharded for counting, 1 all-gather (NCCL), kNN row tiles dealt to the ranks by work (equal row blocks with --no-prune)'}
    return {'workload': 'C2 per GPU: %d scaffolds x %d GPU(s), 1-100 kb log-uniform (~1.07 Gbp per GPU), tetramer signatures + '
                        'k=%d TD kNN + TD neighbourhood (per=%d) over all %d scaffolds'
                        % (int(50000 * args.scale), world, K_NN, TD_PER, int(50000 * args.scale) * world),
            'k': K_NN, 'td_percentile': TD_PER, 'scale': args.scale,
            'l2_policy': 'inputs larger than L2 (packed sequence ~400 MB per GPU re-read every step; stage 2 is FP32-bound)',
            'parallelism': 'scaffolds LPT-sharded for counting, 1 all-gather (NCCL), kNN row tiles dealt to the ranks by work (equal row blocks with --no-prune)' if world > 1 else 'single GPU'}


def run_ours(args):
    import torch
    from dbb_b200 import _lib, device, parallel, synthetic
    from dbb_b200.distributions import PairTables
    from dbb_b200.data import load_td_dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; this repo has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- workload ----------------
    meta, kw = workload(world, args.scale, args.workload)
    parts = parallel.lpt_partition(meta.length, world)
    order, _ = parallel.gathered_order(parts)
    mine = parts[rank]
    my_meta = sub_meta(meta, mine, 1000003 * rank)
    n_total = meta.n_scaffolds
    counting_only = args.workload == 'C5'
    if counting_only:
        args.no_e2e = True                           # 15 GB of pinned host ASCII is not worth holding for a stress config
    d_ascii, seq_off = device.synth_ascii(my_meta, dev)
    packed = device.pack_ascii(d_ascii, seq_off)
    plan = device.CountPlan(packed)
    my_n = my_meta.n_scaffolds
    my_bp = int(my_meta.length.sum())
    total_bp = int(meta.length.sum())
    counts = torch.zeros((my_n, 136), dtype=torch.int32, device=dev)
    freq = torch.zeros((my_n, _lib.SIG_STRIDE), dtype=torch.float32, device=dev)
    # K1 -> K2 hand-off of the per-scaffold scalars: a rank contributes its own shard's (id, length, GC, coverage) and
    # builds the tables of stage 2 from the gathered arrays (single GPU: its own arrays)
    g_ids, g_len, g_gc, g_cov = order, meta.length[order], meta.gc_obs[order], meta.coverage[order]
    if world > 1 and not counting_only:
        g_ids, g_len, g_gc, g_cov = parallel.all_gather_scalars(mine, my_meta.length, my_meta.gc_obs, my_meta.coverage,
                                                                counts=[len(p) for p in parts], device=dev)
        assert np.array_equal(g_ids, order)
    if counting_only:
        tables = dt = None
        topk_filter = 0
    elif args.workload == 'C3':      # BASELINE.json configs[2]: all three predicates; top-k restricted by the GC and coverage filters
        tables = PairTables(g_len, GC=g_gc, coverage=g_cov, tdDist=load_td_dist(),
                            tdDistPer=TD_PER, gcDist=synthetic.synthetic_gc_dist(), gcDistPer=TD_PER,
                            covDist=synthetic.synthetic_cov_dist(seed=3), covDistPer=TD_PER)
        topk_filter = device.FILTER_GC | device.FILTER_COV
    else:
        tables = PairTables(g_len, tdDist=load_td_dist(), tdDistPer=TD_PER)
        topk_filter = 0
    if not counting_only:
        dt = device.DevicePairTables(tables, dev)
    shard_counts = [len(p) for p in parts]
    r0, r1 = parallel.row_block(n_total, rank, world)
    torch.cuda.synchronize()

    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    result = {}
    sig_keep = [None]
    phase_events, timed_now = [], [False]

    def stage2(sig_all):
        # [r0, r1) is this rank's row block: of the caller's order for the full sweep, of the sorted order for the
        # pruned one (every rank derives the same sort from the gathered signatures)
        if args.no_prune:
            return device.pairs(sig_all, dt, r0, r1, k=K_NN, topk_filter=topk_filter, want_count=True)
        return device.pairs_pruned(sig_all, dt, k=K_NN, topk_filter=topk_filter, want_count=True,
                                   timings=phase_events if timed_now[0] else None, part=(rank, world) if world > 1 else None)

    def step(e=None):
        if e: e[0].record()
        plan.run(counts, freq)
        if e: e[1].record()
        if not counting_only:
            sig_all = parallel.all_gather_rows(freq, shard_counts)
            sig_keep[0] = sig_all
        if e: e[2].record()
        if not counting_only:
            result['knn'] = stage2(sig_all)
        if e: e[3].record()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local); sampler.start()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    timed_now[0] = True
    for s in range(args.steps):
        step(ev[s])
    timed_now[0] = False
    t_end.record()
    torch.cuda.synchronize(); barrier()
    clocks = sampler.stop()
    total_ms = max_over_ranks(t_start.elapsed_time(t_end))
    my_ms_count = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    ms_count = max_over_ranks(my_ms_count)
    per_rank = [[my_ms_count, float(my_bp)]]
    if dist is not None:
        t = torch.zeros((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(t, torch.tensor([[my_ms_count, float(my_bp)]], dtype=torch.float64, device=dev))
        per_rank = t.cpu().tolist()
    ms_gather = max_over_ranks(sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps)
    ms_knn = max_over_ranks(sum(e[2].elapsed_time(e[3]) for e in ev) / args.steps)
    ms_step = total_ms / args.steps
    pairs_total = float(n_total) * float(n_total)
    if counting_only:
        launches_per_step = plan.launches
    elif args.no_prune:
        launches_per_step = plan.launches + int(_lib.lib().dbb_pairs_launches(n_total, r1 - r0))
    else:
        launches_per_step = plan.launches + int(result['knn']['launches'])     # pairs + merge kernels of one or two sweeps

    e2e_s = e2e_count_s = e2e_page_s = e2e_list_s = None
    pageable_ok = True
    h2d_ceiling_gbs = None
    h2d = d2h = 0
    rows = r1 - r0
    if not args.no_e2e:
        # ---------------- end-to-end through the host-buffer C ABI (H2D + D2H inside the timed region) --------
        h_ascii = torch.empty(d_ascii.shape, dtype=torch.uint8, pin_memory=True)
        h_ascii.copy_(d_ascii); torch.cuda.synchronize()
        a_np = h_ascii.numpy()
        h_counts = torch.empty((my_n, 136), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        h_freq = torch.empty((my_n, 136), dtype=torch.float32, pin_memory=True).numpy()
        L = _lib.lib()
        from dbb_b200 import _pairs

        def e2e_count():
            _lib.check(L.dbb_tetra_signatures_host(_lib.ptr(a_np), _lib.ptr(seq_off), my_n, _lib.ptr(h_counts), _lib.ptr(h_freq)))

        def e2e_pairs():
            if world == 1:
                # host buffers in, host buffers out: the call TdNeighbours.knn makes
                return _pairs.run_knn_host(h_freq, tables, k=K_NN, topk_filter=topk_filter, prune=0 if args.no_prune else 1)
            # the host-produced signatures go up (and, N > 1, are gathered through the device), the row block runs,
            # its results are copied back to the host
            f = torch.from_numpy(h_freq).to(dev, non_blocking=True)
            sig_all = parallel.all_gather_rows(device.pad_signatures(f), shard_counts)
            res = stage2(sig_all)
            return {k_: v.cpu() for k_, v in res.items() if torch.is_tensor(v)}

        e2e_steps = max(1, min(args.steps, 5))
        e2e_count(); e2e_pairs()
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        tc = 0.0
        for _ in range(e2e_steps):
            t1 = time.perf_counter()
            e2e_count()
            tc += time.perf_counter() - t1
            e2e_pairs()
        torch.cuda.synchronize(); barrier()
        e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        e2e_count_s = max_over_ranks(tc / e2e_steps)
        # how the last stage-1 call moved its input: ASCII by the copy engine / planes packed by host threads
        import ctypes
        up_ascii, up_planes, up_thr = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        e2e_count()
        _lib.check(L.dbb_host_upload_stats(ctypes.byref(up_ascii), ctypes.byref(up_planes), ctypes.byref(up_thr)))
        # the same stage-1 call with PAGEABLE numpy buffers (what GenomicSignatures.signatures hands over): the library
        # stages them through its own pinned ring (csrc/host_api.cu::h2d_any)
        a_page = np.array(a_np)
        pc, pf = np.empty((my_n, 136), dtype=np.uint32), np.empty((my_n, 136), dtype=np.float32)
        _lib.check(L.dbb_tetra_signatures_host(_lib.ptr(a_page), _lib.ptr(seq_off), my_n, _lib.ptr(pc), _lib.ptr(pf)))
        barrier()
        t1 = time.perf_counter()
        for _ in range(3):
            _lib.check(L.dbb_tetra_signatures_host(_lib.ptr(a_page), _lib.ptr(seq_off), my_n, _lib.ptr(pc), _lib.ptr(pf)))
        e2e_page_s = max_over_ranks((time.perf_counter() - t1) / 3)
        pageable_ok = bool(np.array_equal(pc, h_counts))
        # the reference-facing CLASS call on a list of separate buffers (how the reference holds its sequences, one Python
        # object per scaffold): GenomicSignatures.signatures -> dbb_tetra_signatures_host_ptrs, nothing concatenated; the
        # time includes the Python loop that collects the 50 000 pointers
        from dbb_b200.genomicSignatures import GenomicSignatures
        gs = GenomicSignatures(4, 1)
        seq_list = [a_page[int(seq_off[i]):int(seq_off[i + 1])] for i in range(my_n)]
        lc, _ = gs.signatures(seq_list, want_freq=False)
        barrier()
        t1 = time.perf_counter()
        for _ in range(3):
            lc, _ = gs.signatures(seq_list, want_freq=False)
        e2e_list_s = max_over_ranks((time.perf_counter() - t1) / 3)
        pageable_ok = pageable_ok and bool(np.array_equal(lc, h_counts))
        del a_page, seq_list
        # the box's concurrent host-to-device ceiling: every rank copies its pinned ASCII buffer at the same time,
        # nothing else runs (the e2e line above is bound by this link, not by the kernels)
        d_tmp = torch.empty_like(d_ascii)
        d_tmp.copy_(h_ascii, non_blocking=True); torch.cuda.synchronize(); barrier()
        t1 = time.perf_counter()
        for _ in range(3):
            d_tmp.copy_(h_ascii, non_blocking=True)
        torch.cuda.synchronize(); barrier()
        h2d_ceiling_gbs = float(total_bp) * 3 / max_over_ranks(time.perf_counter() - t1) / 1e9
        del d_tmp
        rows = r1 - r0
        # bytes that crossed the link per step (stage 1: ASCII route + packed route + offsets; stage 2: signatures + tables)
        h2d = int(up_ascii.value + up_planes.value + seq_off.nbytes * 2 + (n_total * 136 * 4 + n_total * 8 if world == 1 else my_n * 136 * 4))
        d2h = int(my_n * 136 * 8 + rows * (K_NN * 8 + 12))

    # ---------------- parity spot check (untimed): sampled scaffolds / rows against the oracle ----------
    verified = None
    oracle_check = None
    if rank == 0:
        from oracle import dbb_oracle as O
        rng = np.random.RandomState(0)
        sample = rng.choice(my_n, size=min(16, my_n), replace=False)
        hc = counts.cpu().numpy().astype(np.int64)
        ok = True
        if not args.no_e2e:                       # every scaffold of the shard, exactly, with the C restatement
            from oracle import c_oracle
            ok &= bool(np.array_equal(hc, c_oracle.counts_batch(a_np, seq_off)))
        for i in sample:
            s = d_ascii[int(seq_off[i]):int(seq_off[i + 1])].cpu().numpy()
            ok &= bool(np.array_equal(hc[i], O.seq_counts_np(s)))
        if not args.no_e2e:
            ok &= bool(np.array_equal(h_counts.astype(np.int64), hc)) and pageable_ok
        if not counting_only:
            # stage 2 against the ORACLE at this size (every N, every workload): a seeded sample of this rank's rows,
            # float64 distances to all n_total columns by the C restatement of tdNeighbours.py:41-51, predicates by the
            # numpy restatement; top-k sets within the 2e-6 tie band, count / neighbour_bp exact outside it
            from oracle import checks
            kn = result['knn']
            ids = kn['row_ids'].cpu().numpy() if 'row_ids' in kn else np.arange(r0, r1)
            pos = np.sort(np.random.RandomState(1).choice(ids.shape[0], size=min(args.check_rows, ids.shape[0]), replace=False))
            tpos = torch.from_numpy(pos).to(dev)
            sig32 = sig_keep[0][:, :136].cpu().numpy()
            c3 = args.workload == 'C3'
            ot = O.DistTables(synthetic.synthetic_gc_dist() if c3 else None, TD_PER if c3 else None, load_td_dist(), TD_PER,
                              synthetic.synthetic_cov_dist(seed=3) if c3 else None, TD_PER if c3 else None)
            try:
                oracle_check = checks.check_knn_sample(
                    ids[pos], kn['idx'][tpos].cpu().numpy(), kn['dist'][tpos].cpu().numpy(), kn['count'][tpos].cpu().numpy(),
                    kn['neighbour_bp'][tpos].cpu().numpy(), sig32, meta.length[order], K_NN, ot,
                    GC=meta.gc_obs[order] if c3 else None, coverage=meta.coverage[order] if c3 else None, topk_filter_gc_cov=c3)
            except AssertionError as exc:
                raise SystemExit('bench.py: stage-2 parity check against the oracle FAILED: %s' % exc)
            del sig32
            if not args.no_prune and n_total <= 300000:
                # additionally (cheap at these sizes): the pruned sweep against the full sweep, identical bit for bit
                full = device.pairs(sig_keep[0], dt, 0, n_total, k=K_NN, topk_filter=topk_filter, want_count=True)
                tid = kn['row_ids']
                for key in ('idx', 'dist', 'count', 'neighbour_bp'):
                    ok &= bool(torch.equal(kn[key], full[key][tid]))
        verified = ok
        if not ok:
            raise SystemExit('bench.py: parity spot check FAILED (counts differ from the oracle)')

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline ----------------
    hbm_peak, sm_max_mhz, peak_src = load_peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    alg_bytes = packed.algorithmic_bytes()          # rank 0's shard: bytes one stage-1 launch set moves
    traffic = load_traffic()
    gbs = alg_bytes / (ms_count * 1e-3) / 1e9
    fp32_peak = sm_count * 128 * sm_max_mhz * 1e6
    pruning = None
    pairs_executed = float(rows) * float(n_total)
    if not counting_only and not args.no_prune:
        kn = result['knn']
        pruning = {'tiles_visited': int(kn['tiles_visited']), 'tiles_total': int(kn['tiles_total']),
                   'frac_visited': kn['tiles_visited'] / max(kn['tiles_total'], 1),
                   'what': 'exact: scaffolds sorted by (TD-bound class, GC projection p of the signature); L1 >= 2|p_a - p_b| lets whole '
                           '128 x 128 tiles be skipped; results identical to the full sweep (checked below)'}
        rows = int(kn['row_ids'].shape[0])
        pairs_executed = min(float(rows) * float(n_total), kn['tiles_visited'] * 128.0 * 128.0)
    ms_kernels = ms_knn
    if phase_events:
        # pairs_kernel + merge_kernel launches only (the sweeps), without the planning ops around them
        per_step = phase_events                     # one dict of device times per timed step (dbb_knn_output.ms_*)
        ms_plan = sum(d['plan'] for d in per_step) / len(per_step)
        ms_sweep = sum(d['sweep'] for d in per_step) / len(per_step)
        ms_second = sum(d['second sweep + merge'] for d in per_step) / len(per_step)
        ms_kernels = ms_sweep + (ms_second if pruning and int(result['knn']['launches']) > 2 else 0.0)
        pruning.update({'ms_plan': ms_plan, 'ms_sweep': ms_sweep, 'ms_second_sweep_and_merge': ms_second})
    lane_ops = 272.0 * pairs_executed / (ms_kernels * 1e-3)        # executed work only (rank 0's row block)
    lane_ops_resolved = 272.0 * float(rows) * float(n_total) / (ms_knn * 1e-3)
    line = {
        'metric': METRIC,
        'value': total_bp / (ms_count * 1e-3) / 1e9, 'unit': 'Gbp/s',
        'value2': None if counting_only else pairs_total / (ms_knn * 1e-3), 'unit2': 'scaffold-pairs/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_step,
        'stages_ms': {'signatures': ms_count, 'allgather': None if counting_only else ms_gather, 'knn': None if counting_only else ms_knn},
        'pipeline': {'gbp_per_s': total_bp / (ms_step * 1e-3) / 1e9, 'pairs_per_s': None if counting_only else pairs_total / (ms_step * 1e-3)},
        'higher_is_better': True, 'scaling': 'weak' if args.workload == 'C2' else 'strong', 'vs_baseline': None, 'dtype': 'u32 counts / f32 distances',
        'data': 'synthetic', 'config': workload_config(args, world),
        'roofline': {'kernel': 'count_kernel<S> (stage 1, all classes of one launch set)', 'bound': 'hbm', 'achieved': gbs,
                     'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak, 'traffic': traffic.get('count_dram_bytes_per_step') if args.workload == 'C2' and args.scale == 1.0 else None,
                     'algorithmic_bytes': alg_bytes, 'peak_source': peak_src},
        'roofline_pairs': None if counting_only else {'kernel': 'pairs_kernel (stage 2)', 'bound': 'fp32', 'achieved': lane_ops / 1e12,
                           'peak': fp32_peak / 1e12, 'unit': 'T lane-op/s', 'frac': lane_ops / fp32_peak,
                           'pairs_resolved_per_executed': float(rows) * float(n_total) / max(pairs_executed, 1.0),
                           'frac_resolved': lane_ops_resolved / fp32_peak,
                           'note': 'achieved / frac: lane-ops of the pairs actually computed over the time of the sweep launches; frac_resolved: '
                                   '272 lane-ops for every pair of the N x N problem over the whole stage (planning included), i.e. what a full sweep '
                                   'would have to sustain to be as fast',
                           'traffic': traffic.get('pairs_dram_bytes_per_step' if args.no_prune else 'pairs_pruned_dram_bytes_per_launch') if args.workload == 'C2' and args.scale == 1.0 and world == 1 else None,
                           'peak_source': '%d SMs x 128 FP32 lanes x %.0f MHz (272 lane-ops per pair)' % (sm_count, sm_max_mhz)},
        'e2e': None if args.no_e2e else {'value': total_bp / e2e_count_s / 1e9, 'unit': 'Gbp/s', 'value2': pairs_total / max(e2e_s - e2e_count_s, 1e-9),
                'unit2': 'scaffold-pairs/s', 'ms_per_step': e2e_s * 1e3, 'steps': e2e_steps,
                'value_pageable_host_buffers': total_bp / e2e_page_s / 1e9, 'value_class_api_list_of_buffers': total_bp / e2e_list_s / 1e9,
                'h2d_concurrent_ceiling_gbs': h2d_ceiling_gbs,
                'frac_of_h2d_ceiling': (total_bp / e2e_count_s / 1e9) / h2d_ceiling_gbs, 'pipeline_gbp_per_s': total_bp / e2e_s / 1e9, 'pipeline_pairs_per_s': pairs_total / e2e_s,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'host_input_bytes_per_step': int(a_np.nbytes),
                'upload': {'ascii_route_bytes': int(up_ascii.value), 'host_packed_route_bytes': int(up_planes.value),
                           'host_pack_threads': int(up_thr.value),
                           'note': 'dbb_tetra_signatures_host sends part of every batch as ASCII (copy engine) and part as planes packed by host threads (0.375 B/base); rank 0 values'},
                'api': 'dbb_tetra_signatures_host + dbb_pairs_host (pinned host buffers)' if args.no_prune else
                       'dbb_tetra_signatures_host + dbb_knn_host (pinned host buffers)' if world == 1 else 'dbb_tetra_signatures_host (pinned host buffers) + H2D of the signatures, all-gather, dbb_knn_dev, D2H of the results'},
        'pruning': pruning,
        'gpu_launches': launches_per_step * args.steps, 'launches_per_step': launches_per_step,
        'clocks': clocks, 'verified_vs_oracle': verified, 'oracle_check_stage2': oracle_check,
        'shard': {'scaffolds_rank0': my_n, 'bp_rank0': my_bp, 'lpt_imbalance': parallel.imbalance(meta.length, parts),
                  'count_ms_per_rank': [round(r[0], 4) for r in per_rank], 'bp_per_rank': [int(r[1]) for r in per_rank],
                  'count_time_imbalance': max(r[0] for r in per_rank) / (sum(r[0] for r in per_rank) / len(per_rank))},
    }
    if world == 1 and not args.no_cpu and not counting_only:
        cores = os.cpu_count() or 1
        c = cpu_reference_sample(cpu_sample_meta(args.scale), args.cpu_seconds, cores)
        line['cpu_baseline'] = {
            'value': c['gbp_s'], 'unit': 'Gbp/s', 'value2': c['pairs_s'], 'unit2': 'scaffold-pairs/s', 'cores': cores, 'kind': 'port',
            'sample': 'oracle port of genomicSignatures.py:134-150 / tdNeighbours.py:41-51, fork-join over %d processes: %.1f Mbp '
                      'counted in %.1f s, %d rows x %d columns of TD pairs in %.1f s (same generator and length law as the GPU '
                      'workload)' % (cores, c['sample_bp'] / 1e6, c['t_count_s'], c['sample_rows'], c['sample_cols'], c['t_rows_s'])}
        b = cpu_best_effort_sample(cpu_sample_meta(args.scale), cores)
        line['cpu_best_effort'] = {
            'value': b['gbp_s'], 'unit': 'Gbp/s', 'value2': b['pairs_s'], 'unit2': 'scaffold-pairs/s', 'cores': cores, 'kind': 'vectorised numpy/scipy '
            'comparator, not the reference algorithm', 'sample': '%.1f Mbp counted, %d x %d cdist(cityblock) + argpartition' % (b['sample_bp'] / 1e6, b['sample_n'], b['sample_n'])}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scale', type=float, default=1.0, help='fraction of the C2 scaffold count (tests only; 1.0 = the named config)')
    ap.add_argument('--cpu-seconds', type=float, default=8.0, help='target wall time of each CPU-baseline leg')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline leg')
    ap.add_argument('--no-prune', action='store_true', help='stage 2 as a full sweep over all column tiles (default: exact tile skipping, device.pairs_pruned)')
    ap.add_argument('--check-rows', type=int, default=96, help='rows of the all-pairs stage recomputed by the oracle after the timed region')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer end-to-end leg (large exploratory workloads)')
    ap.add_argument('--workload', default='C2', choices=['C2', 'C3', 'C4', 'C5'], help='C2 = the benchmark of record (weak scaling); C3 = 250k scaffolds with GC + coverage filters, C4 = 1M scaffolds, C5 = 10M short contigs + 8 x 5 Mb, counting only (all strong scaling)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
