#!/usr/bin/env python
"""bench.py -- search-step throughput of the BM-NAS fusion-cell hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config ntu|mmimdb|ego]

One "step" = one search step (arch step on a dev batch + weight step on a train batch,
train_searchable/ntu.py:70-93 + architect.py:21-29) over synthetic frozen-backbone
features.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the exact
definitions of value / e2e / roofline / cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200'))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # SURVEY 8 / App. B defaults
    'ntu': dict(C=128, L=8, num_input_nodes=8, steps=2, multiplier=2, node_steps=2, node_multiplier=2, drpt=0.2,
                B=96, classes=60, loss='ce', eta_max=1e-3, weight_decay=3e-4),
    'mmimdb': dict(C=192, L=16, num_input_nodes=6, steps=2, multiplier=2, node_steps=1, node_multiplier=1, drpt=0.1,
                   B=32, classes=23, loss='bce', eta_max=1e-3, weight_decay=1e-4),
    'ego': dict(C=128, L=8, num_input_nodes=8, steps=2, multiplier=2, node_steps=3, node_multiplier=3, drpt=0.05,
                B=96, classes=83, loss='ce', eta_max=3e-3, weight_decay=1e-4),
    # BASELINE configs[3]: "larger C/L and more Step nodes" (the authors' C=256, steps=4, multiplier=4 exploration,
    # structure_vis.ipynb cell 3)
    'ego_large': dict(C=256, L=16, num_input_nodes=8, steps=4, multiplier=4, node_steps=3, node_multiplier=3, drpt=0.05,
                      B=96, classes=83, loss='ce', eta_max=3e-3, weight_decay=1e-4),
}
L2_BYTES = 126 * 1024 * 1024


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sustained=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src='fallback')


# ----------------------------------------------------------------------------- CPU reference arm (oracle port)
def cpu_reference(cfgname, steps, warmup, max_seconds=25.0, device='cpu'):
    """times oracle/bmnas_oracle.py (the CPU restatement of the reference's path, pinned against the
    reference by tests/golden) on all host cores.  Returns (samples_per_s, ms_per_step, cores, steps_done).
    device='cuda': the same functional PyTorch code run eagerly on the GPU (cuDNN/cuBLAS kernels, TF32 off) --
    the "reference-style eager PyTorch on B200" baseline SURVEY 8d asks for next to the CPU number."""
    from oracle import bmnas_oracle as O
    c = CONFIGS[cfgname]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gpu = device != 'cpu'
    if gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    cfg = O.Cfg(c['C'], c['L'], c['num_input_nodes'], c['steps'], c['multiplier'], c['node_steps'],
                c['node_multiplier'], c['drpt'])
    P = O.init_params(cfg, c['classes'], seed=2, prefix='cell')
    arch = O.init_arch(cfg, seed=2)
    st = O.SearchState(cfg, P, arch, eta_max=c['eta_max'], weight_decay=c['weight_decay'], loss=c['loss'])
    dev = O.synthetic_batch(cfg, c['B'], c['classes'], seed=2, loss=c['loss'])
    trn = O.synthetic_batch(cfg, c['B'], c['classes'], seed=3, loss=c['loss'])
    g = torch.Generator().manual_seed(7)
    if gpu:
        for k in list(P):
            P[k] = P[k].to(device)
        st.arch = [a.to(device) for a in st.arch]
        dev = ([f.to(device) for f in dev[0]], dev[1].to(device))
        trn = ([f.to(device) for f in trn[0]], trn[1].to(device))
        g = torch.Generator(device=device).manual_seed(7)

    def masks():
        # the reference draws fresh dropout masks every forward; do the same work here
        m = {}
        for name in _dropout_sites(cfg):
            p = 0.1 if name.endswith('_ops.1.dropout') else cfg.drpt
            m['fusion_net.' + name] = (torch.rand(c['B'], cfg.C, cfg.L, generator=g, device=device) >= p).to(torch.uint8)
        return m

    def sync():
        if gpu:
            torch.cuda.synchronize()
    for _ in range(warmup):
        st.search_step(dev, trn, masks(), masks())
    sync()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        st.search_step(dev, trn, masks(), masks())
        done += 1
        if gpu and done % 8 == 0:
            sync()
        if time.perf_counter() - t0 > max_seconds:
            break
    sync()
    dt = time.perf_counter() - t0
    return c['B'] * done / dt, 1e3 * dt / done, cores, done


def _dropout_sites(cfg):
    sites = []
    for i in range(cfg.steps):
        nc = f'cell._step_nodes.{i}.node_cell'
        for j in range(cfg.node_steps):
            for k in (1, 2, 3):
                sites.append(f'{nc}.node_ops.{j}._ops.{k}.dropout')
        if cfg.node_multiplier != 1:
            sites.append(f'{nc}.out_dropout')
    return sites


# ----------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------- our arm
def make_pool(c, n_batches, seed, device, pinned=False):
    """synthetic (B,C,L) unit-normal features + labels; one flat tensor per batch"""
    g = torch.Generator().manual_seed(seed)
    pool = []
    for _ in range(n_batches):
        f = torch.randn(c['num_input_nodes'], c['B'], c['C'], c['L'], generator=g)
        if c['loss'] == 'ce':
            y = torch.randint(0, c['classes'], (c['B'],), generator=g)
        else:
            y = (torch.rand(c['B'], c['classes'], generator=g) < 0.2).float()
        if pinned:
            pool.append((f.pin_memory(), y.pin_memory()))
        else:
            pool.append((f.to(device), y.to(device)))
    return pool


def graph_time_us(call, R=50, reps=4):
    """device time per launch of one prepared C-ABI call: R launches captured into a CUDA graph and replayed
    (CUDA events on the replay stream; no host launch overhead inside the number)"""
    import ctypes
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        for _ in range(3):
            call(sp)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for _ in range(R):
                call(sp)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * R)


def algorithmic_bytes(call, c, B):
    """HBM bytes one launch must move: tensors read once + tensors written once, fp32"""
    st, T = call.st, B * c['C'] * c['L'] * 4
    n = call.name
    if n == 'bmnas_mix_fwd':
        return (st.n + 1) * T
    if n == 'bmnas_mix_bwd':   # gout in; the x_j in when d(alpha) is wanted; the gx_j out (the two halves are separate launches)
        n_gx = sum(1 for j in range(st.n) if st.gx[j])
        return (1 + (st.n if st.gw else 0) + n_gx) * T
    if n == 'bmnas_node_fwd':
        return (1 if st.alias_xy else 2) * T + st.M * c['L'] * 4 * B + T
    if n == 'bmnas_node_bwd':
        return (2 if st.alias_xy else 3) * T + 2 * st.M * c['L'] * 4 * B + (1 if st.alias_xy else 2) * T
    if n == 'bmnas_conv_fwd':
        return (st.K + st.M) * c['L'] * 4 * B + st.K * st.M * 4
    if n == 'bmnas_conv_dgrad':
        return (2 * st.M + st.K) * c['L'] * 4 * B + st.K * st.M * 4
    if n == 'bmnas_conv_wgrad':
        return (2 * st.M + st.K) * c['L'] * 4 * B + st.K * st.M * 4
    if n == 'bmnas_ln_fwd':
        return (st.Ctot + (st.Ctot if st.mode == 1 else 0) + st.Ctot) * c['L'] * 4 * B
    if n == 'bmnas_ln_bwd':
        return 4 * st.Ctot * c['L'] * 4 * B
    return 0


def large_batch_roofline(args, c, pk, device, B):
    """the same kernels at a batch where they are bandwidth bound: per-kernel device time (graph-replayed
    launches), algorithmic bytes and fraction of the HBM peak; heaviest instance of each kernel"""
    import types
    from bmnas.nn import SearchHead, CrossEntropyLoss, BCEWithLogitsLoss
    from bmnas.search import SearchStep
    c = dict(c, B=B)
    a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps',
                                                    'node_multiplier', 'drpt')}, weight_decay=c['weight_decay'])
    crit = CrossEntropyLoss() if c['loss'] == 'ce' else BCEWithLogitsLoss()
    head = SearchHead(a, c['classes'], criterion=crit).to(device)
    ss = SearchStep(head, crit, B, c['classes'], loss_kind=c['loss'], use_graphs=False)
    pool = make_pool(c, 2, 77, device)
    ss.load('dev', *pool[0]); ss.load('train', *pool[1])
    for _ in range(2):
        ss.step()
    torch.cuda.synchronize()
    runner = [r for r in head.fusion_net._bm_cache.values() if r.prog.training][0]
    keep_gout = torch.zeros_like(runner.out)
    runner.prog.bind('gout', keep_gout)        # the upstream-gradient slot pointed at a freed autograd temporary
    best = {}
    for call in runner.prog.fwd + runner.prog.bwd:
        by = algorithmic_bytes(call, c, B)
        if by and (call.name not in best or by > best[call.name][1]):
            best[call.name] = (call, by)
    out = {'B': B, 'peak_GBs': pk['hbm'], 'peak_source': pk['src'], 'kernels': {}}
    for name, (call, by) in sorted(best.items()):
        us = graph_time_us(call, R=10, reps=3)
        gbs = by / (us * 1e-6) / 1e9
        out['kernels'][name] = {'us': round(us, 1), 'algorithmic_MB': round(by / 1e6, 1), 'GBs': round(gbs, 1),
                                'frac': round(gbs / pk['hbm'], 4)}
    del ss, head, pool
    torch.cuda.empty_cache()
    return out


def ncu_traffic(kernel, B):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture (profiles/)"""
    try:
        d = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        return d.get(f'{kernel}@B{B}')
    except Exception:
        return None


def kernel_roofline(head, ss, c, pk):
    """average duration of the dominant fused MixedOp kernels (bmnas_node_fwd / bmnas_node_bwd), timed live with
    CUDA events around graph-replayed back-to-back launches of the prepared parameter blocks."""
    from bmnas import native as N
    runner = [r for r in head.fusion_net._bm_cache.values() if r.prog.training][0]
    prog = runner.prog
    keep_gout = torch.zeros_like(runner.out)
    prog.bind('gout', keep_gout)
    res = {}
    for name, calls in (('bmnas_node_fwd', prog.fwd), ('bmnas_node_bwd', prog.bwd)):
        cand = [x for x in calls if x.name == name]
        plain = [x for x in cand if not (x.st.out2 or x.st.gout2)]   # an instance without a chained edge mix: the bytes below
        call = (plain or cand)[0]
        res[name] = graph_time_us(call)                # us per launch
    T1 = c['C'] * c['L'] * 4
    M = 3 * c['C']
    # algorithmic bytes per sample of the fused node kernel: x (aliased with y) + Z (3C rows) in, out
    fwd_bytes = c['B'] * (T1 + M * c['L'] * 4 + T1)
    t = res['bmnas_node_fwd'] * 1e-6
    ach = fwd_bytes / t / 1e9
    return {'bound': 'hbm', 'kernel': 'bmnas_node_fwd (fused NodeMixedOp forward)', 'achieved': round(ach, 2),
            'peak': pk['hbm'], 'peak_source': pk['src'], 'unit': 'GB/s', 'frac': round(ach / pk['hbm'], 5),
            'traffic': ncu_traffic('bmnas_node_fwd', c['B']), 'avg_launch_us': {k: round(v, 3) for k, v in res.items()},
            'algorithmic_bytes_per_launch': fwd_bytes,
            'note': 'B=%d working set is L2-resident and the kernel is latency-bound at this size (graph-replayed launches, '
                    'programmatic dependent launch overlaps the early section of launch i+1 with launch i); '
                    'roofline_large_batch times the same kernels where they are bandwidth bound' % c['B']}


def profile_kernels(head, R=50):
    """device time of every prepared call of the training plan: R back-to-back launches of the SAME call captured
    into one CUDA graph and replayed (no host launch overhead in the number; stream launches through ctypes are
    host-bound at ~3 us and hide anything shorter)"""
    from bmnas import native as N
    import ctypes
    runner = [r for r in head.fusion_net._bm_cache.values() if r.prog.training][0]
    prog = runner.prog
    keep_gout = torch.zeros_like(runner.out)
    prog.bind('gout', keep_gout)               # the upstream-gradient slot pointed at a freed autograd temporary
    rows = []
    for phase, calls in (('fwd', prog.fwd), ('bwd', prog.bwd)):
        for i, call in enumerate(calls):
            st = call.st
            dims = {k: getattr(st, k) for k in ('B', 'L', 'K', 'M', 'C', 'n', 'Ctot', 'n_src', 'w_fold', 'mode', 'n_ops')
                    if hasattr(st, k)}
            rows.append((phase, i, call.name, round(graph_time_us(call), 2), dims))
    return rows


def run_ours(args):
    import torch.distributed as dist
    from bmnas import native as N
    from bmnas.nn import SearchHead, CrossEntropyLoss, BCEWithLogitsLoss
    from bmnas.search import SearchStep
    import types
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
        group = dist.group.WORLD
    c = dict(CONFIGS[args.config])
    if args.batch:
        c['B'] = args.batch
    torch.manual_seed(2)                              # main_darts_searchable_ntu.py:17 (all ranks: identical replicas)
    a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps',
                                                    'node_multiplier', 'drpt')}, weight_decay=c['weight_decay'])
    crit = CrossEntropyLoss() if c['loss'] == 'ce' else BCEWithLogitsLoss()
    head = SearchHead(a, c['classes'], criterion=crit).to(device)
    from bmnas import runtime as rt
    rt.SAMPLE_OFFSET[0] = rank * c['B']               # world-size-invariant dropout streams
    ss = SearchStep(head, crit, c['B'], c['classes'], loss_kind=c['loss'], eta_max=c['eta_max'],
                    weight_decay=c['weight_decay'], nbpe=400.0, use_graphs=not args.no_graphs, group=group)
    # input pool larger than L2 so every step's inputs come from HBM
    per_batch = c['num_input_nodes'] * c['B'] * c['C'] * c['L'] * 4
    n_pool = max(4, int(1.3 * L2_BYTES / per_batch) + 1)
    n_pool += n_pool % 2
    pool = make_pool(c, n_pool, 100 + rank, device)
    ss.load('dev', *pool[0]); ss.load('train', *pool[1])
    ss.prepare(warmup=3, restore=False)
    n_weights = sum(p.numel() for p in head.parameters())
    n_arch = sum(p.numel() for p in head.arch_parameters())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K, loader, read_loss):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for i in range(K):
            loader(i)
            la, lw = ss.step()
            if read_loss:
                lw_host = lw.item()               # device->host read of the step's result
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        return e0.elapsed_time(e1), wall          # device time span of the K steps (idle gaps included)

    def load_dev(i):
        ss.load('dev', *pool[(2 * i) % n_pool]); ss.load('train', *pool[(2 * i + 1) % n_pool])
    for i in range(args.warmup):
        load_dev(i); ss.step()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    dev_ms, wall = timed(args.steps, load_dev, read_loss=False)
    # ---- end to end: pinned host inputs, H2D every step, D2H loss read every step
    hpool = make_pool(c, 8, 500 + rank, device, pinned=True)

    def load_host(i):
        ss.load('dev', *hpool[(2 * i) % 8]); ss.load('train', *hpool[(2 * i + 1) % 8])
    for i in range(max(3, args.warmup)):
        load_host(i); ss.step()
    serial_ms, serial_wall = timed(args.steps, load_host, read_loss=True)   # copy, then compute, then read: no overlap

    def timed_pipelined(K):
        """the input pipeline a training loop runs: SearchStep.prefetch() queues batch i+1 on the copy stream right
        after step i was launched, then the host reads step i's loss.  All K batches (step 0's included, which
        nothing can hide) are copied from pinned host memory inside the timed region."""
        barrier()
        t0 = time.perf_counter()
        ss.prefetch('dev', *hpool[0]); ss.prefetch('train', *hpool[1])
        for i in range(K):
            la, lw = ss.step()
            if i + 1 < K:
                ss.prefetch('dev', *hpool[(2 * i + 2) % 8]); ss.prefetch('train', *hpool[(2 * i + 3) % 8])
            lw_host = lw.item()
        barrier()
        return time.perf_counter() - t0
    timed_pipelined(max(3, args.warmup))
    e2e_wall = timed_pipelined(args.steps)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_wall * 1e3, serial_wall * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms_wall, serial_ms_wall = t.tolist()
    gB = c['B'] * world
    ms_per_step = dev_ms / args.steps
    value = gB / (ms_per_step * 1e-3)
    e2e_value = gB * args.steps / (e2e_ms_wall * 1e-3)
    lab_bytes = c['B'] * (8 if c['loss'] == 'ce' else 4 * c['classes'])
    out = None
    if rank == 0 and args.profile_kernels and ss.graphs:
        for which in ('dev', 'train'):
            g = ss.graphs[which]
            for _ in range(5):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            print('# graph replay (%s half step, back to back, no input copies): %.1f us' % (which, e0.elapsed_time(e1) * 10))
    if rank == 0 and args.profile_kernels:
        rows = profile_kernels(head)
        tot = sum(r[3] for r in rows)
        print('# warm per-launch time of the training plan (one fwd+bwd of the fusion network), total %.1f us' % tot)
        for r in rows:
            print('%s %3d %-18s %8.2f us  %s' % r)
    if rank == 0:
        pk = peaks()
        roof = kernel_roofline(head, ss, c, pk)
        big = None
        if world == 1 and args.roofline_batch > 0:
            big = large_batch_roofline(args, c, pk, device, args.roofline_batch)
        cpu = None
        if world == 1 and not args.no_cpu:
            v, ms, cores, done = cpu_reference(args.config, 40, 3, max_seconds=20.0)
            cpu = {'value': round(v, 1), 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'ms_per_step': round(ms, 2),
                   'sample': f'{done} search steps of the same workload (B={c["B"]}) on the oracle port, '
                             f'torch CPU fp32, {cores} threads'}
            try:    # the same functional PyTorch code, eager on this GPU (reported next to the CPU number)
                gv, gms, _, gdone = cpu_reference(args.config, 60, 3, max_seconds=10.0, device=str(device))
                cpu['torch_eager_gpu'] = {'value': round(gv, 1), 'unit': 'samples/s', 'ms_per_step': round(gms, 3),
                                          'sample': f'{gdone} search steps, eager PyTorch fp32 (TF32 off) on the same B200'}
            except Exception as e:
                cpu['torch_eager_gpu'] = {'unavailable': repr(e)[:200]}
        out = {
            'metric': 'search-step samples/sec (fwd+bwd+arch step, fwd+bwd+weight step)', 'value': round(value, 1),
            'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{args.config.upper()} fusion search step on synthetic frozen-backbone features, '
                                   f'B={c["B"]} per GPU, C={c["C"]}, L={c["L"]}, n_in={c["num_input_nodes"]}, '
                                   f'steps={c["steps"]}, node_steps={c["node_steps"]}, classes={c["classes"]}',
                       'global_batch': gB, 'weights': n_weights, 'arch_scalars': n_arch,
                       'parallelism': f'dp{world} (batch-sharded, one NCCL all-reduce of the flat grad arena per half step)',
                       'cuda_graphs': not args.no_graphs,
                       'l2_policy': f'inputs larger than L2: pool of {n_pool} distinct resident batches '
                                    f'({n_pool * per_batch / 2**20:.0f} MiB) rotated every step'},
            'e2e': {'value': round(e2e_value, 1), 'unit': 'samples/s',
                    'h2d_bytes_per_step': 2 * (per_batch + lab_bytes), 'd2h_bytes_per_step': 4,
                    'ms_per_step': round(e2e_ms_wall / args.steps, 4),
                    'how': 'SearchStep.prefetch() from pinned host memory (copy stream; batch i+1 travels while step i '
                           'computes) + SearchStep.step() + loss.item() every step, wall clock between barriers',
                    'serial_value': round(gB * args.steps / (serial_ms_wall * 1e-3), 1),
                    'serial_how': 'SearchStep.load() + step() + loss.item(): copy, compute and read strictly in sequence'},
            'gpu_launches': (ss.launches_per_step or 0) * args.steps,
            'launches_per_step': ss.launches_per_step,
            'roofline': roof, 'roofline_large_batch': big, 'cpu_baseline': cpu, 'clocks': clk,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        # the captured graphs hold NCCL kernels: tearing the process group down underneath them can hang, so
        # every rank synchronises, meets at one last barrier and leaves without the NCCL destructor
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    c = CONFIGS[args.config]
    v, ms, cores, done = cpu_reference(args.config, max(args.steps, 1), max(args.warmup, 1), max_seconds=120.0)
    print(json.dumps({
        'impl': 'reference', 'metric': 'search-step samples/sec (fwd+bwd+arch step, fwd+bwd+weight step)',
        'value': round(v, 1), 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': done, 'warmup': args.warmup,
        'ms_per_step': round(ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.config.upper()} fusion search step on synthetic frozen-backbone features, '
                               f'B={c["B"]}, C={c["C"]}, L={c["L"]}', 'global_batch': c['B']},
        'cpu_baseline': {'value': round(v, 1), 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{done} search steps on the oracle port (torch CPU fp32, {cores} threads)'},
        'e2e': {'value': round(v, 1), 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='ntu', choices=list(CONFIGS))
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch override')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--profile-kernels', action='store_true')
    ap.add_argument('--roofline-batch', type=int, default=8192,
                    help='also time the kernels at this per-GPU batch (bandwidth-bound regime); 0 = skip')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == '__main__':
    main()
