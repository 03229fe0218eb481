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


# ----------------------------------------------------------------------------- reference arm
def reference_arm(cfgname, steps, warmup, max_seconds, device='cpu', full_fidelity=False, batch=None):
    """The reference's own modules (FusionNetwork, Architect, LRCosineAnnealingScheduler, torch.optim.Adam, unmodified,
    vendored by tools/vendor_ref.sh into the git-ignored oracle/_ref/) driven through the reference's loop body by
    oracle/ref_harness.py.  Runs in THIS process: only call it from a process that never imported the product package
    (`bench.py --impl reference`; run_ours() reaches it through a subprocess).  Falls back to the oracle port (kind
    "port") when oracle/_ref is absent.  Returns (kind, samples/s, ms/step, threads, steps done)."""
    from oracle import ref_harness as H
    c = dict(CONFIGS[cfgname])
    if batch:
        c['B'] = batch
    if H.available():
        v, ms, thr, done = H.time_search(c, steps, warmup, max_seconds, device=device, full_fidelity=full_fidelity)
        return 'reference', v, ms, thr, done
    v, ms, thr, done = cpu_reference(cfgname, steps, warmup, max_seconds=max_seconds, device=device)
    return 'port', v, ms, thr, done


def reference_subprocess(cfgname, steps, warmup, max_seconds, device='cpu', full_fidelity=False, batch=None):
    """run `bench.py --impl reference` in a child process and return its JSON line (dict) or {'unavailable': why}"""
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', cfgname, '--steps', str(steps),
           '--warmup', str(warmup), '--max-seconds', str(max_seconds), '--ref-device', device]
    if full_fidelity:
        cmd.append('--full-fidelity')
    if batch:
        cmd += ['--batch', str(batch)]
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE', 'MASTER_ADDR', 'MASTER_PORT')}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=max_seconds * 3 + 180, env=env)
        lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
        return json.loads(lines[-1]) if lines else {'unavailable': (out.stderr or 'no output')[-300:]}
    except Exception as e:
        return {'unavailable': repr(e)[:300]}


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
    if n == 'bmnas_mixed_fwd':     # SURVEY 8(d) node_mixed fwd, aliased inputs: t in, out out, weights once per launch
        return 2 * T + 6 * c['C'] * c['C'] * 4
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


def mode_name():
    from bmnas import native as N
    return {0: 'fp32 FFMA', 1: '3xTF32 tcgen05 (fp32-class)', 2: '1xTF32 tcgen05', 3: 'bf16 operands in the fused MixedOp forward, 3xTF32 elsewhere'}[
        int(N.lib().bmnas_get_gemm_mode())]


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
    runner = training_runner(head)
    keep_gout = torch.zeros_like(runner.out)
    runner.prog.bind('gout', keep_gout)        # the upstream-gradient slot pointed at a freed autograd temporary
    mo = mixedop_roofline(runner.prog, c, B, pk, R=10, reps=3)
    best = {}
    for call in runner.prog.fwd + runner.prog.bwd:
        by = algorithmic_bytes(call, c, B)
        if by and (call.name not in best or by > best[call.name][1]):
            best[call.name] = (call, by)
    tr = ncu_traffic()
    for ph, d in mo.items():
        t = sum(tr.get(f'{k}@B{B}', 0) for k in d['launches'])
        d['traffic'] = t or None
        d['traffic_over_algorithmic'] = round(t / d['algorithmic_bytes'], 2) if t else None
    out = {'B': B, 'peak_GBs': pk['hbm'], 'peak_TFLOPs_bf16': pk['bf16'], 'peak_source': pk['src'], 'gemm_mode': mode_name(),
           'mixedop': mo, 'kernels': {}}
    for name, (call, by) in sorted(best.items()):
        us = graph_time_us(call, R=10, reps=3)
        gbs = by / (us * 1e-6) / 1e9
        out['kernels'][name] = {'us': round(us, 1), 'algorithmic_MB': round(by / 1e6, 1), 'GBs': round(gbs, 1),
                                'frac': round(gbs / pk['hbm'], 4)}
    del ss, head, pool
    torch.cuda.empty_cache()
    return out


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the ncu --set full captures committed under profiles/
    (tools/make_profiles.py rewrites the file from the .ncu-rep files of a build and stamps that build's git SHA)"""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
    except Exception:
        return {}


def training_runner(head):
    """the weight-step plan (runtime.GRAD_MODE 'weights', else 'all'): it carries every GEMM of a training step -- the
    arch-step plan omits the weight-gradient GEMMs"""
    rs = {k: r for k, r in head.fusion_net._bm_cache.items() if r.prog.training and r.prog.want_backward}
    for mode in ('weights', 'all', 'arch'):
        for k, r in rs.items():
            if mode in k:
                return r
    raise RuntimeError('no training plan built yet')


def mixedop_roofline(prog, c, B, pk, R=50, reps=4):
    """Roofline of the NodeMixedOp as SURVEY 8(d) defines it -- ALL launches that implement one mixed op, summed:
         forward  = bmnas_mixed_fwd, or bmnas_conv_fwd + bmnas_node_fwd on the two-kernel path
         backward = bmnas_node_bwd + bmnas_conv_dgrad + bmnas_conv_wgrad
       algorithmic bytes (fp32, inputs aliased as in the searchable cell): fwd 2*T1*B + 6C^2*4 (t in, out out, weights once),
       bwd 3*T1*B + 2*6C^2*4 (g and t in, gt out; weights in, weight gradients out);  T1 = C*L*4
       algorithmic FLOPs: fwd (12*L*C^2 + 4*L^2*C)*B, bwd twice that.
       Time = sum of the launches' device times, each measured live (CUDA events around graph-replayed launches)."""
    C, L = c['C'], c['L']
    T1 = C * L * 4
    groups = {}
    for phase, calls in (('fwd', prog.fwd), ('bwd', prog.bwd)):
        for call in calls:
            if getattr(call, 'tag', None):
                groups.setdefault((call.tag, phase), []).append(call)
    out = {}
    for phase in ('fwd', 'bwd'):
        cand = [(k, v) for k, v in groups.items() if k[1] == phase]
        if not cand:
            continue
        # the first mixed op of the plan that carries no chained edge mix (its bytes are the plain SURVEY 8(d) ones)
        plain = [(k, v) for k, v in cand if not any(getattr(x.st, 'out2', None) or getattr(x.st, 'gout2', None) for x in v
                                                    if x.name.startswith('bmnas_node'))]
        key, calls = (plain or cand)[0]
        times = {x.name: graph_time_us(x, R=R, reps=reps) for x in calls}
        us = sum(times.values())
        mult = 1 if phase == 'fwd' else 2
        by = (2 if phase == 'fwd' else 3) * T1 * B + mult * 6 * C * C * 4
        fl = mult * (12 * L * C * C + 4 * L * L * C) * B
        gbs, tfs = by / (us * 1e-6) / 1e9, fl / (us * 1e-6) / 1e12
        t_roof = max(by / (pk['hbm'] * 1e9), fl / (pk['bf16'] * 1e12)) * 1e6
        out[phase] = {'launches': {k: round(v, 2) for k, v in times.items()}, 'us': round(us, 2),
                      'algorithmic_bytes': by, 'algorithmic_flops': fl, 'GBs': round(gbs, 1), 'TFLOPs': round(tfs, 2),
                      'hbm_frac': round(gbs / pk['hbm'], 4), 'tensor_frac': round(tfs / pk['bf16'], 4),
                      'roofline_us': round(t_roof, 2), 'roofline_frac': round(t_roof / us, 4)}
    return out


def step_shares(prog, R=20):
    """device time of every launch of a training plan grouped by kernel: {name: (us per fwd+bwd, launches)}"""
    tot = {}
    for call in prog.fwd + prog.bwd:
        us = graph_time_us(call, R=R, reps=2)
        t, n = tot.get(call.name, (0.0, 0))
        tot[call.name] = (t + us, n + 1)
    return tot


def kernel_roofline(head, c, pk, B):
    """the `roofline` object of the bench line, for the batch the line is quoted on: the MixedOp kernel group that takes
    the largest share of the step (measured), against the HBM roofline, SURVEY 8(d) bytes."""
    runner = training_runner(head)
    prog = runner.prog
    keep_gout = torch.zeros_like(runner.out)
    prog.bind('gout', keep_gout)               # the upstream-gradient slot pointed at a freed autograd temporary
    mo = mixedop_roofline(prog, c, B, pk)
    shares = step_shares(prog)
    total = sum(t for t, _ in shares.values())
    n_mixed = len({x.tag for x in prog.fwd if getattr(x, 'tag', None)})
    dom = max(('fwd', 'bwd'), key=lambda ph: mo.get(ph, {}).get('us', 0.0))
    d = mo[dom]
    tr = ncu_traffic()
    traffic = sum(tr.get(f'{k}@B{B}', 0) for k in d['launches']) or None
    return {'bound': 'hbm', 'kernel': f'NodeMixedOp {dom} = ' + ' + '.join(d['launches']), 'achieved': d['GBs'], 'peak': pk['hbm'],
            'peak_source': pk['src'] + ' (burst: kernels timed alone)', 'unit': 'GB/s', 'frac': d['hbm_frac'],
            'traffic': traffic, 'traffic_git': tr.get('_git'), 'traffic_over_algorithmic': round(traffic / d['algorithmic_bytes'], 2) if traffic else None,
            'tensor_frac_vs_bf16_peak': d['tensor_frac'], 'roofline_frac': d['roofline_frac'],
            'share_of_fwd_bwd': round(n_mixed * d['us'] / total, 3), 'mixedop': mo,
            'time_share_by_kernel': {k: {'us': round(t, 1), 'launches': n, 'share': round(t / total, 3)}
                                     for k, (t, n) in sorted(shares.items(), key=lambda kv: -kv[1][0])},
            'note': f'B={B}: the whole working set is L2-resident and every launch is latency bound; roofline_large_batch '
                    'times the same plan where the kernels are throughput bound'}


def profile_kernels(head, R=50):
    """device time of every prepared call of the training plan: R back-to-back launches of the SAME call captured
    into one CUDA graph and replayed (no host launch overhead in the number; stream launches through ctypes are
    host-bound at ~3 us and hide anything shorter)"""
    from bmnas import native as N
    import ctypes
    runner = training_runner(head)
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


METRIC = 'search-step samples/sec (fwd+bwd+arch step, fwd+bwd+weight step)'


def workload_string(cfgname, c, gB):
    """identical in both arms (the driver compares the strings)"""
    return (f'{cfgname.upper()} fusion search step on synthetic frozen-backbone features, global batch {gB}, C={c["C"]}, '
            f'L={c["L"]}, n_in={c["num_input_nodes"]}, steps={c["steps"]}, node_steps={c["node_steps"]}, classes={c["classes"]}')


def build_search(c, device, group=None, use_graphs=True, nbpe=400.0, peer_step=None):
    import types
    from bmnas.nn import SearchHead, CrossEntropyLoss, BCEWithLogitsLoss
    from bmnas.search import SearchStep
    a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps',
                                                    'node_multiplier', 'drpt')}, weight_decay=c['weight_decay'])
    crit = CrossEntropyLoss() if c['loss'] == 'ce' else BCEWithLogitsLoss()
    head = SearchHead(a, c['classes'], criterion=crit).to(device)
    ss = SearchStep(head, crit, c['B'], c['classes'], loss_kind=c['loss'], eta_max=c['eta_max'],
                    weight_decay=c['weight_decay'], nbpe=nbpe, use_graphs=use_graphs, group=group, peer_step=peer_step)
    return head, ss


def quick_value(cfgname, device, steps=60, warmup=5, batch=None):
    """value / ms_per_step of another BASELINE config (single GPU, graphs, inputs resident and rotated through a pool
    larger than L2), for the `configs` block of the bench line"""
    c = dict(CONFIGS[cfgname])
    if batch:
        c['B'] = batch
    torch.manual_seed(2)
    head, ss = build_search(c, device)
    per_batch = c['num_input_nodes'] * c['B'] * c['C'] * c['L'] * 4
    n_pool = max(4, int(1.3 * L2_BYTES / per_batch) + 1)
    n_pool += n_pool % 2
    pool = make_pool(c, n_pool, 300, device)
    ss.load('dev', *pool[0]); ss.load('train', *pool[1])
    ss.prepare(warmup=3, restore=False)
    for i in range(warmup):
        ss.load('dev', *pool[(2 * i) % n_pool]); ss.load('train', *pool[(2 * i + 1) % n_pool]); ss.step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        ss.load('dev', *pool[(2 * i) % n_pool]); ss.load('train', *pool[(2 * i + 1) % n_pool]); ss.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {'workload': workload_string(cfgname, c, c['B']), 'value': round(c['B'] / ms * 1e3, 1), 'unit': 'samples/s',
           'ms_per_step': round(ms, 4), 'steps': steps, 'launches_per_step': ss.launches_per_step,
           'weights': sum(p.numel() for p in head.parameters())}
    del ss, head, pool
    torch.cuda.empty_cache()
    return out


def found_sweep(device, batches=(96, 1024, 8192), steps=30):
    """BASELINE configs[4]: found (fixed-genotype) NTU network, train and inference throughput (bench_found.py)"""
    import bench_found as BF
    return BF.sweep(device, batches, steps)


def dp_check(ss, head, group, world, rank, device):
    """N > 1 only, before timing, on the REAL optimiser path (peer-memory reduce-scatter + Adam + all-gather kernel, or NCCL
    all-reduce + FusedAdam): for each half step, the parameters every rank holds after the data-parallel step must equal
    one torch Adam step (zero moments, step 1) on the gradient summed over the ranks and divided by the world size --
    the per-rank gradients are all-gathered and summed here independently of the path under test.  State is restored
    afterwards.  Raises on mismatch."""
    import torch.distributed as dist
    from bmnas import runtime as rt
    from bmnas.program import join_side
    snap = ss._snapshot()
    res = {'path': 'peer-memory fused reduce-scatter + Adam + all-gather (bmnas_dp_adam_step)' if ss.peer is not None
           else 'ncclAllReduce + bmnas_adam_step'}
    hyp = {'dev': (ss.a_opt.defaults['lr'], ss.a_opt.defaults['betas'], ss.a_opt.defaults['weight_decay']),
           'train': (ss.w_opt.param_groups[0]['lr'], ss.w_opt.defaults['betas'], ss.w_opt.defaults['weight_decay'])}
    for which, mode in (('dev', 'arch'), ('train', 'weights')):
        tensors = head.arch_parameters() if which == 'dev' else [p for p in head.parameters() if p.requires_grad]
        before = [t.detach().clone() for t in tensors]
        with rt.grad_mode(mode), rt.static_io():
            loss = ss.criterion(head(ss.feats[which]), ss.labels[which])
            loss.backward()
        join_side(device)
        torch.cuda.synchronize()
        local = torch.cat([t.grad.detach().reshape(-1) for t in tensors])
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local, group=group)
        g = (torch.stack(gathered).double().sum(0) / world).float()
        if ss.peer is not None:
            ss.peer.step(which)
        else:
            ss.allreduce_grads(which)
            (ss.a_opt if which == 'dev' else ss.w_opt).step()
        torch.cuda.synchronize()
        lr, (b1, b2), wd = hyp[which]
        p0 = torch.cat([t.reshape(-1) for t in before])
        g = g + wd * p0
        m, v = (1 - b1) * g, (1 - b2) * g * g
        expect = p0 - (lr / (1 - b1)) * (m / (v.sqrt() / (1 - b2) ** 0.5 + 1e-8))
        got = torch.cat([t.detach().reshape(-1) for t in tensors])
        err = (got - expect).abs().max().item()
        moved = (expect - p0).abs().max().item()
        if not err <= 1e-3 * moved + 1e-9:
            raise SystemExit(f'dp_check: {which} parameters after the data-parallel step differ from Adam(sum of per-rank gradients / world): '
                             f'err {err:.3e}, step size {moved:.3e}')
        res[which] = {'max_err': err, 'max_update': moved}
    ss._restore(snap)
    torch.cuda.synchronize()
    dist.barrier(group)
    return res


def replica_checksum(head, group, world, device):
    import torch.distributed as dist
    flat = torch.cat([p.detach().reshape(-1) for p in head.parameters()] + [a.detach().reshape(-1) for a in head.arch_parameters()])
    bits = flat.view(torch.int32).to(torch.int64)
    sig = torch.stack([bits.sum(), (bits * torch.arange(1, bits.numel() + 1, device=device) % 1000003).sum()])
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig, group=group)
    if not all(torch.equal(sigs[0], x) for x in sigs):
        raise SystemExit('dp_check: replicas diverged (parameter checksums differ across ranks)')
    return 'ok'


def run_ours(args):
    import torch.distributed as dist
    from bmnas import native as N
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
    if args.gemm_mode is not None:
        N.lib().bmnas_set_gemm_mode(args.gemm_mode)
    c = dict(CONFIGS[args.config])
    if args.batch:
        c['B'] = args.batch
    gB = c['B'] * world if args.scaling == 'weak' else c['B']
    if args.scaling == 'strong':
        # SURVEY 8(d) config 3 / nn.DataParallel semantics (ntu_darts_searchable.py:50-51): the GLOBAL batch stays at
        # the script's value and is chunked over the GPUs
        if c['B'] % world:
            raise SystemExit(f'--scaling strong: global batch {c["B"]} is not divisible by {world} GPUs')
        c['B'] = c['B'] // world
    torch.manual_seed(2)                              # main_darts_searchable_ntu.py:17 (all ranks: identical replicas)
    head, ss = build_search(c, device, group=group, use_graphs=not args.no_graphs, peer_step=(False if args.nccl else None))
    # input pool larger than L2 so every step's inputs come from HBM
    per_batch = c['num_input_nodes'] * c['B'] * c['C'] * c['L'] * 4
    n_pool = max(4, int(1.3 * L2_BYTES / per_batch) + 1)
    n_pool += n_pool % 2
    pool = make_pool(c, n_pool, 100 + rank, device)
    ss.load('dev', *pool[0]); ss.load('train', *pool[1])
    # one resident byte tensor per step (SearchStep.pack_step: [dev feats | train feats | dev labels | train labels]):
    # a step's inputs reach the static buffers with ONE device-to-device copy
    ppool = [ss.pack_step(pool[2 * j][0], pool[2 * j][1], pool[2 * j + 1][0], pool[2 * j + 1][1], device=device)
             for j in range(n_pool // 2)]
    del pool[2:]
    dp = None
    if world > 1:
        dp = dp_check(ss, head, group, world, rank, device)
    ss.prepare(warmup=3, restore=False)
    n_weights = sum(p.numel() for p in head.parameters())
    n_arch = sum(p.numel() for p in head.arch_parameters())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(K, loader, read_loss, metrics_fwd=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for i in range(K):
            loader(i)
            if metrics_fwd:                       # the reference's dev phase: Architect.step, then a no-grad metrics forward
                ss.half('dev')
                ss.metrics_forward('dev')
                lw = ss.half('train')
                ss.steps_done += 1
            else:
                la, lw = ss.step()
            if read_loss:
                lw_host = lw.item()               # device->host read of the step's result
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        return e0.elapsed_time(e1), wall          # device time span of the K steps (idle gaps included)

    def load_dev(i):
        ss.load_step(ppool[i % len(ppool)])
    for i in range(args.warmup):
        load_dev(i); ss.step()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    # EXACTLY --steps steps per timed region; the region is repeated so that at least ~200 steps are inside timed regions
    # in total and the spread is visible (a 20-step region is 13 ms)
    reps = max(1, min(10, -(-200 // max(args.steps, 1))))
    regions = [timed(args.steps, load_dev, read_loss=False)[0] for _ in range(reps)]
    dev_ms = sorted(regions)[len(regions) // 2]
    # full fidelity: + the dev-phase no-grad metrics forward of the reference loop (train_searchable/ntu.py:81-85)
    ss.capture_metrics_forward('dev')
    timed(max(3, args.warmup), load_dev, read_loss=False, metrics_fwd=True)
    ff_ms, _ = timed(args.steps, load_dev, read_loss=False, metrics_fwd=True)
    # ---- end to end: pinned host inputs, H2D every step, D2H loss read every step
    hpool = make_pool(c, 8, 500 + rank, device, pinned=True)

    hppool = [ss.pack_step(hpool[2 * j][0], hpool[2 * j][1], hpool[2 * j + 1][0], hpool[2 * j + 1][1], pinned=True) for j in range(4)]

    def load_host(i):
        ss.load_step(hppool[i % 4])
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
    half_wall = sorted(timed_pipelined(args.steps) for _ in range(min(reps, 5)))[min(reps, 5) // 2]
    ss.reset_pipeline()

    def timed_step_pipeline(K):
        """the same loop with the step-level pipeline: SearchStep.prefetch_step() sends the packed inputs of step i+1 to a
        staging buffer on the copy stream while step i computes, step() moves them into the static buffers (one device copy)
        and replays the one-graph step, read_loss_async() copies the step's two losses to pinned host memory behind the
        step; the host consumes the loss of step i while step i+1 runs (every step's loss is read inside the region)."""
        barrier()
        t0 = time.perf_counter()
        ss.prefetch_step(hppool[0])
        pending = None
        for i in range(K):
            ss.step()
            h = ss.read_loss_async()
            if i + 1 < K:
                ss.prefetch_step(hppool[(i + 1) % 4])
            if pending is not None:
                la_host, lw_host = pending.get()
            pending = h
        la_host, lw_host = pending.get()
        barrier()
        return time.perf_counter() - t0
    timed_step_pipeline(max(3, args.warmup))
    e2e_wall = sorted(timed_step_pipeline(args.steps) for _ in range(min(reps, 5)))[min(reps, 5) // 2]
    ss.reset_pipeline()
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_wall * 1e3, serial_wall * 1e3, ff_ms, half_wall * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms_wall, serial_ms_wall, ff_ms, half_ms_wall = t.tolist()
    if world > 1:
        dp['replicas_after_timed_steps'] = replica_checksum(head, group, world, device)
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
        roof = kernel_roofline(head, c, pk, c['B'])
        big = big16 = None
        extra = None
        if world == 1 and args.roofline_batch > 0:
            big = large_batch_roofline(args, c, pk, device, args.roofline_batch)
            if args.gemm_mode is None:             # the same plan with bf16 operands in the fused MixedOp forward
                N.lib().bmnas_set_gemm_mode(3)
                try:
                    big16 = large_batch_roofline(args, c, pk, device, args.roofline_batch)
                finally:
                    N.lib().bmnas_set_gemm_mode(1)
        if world == 1 and not args.no_configs and args.config == 'ntu' and not args.batch:
            extra = {}
            for name in ('mmimdb', 'ego', 'ego_large'):
                try:
                    extra[name] = quick_value(name, device)
                except Exception as e:
                    extra[name] = {'error': repr(e)[:200]}
            try:
                extra['found_ntu'] = found_sweep(device)
            except Exception as e:
                extra['found_ntu'] = {'error': repr(e)[:200]}
        cpu = None
        if world == 1 and not args.no_cpu:
            r = reference_subprocess(args.config, 40, 3, 20.0, batch=c['B'])
            if 'unavailable' in r:
                cpu = r
            else:
                cpu = dict(r['cpu_baseline'], ms_per_step=r['ms_per_step'])
                g = reference_subprocess(args.config, 60, 3, 10.0, device='cuda', batch=c['B'])
                cpu['torch_eager_gpu'] = ({'value': g['value'], 'unit': 'samples/s', 'ms_per_step': g['ms_per_step'],
                                           'sample': g['cpu_baseline']['sample']} if 'value' in g else g)
        out = {
            'metric': METRIC, 'value': round(value, 1),
            'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': round(ms_per_step, 4), 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32' if N.lib().bmnas_get_gemm_mode() != 3 else 'f32 (bf16 operands in the fused MixedOp forward)',
            'data': 'synthetic',
            'config': {'workload': workload_string(args.config, c, gB),
                       'global_batch': gB, 'per_gpu_batch': c['B'], 'weights': n_weights, 'arch_scalars': n_arch,
                       'parallelism': f'dp{world} (batch-sharded; per half step ONE ' +
                                      ('peer-memory kernel: reduce-scatter + Adam on the shard + all-gather of parameters over NVLink'
                                       if ss.peer is not None else 'NCCL all-reduce of the half\'s gradient span, then fused Adam') +
                                      f'; arch bucket {n_arch} floats, weight bucket {n_weights} floats)',
                       'cuda_graphs': not args.no_graphs, 'gemm_mode': mode_name(),
                       'l2_policy': f'inputs larger than L2: pool of {n_pool} distinct resident batches '
                                    f'({n_pool * per_batch / 2**20:.0f} MiB) rotated every step'},
            'timed_regions': {'repeats': reps, 'steps_each': args.steps, 'ms_per_step_min': round(min(regions) / args.steps, 4),
                              'ms_per_step_max': round(max(regions) / args.steps, 4), 'reported': 'median region'},
            'value_full_fidelity': round(gB / (ff_ms / args.steps * 1e-3), 1),
            'full_fidelity_how': 'arch half + no-grad train-mode metrics forward on the dev batch (train_searchable/ntu.py:81-85) + '
                                 'weight half; device time',
            'e2e': {'value': round(e2e_value, 1), 'unit': 'samples/s',
                    'h2d_bytes_per_step': 2 * (per_batch + lab_bytes), 'd2h_bytes_per_step': 8,
                    'ms_per_step': round(e2e_ms_wall / args.steps, 4),
                    'how': 'SearchStep.prefetch_step() from pinned host memory (copy stream -> staging buffer; the packed inputs of '
                           'step i+1 travel while step i computes) + SearchStep.step() (device copy staging -> static buffers, one '
                           'graph) + read_loss_async() every step (both losses to pinned host memory behind the step; the host '
                           'reads the value of step i while step i+1 runs), wall clock between barriers',
                    'sync_read_value': round(gB * args.steps / (half_ms_wall * 1e-3), 1),
                    'sync_read_how': 'round-1 loop: per-half SearchStep.prefetch() + step() + a blocking loss.item() every step '
                                     '(the GPU idles while the host reads and relaunches)',
                    'serial_value': round(gB * args.steps / (serial_ms_wall * 1e-3), 1),
                    'serial_how': 'SearchStep.load() + step() + loss.item(): copy, compute and read strictly in sequence'},
            'gpu_launches': (ss.launches_per_step or 0) * args.steps,
            'launches_per_step': ss.launches_per_step,
            'roofline': roof, 'roofline_large_batch': big, 'roofline_large_batch_bf16': big16, 'configs': extra,
            'cpu_baseline': cpu, 'dp_check': dp, 'clocks': clk,
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
    """`--impl reference`: the reference's own CPU implementation of the path (oracle/_ref driven by
    oracle/ref_harness.py; the oracle port if that is absent) on all host cores, same config / metric / unit;
    at N > 1 rank 0 alone runs it, on the GLOBAL batch of our arm."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    c = dict(CONFIGS[args.config])
    if args.batch:
        c['B'] = args.batch
    gB = c['B'] * max(args.gpus, 1) if args.scaling == 'weak' else c['B']
    dev = args.ref_device
    kind, v, ms, cores, done = reference_arm(args.config, max(args.steps, 1), max(args.warmup, 1), args.max_seconds,
                                             device=dev, full_fidelity=args.full_fidelity, batch=gB)
    what = ('reference modules (oracle/_ref: FusionNetwork + Architect + torch.optim.Adam + LRCosineAnnealingScheduler, unmodified)'
            if kind == 'reference' else 'oracle port (oracle/_ref absent)')
    where = f'torch CPU fp32, {cores} threads' if dev == 'cpu' else 'eager PyTorch fp32 (TF32 off) on the B200, alpha/beta/gamma left on the CPU as shipped'
    sample = f'{done} search steps of the same workload (global batch {gB}) through the {what}; {where}'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC,
        'value': round(v, 1), 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': done, 'warmup': args.warmup,
        'ms_per_step': round(ms, 3), 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_string(args.config, c, gB), 'global_batch': gB},
        'cpu_baseline': {'value': round(v, 1), 'unit': 'samples/s', 'cores': cores, 'kind': kind, 'sample': sample},
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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: the config batch per GPU; strong: the config batch is the GLOBAL batch (nn.DataParallel chunking)')
    ap.add_argument('--gemm-mode', type=int, default=None, help='bmnas_set_gemm_mode: 1 = 3xTF32 (default), 3 = bf16 fused MixedOp forward')
    ap.add_argument('--no-graphs', action='store_true')
    ap.add_argument('--nccl', action='store_true', help='N > 1: NCCL all-reduce + FusedAdam instead of the peer-memory fused optimiser step')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the other BASELINE configs (mmimdb, ego, ego_large, found sweep)')
    ap.add_argument('--profile-kernels', action='store_true')
    ap.add_argument('--roofline-batch', type=int, default=8192,
                    help='also time the kernels at this per-GPU batch (bandwidth-bound regime); 0 = skip')
    ap.add_argument('--max-seconds', type=float, default=120.0, help='--impl reference: wall-clock bound of the timed loop')
    ap.add_argument('--ref-device', default='cpu', help='--impl reference: cpu (the arm) or cuda (eager-GPU baseline)')
    ap.add_argument('--full-fidelity', action='store_true', help='--impl reference: include the dev-phase metrics forward')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == '__main__':
    main()
