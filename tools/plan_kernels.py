"""plan driver: run every prepared call of the training plan standalone at a large batch, one at a time with a
synchronize, printing its name first (an illegal access is sticky: the first failure is the culprit).
    python tools/plan_kernels.py [B] [name-filter] [eager|graph]
    compute-sanitizer --print-limit 8 python tools/plan_kernels.py 8192 conv_dgrad"""
import ctypes
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
filt = sys.argv[2] if len(sys.argv) > 2 else ''
mode = sys.argv[3] if len(sys.argv) > 3 else 'eager'
from bmnas.nn import SearchHead, CrossEntropyLoss  # noqa: E402
from bmnas.search import SearchStep  # noqa: E402
device = torch.device('cuda:0')
c = dict(bench.CONFIGS['ntu'], B=B)
a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps',
                                                'node_multiplier', 'drpt')}, weight_decay=c['weight_decay'])
crit = CrossEntropyLoss()
head = SearchHead(a, c['classes'], criterion=crit).to(device)
ss = SearchStep(head, crit, B, c['classes'], loss_kind='ce', use_graphs=False)
pool = bench.make_pool(c, 2, 77, device)
ss.load('dev', *pool[0]); ss.load('train', *pool[1])
for _ in range(2):
    ss.step()
torch.cuda.synchronize()
print('steps ok', flush=True)
runner = [r for r in head.fusion_net._bm_cache.values() if r.prog.training][0]
keep_gout = torch.zeros_like(runner.out)
runner.prog.bind('gout', keep_gout)
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
for phase, calls in (('fwd', runner.prog.fwd), ('bwd', runner.prog.bwd)):
    for i, call in enumerate(calls):
        if filt and filt not in call.name:
            continue
        st = call.st
        dims = {k: getattr(st, k) for k in ('B', 'L', 'K', 'M', 'C', 'n', 'Ctot', 'n_src', 'w_fold', 'mode', 'n_ops', 'wimg_fmt', 'early_ok')
                if hasattr(st, k)}
        print(phase, i, call.name, dims, end=' ... ', flush=True)
        if mode == 'eager':
            for _ in range(3):
                flush.zero_()                  # evict L2 (126 MB) so that a profiled launch reads its operands from DRAM
                call(s)
            torch.cuda.synchronize()
            print('ok', flush=True)
        else:
            print('%.1f us' % bench.graph_time_us(call, R=10, reps=3), flush=True)
print('all ok')
