"""fused (bmnas_mixed_fwd) vs two-kernel (bmnas_conv_fwd + bmnas_node_fwd) NodeMixedOp forward over the batch size:
device time per forward, graph-replayed.  python tools/fused_crossover.py [mode]"""
import sys, os, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch
import test_gpu_mixed as T
import gpu_util as U
import bench as BN
from bmnas import program, native as N
if len(sys.argv) > 1:
    N.lib().bmnas_set_gemm_mode(int(sys.argv[1]))
L = 8
print('# NodeMixedOp forward, C=128, L=8, train-mode BatchNorm, Philox dropout; us per forward (graph replay)')
print('# %8s %12s %12s %12s %12s' % ('B', 'fused+Z', 'fused noZ', 'two-kernel', 'conv+node'))
for B in (96, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    row = []
    for fused, grad in (('1', True), ('1', False), ('0', True)):
        program.FUSED_MIXED = fused
        program.FUSED_MIXED_MIN_B = 0
        mod = T._mixed(L).to(U.DEV).train()
        x = torch.randn(B, T.C, L, device=U.DEV).requires_grad_(grad)
        w = torch.softmax(torch.randn(4), -1).to(U.DEV)
        with torch.set_grad_enabled(grad):
            for _ in range(2):
                mod(x, x, w)
        torch.cuda.synchronize()
        runner = [r for r in mod._bm_cache.values() if r.prog.want_backward == grad][0]
        calls = [c for c in runner.prog.fwd if c.name in ('bmnas_mixed_fwd', 'bmnas_conv_fwd', 'bmnas_node_fwd')]
        ts = [BN.graph_time_us(c, R=10, reps=3) for c in calls]
        row.append((sum(ts), ts))
    print('  %8d %12.1f %12.1f %12.1f   %s' % (B, row[0][0], row[1][0], row[2][0], ' + '.join('%.1f' % t for t in row[2][1])))
