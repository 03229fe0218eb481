"""phase timeline (globaltimer, ns) of the tcgen05 GEMM kernels at the NTU B=96 shapes; needs the -DBMNAS_TIMELINE build"""
import os, sys, types, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['BMNAS_LIB'] = os.path.join(ROOT, 'tools', 'ubench', 'libbmnas_tl.so')
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200')); sys.path.insert(0, ROOT)
import torch
import bench
from bmnas import native as N
from bmnas.nn import SearchHead, CrossEntropyLoss
from bmnas.search import SearchStep
c = dict(bench.CONFIGS['ntu'])
if len(sys.argv) > 1: c['B'] = int(sys.argv[1])
dev = torch.device('cuda:0')
a = types.SimpleNamespace(**{k: c[k] for k in ('C', 'L', 'num_input_nodes', 'steps', 'multiplier', 'node_steps', 'node_multiplier', 'drpt')}, weight_decay=3e-4)
crit = CrossEntropyLoss()
head = SearchHead(a, c['classes'], criterion=crit).to(dev)
ss = SearchStep(head, crit, c['B'], c['classes'], use_graphs=False)
pool = bench.make_pool(c, 2, 1, dev)
ss.load('dev', *pool[0]); ss.load('train', *pool[1])
for _ in range(600): ss.step()     # ~0.5 s of load: clocks at boost before the timelines are taken
torch.cuda.synchronize()
runner = [r for r in head.fusion_net._bm_cache.values() if r.prog.training][0]
prog = runner.prog
s = N.current_stream()
buf = (ctypes.c_ulonglong * 64)()
names = {0: ['entry', 'prologue done', 'first loads issued', 'chunk0 stored', 'chunk0 synced', 'chunk0 weights landed', 'mainloop done',
             'accumulator done', 'epilogue done', 'teardown sync', 'pre last_block', 'post last_block', 'finalize done', 'slab1 landed', 'slab2 landed', 'slab3 landed', 'slab4 landed', 'slab5 landed', 'slab6 landed']}
for mode, cname in ((0, 'bmnas_conv_fwd'), (1, 'bmnas_conv_dgrad'), (2, 'bmnas_conv_wgrad')):
    calls = [x for x in (prog.fwd + prog.bwd) if x.name == cname and x.st.M == 3 * c['C']]
    call = calls[0]
    for rep in range(3):
        for _ in range(200): call(s)
        torch.cuda.synchronize()
        N.lib().bmnas_debug_timeline(buf)
    t = [buf[mode * 20 + i] for i in range(19)]
    print(cname, 'K', call.st.K, 'M', call.st.M)
    for i in range(19):
        if t[i] >= t[0] and t[i] - t[0] < 10**8:
            print('   %-26s +%6d ns' % (names[0][i], t[i] - t[0]))

names_sg = ['entry', 'copies issued', 'copies landed', 'transform+sync', 'FFMA done', 'epilogue done / pre last_block', 'post last_block', 'finalize done', 'after pdl', 'after bias prefetch', 'A copies issued']
for mode, cname in ((0, 'bmnas_conv_fwd'), (1, 'bmnas_conv_dgrad')):
    calls = [x for x in (prog.fwd + prog.bwd) if x.name == cname and x.st.M == 3 * c['C']]
    call = calls[0]
    for rep in range(3):
        for _ in range(200): call(s)
        torch.cuda.synchronize()
        N.lib().bmnas_debug_timeline_sg(buf)
    t = [buf[mode * 20 + i] for i in range(11)]
    print('sg', cname, 'K', call.st.K, 'M', call.st.M, 'fmt', call.st.wimg_fmt)
    for i in range(11):
        if t[i] >= t[0] and t[i] - t[0] < 10**8:
            print('   %-32s +%6d ns' % (names_sg[i], t[i] - t[0]))
