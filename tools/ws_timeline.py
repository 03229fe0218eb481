"""in-kernel timeline of the warp-specialised conv GEMM (csrc/gemm_ws.cu, bmnas_ws_timeline): %globaltimer stamps of the
middle CTA, us since kernel entry.   python tools/ws_timeline.py [B ...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from bmnas import native as N
import test_gpu_gemm as T
lib = N.lib()
dev = torch.device('cuda:0')
NAMES = ['entry', 'setup + pdl_wait done', 'producers: first stage full', 'producers: done', 'MMA: first tile issued',
         'MMA: all issued', 'epilogue: first tile done', 'epilogue: done', 'teardown barrier passed', 'last_block passed',
         'finalize start (last CTA)', 'finalize done (last CTA)', 'finalize: partials in smem', 'finalize: pass 1 done']
tl = (ctypes.c_ulonglong * 32)()
for B in [int(a) for a in sys.argv[1:]] or [8192]:
    for (L, src_C, seg_M, w_fold) in [(8, [128], [256, 128], 2), (8, [128, 128], [128], 1)]:
        srcs, Ws, bias = T._conv_case(B, L, src_C, seg_M, w_fold, 1, dev)
        M, K = sum(seg_M), sum(src_C)
        imgs = T._images(N, lib, Ws, seg_M, K, w_fold, dev, fmt=0)
        st = T._params(N, B, L, src_C, seg_M, w_fold, srcs, Ws)
        Z = torch.zeros(B, M, L, device=dev); mean = torch.zeros(M, device=dev); rstd = torch.zeros(M, device=dev)
        st.bn_mode = 1
        keep = []
        for i, m in enumerate(seg_M):
            st.bias[i] = bias[i].data_ptr()
            a, b, c = torch.zeros(m, device=dev), torch.ones(m, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
            keep += [a, b, c]
            st.running_mean[i], st.running_var[i], st.num_batches_tracked[i] = a.data_ptr(), b.data_ptr(), c.data_ptr()
        part = torch.zeros(int(lib.bmnas_conv_stat_part_size(ctypes.byref(st))), device=dev)
        cnt = torch.zeros(int(lib.bmnas_conv_num_counters(ctypes.byref(st))), dtype=torch.int32, device=dev)
        st.Z, st.mean, st.rstd, st.stat_part, st.counter = Z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), part.data_ptr(), cnt.data_ptr()
        st.wimg_fwd, st.wimg_fmt = imgs[0].data_ptr(), 0
        GV = torch.randn(B, M, L, device=dev)
        ca, cb, cc = (torch.randn(M, device=dev) for _ in range(3))
        sd = T._params(N, B, L, src_C, seg_M, w_fold, srcs, Ws)
        sd.GV, sd.Z = GV.data_ptr(), Z.data_ptr()
        sd.coef_a, sd.coef_b, sd.coef_c = ca.data_ptr(), cb.data_ptr(), cc.data_ptr()
        sd.wimg_dgrad, sd.wimg_fmt = imgs[1].data_ptr(), 0
        gs = [torch.zeros(B, c, L, device=dev) for c in src_C]
        for i in range(len(src_C)):
            sd.gsrc[i] = gs[i].data_ptr()
        s = N.current_stream()
        for name, stt in (('bmnas_conv_fwd', st), ('bmnas_conv_dgrad', sd)):
            for _ in range(3):
                N.launch(name, ctypes.byref(stt), s)
            torch.cuda.synchronize()
            lib.bmnas_ws_timeline(None, 1)
            N.launch(name, ctypes.byref(stt), s)
            torch.cuda.synchronize()
            lib.bmnas_ws_timeline(tl, 0)
            t0 = tl[0]
            print(f'== {name} B={B} M={M} K={K}')
            for i, nm in enumerate(NAMES):
                if tl[i] >= t0 and (name == 'bmnas_conv_fwd' or i < 9):
                    print(f'  {nm:<34s} {(tl[i] - t0) / 1e3:8.2f} us')
            mhz = 1965.0
            print(f'  stages {tl[21]}: producer thread 0 waited {tl[16] / mhz:.1f} us for its copies, {tl[17] / mhz:.1f} us for a free stage; '
                  f'MMA warp waited {tl[18] / mhz:.1f} us for weight slabs, {tl[19] / mhz:.1f} us for activation stages, '
                  f'{tl[20] / mhz:.1f} us for a free accumulator set')
