"""GPU probe: per-engine error and warm launch time of the conv GEMM family at a given shape.
   python tools/gemm_probe.py [B]"""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from bmnas import native as N
import test_gpu_gemm as T
lib = N.lib()
dev = torch.device('cuda:0')
Bs = [int(a) for a in sys.argv[1:]] or [96]
for B in Bs:
  SHAPES = [(8, [128], [256, 128], 2), (8, [128, 128], [128], 1)]
  if os.environ.get('PROBE_EGO') == '1':      # Ego-large: C = 256, L = 16, node_multiplier 3
      SHAPES = [(16, [256], [512, 256], 2), (16, [256, 256, 256], [256], 1)]
  for (L, src_C, seg_M, w_fold) in SHAPES:
    srcs, Ws, bias = T._conv_case(B, L, src_C, seg_M, w_fold, 1, dev)
    M, K = sum(seg_M), sum(src_C)
    Zr, mr, rr, Weff, U = T._ref_fwd(srcs, Ws, bias, w_fold)
    GV = torch.randn(B, M, L, device=dev)
    dU = torch.einsum('mk,bml->bkl', Weff, GV.double()); dW = torch.einsum('bml,bkl->mk', GV.double(), U)
    FMT = int(os.environ.get('PROBE_FMT', '0'))      # 0: tcgen05 slabs (ws / panel kernels), 1: plain fp32 (gemm_sg.cu FFMA kernels)
    imgs = T._images(N, lib, Ws, seg_M, K, w_fold, dev, fmt=FMT) if os.environ.get('PROBE_IMG', '1') == '1' else None
    for mode in [int(m) for m in os.environ.get('PROBE_MODES', '0,1,2').split(',')]:
        lib.bmnas_set_gemm_mode(mode)
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
        if imgs is not None: st.wimg_fwd = imgs[0].data_ptr(); st.wimg_fmt = FMT
        sd = T._params(N, B, L, src_C, seg_M, w_fold, srcs, Ws); sd.GV = GV.data_ptr()
        if imgs is not None: sd.wimg_dgrad = imgs[1].data_ptr(); sd.wimg_fmt = FMT
        gs = [torch.zeros(B, c, L, device=dev) for c in src_C]
        for i in range(len(src_C)): sd.gsrc[i] = gs[i].data_ptr()
        sw = T._params(N, B, L, src_C, seg_M, w_fold, srcs, Ws); sw.GV = GV.data_ptr()
        gW = [torch.zeros(m, w_fold * K, device=dev) for m in seg_M]; gb = [torch.zeros(m, device=dev) for m in seg_M]
        for i in range(len(seg_M)): sw.gW[i], sw.gbias[i] = gW[i].data_ptr(), gb[i].data_ptr()
        s = N.current_stream()
        res = {}
        for name, stt in (('bmnas_conv_fwd', st), ('bmnas_conv_dgrad', sd), ('bmnas_conv_wgrad', sw)):
            for _ in range(5): N.launch(name, ctypes.byref(stt), s)
            for g_ in gW: g_.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            R = 100
            e0.record()
            for _ in range(R): N.launch(name, ctypes.byref(stt), s)
            e1.record(); torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) * 1e3 / R
        errs = (T._rel(Z, Zr), T._rel(torch.cat(gs, 1), dU), T._rel(torch.cat(gW, 0)[:, :K] / 100.0, dW))
        print(f'B={B} M={M} K={K} fold={w_fold} mode={mode}: fwd {res["bmnas_conv_fwd"]:.2f} us dgrad {res["bmnas_conv_dgrad"]:.2f} us '
              f'wgrad {res["bmnas_conv_wgrad"]:.2f} us | rel err Z {errs[0]:.2e} dU {errs[1]:.2e} dW {errs[2]:.2e}', flush=True)
lib.bmnas_set_gemm_mode(1)
