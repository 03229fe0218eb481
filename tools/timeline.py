"""Kernel timeline of graph-replayed search steps (CUPTI through torch.profiler): start, duration, stream and the idle gap
before every kernel of ONE step, plus per-kernel totals.  python tools/timeline.py [config] [--batch B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    cfgname = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('--') else 'ntu'
    c = dict(bench.CONFIGS[cfgname])
    if '--batch' in sys.argv:
        c['B'] = int(sys.argv[sys.argv.index('--batch') + 1])
    # under torchrun (WORLD_SIZE > 1): the data-parallel step, rank 0 prints its timeline (TL_NCCL=1: NCCL path)
    world, rank, local = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0))
    dev = torch.device('cuda', local)
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        group = dist.group.WORLD
    torch.manual_seed(2)
    head, ss = bench.build_search(c, dev, group=group, use_graphs=True, peer_step=(False if os.environ.get('TL_NCCL') == '1' else None))
    pool = bench.make_pool(c, 4, 100 + rank, dev)
    ss.load('dev', *pool[0]); ss.load('train', *pool[1])
    ss.prepare(warmup=3, restore=False)
    pp = [ss.pack_step(pool[2 * j][0], pool[2 * j][1], pool[2 * j + 1][0], pool[2 * j + 1][1], device=dev) for j in range(2)]
    for i in range(10):
        ss.load_step(pp[i % 2]); ss.step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    NS = 6
    mode = os.environ.get('TL_MODE', 'step')
    if mode == 'upload':          # upload the graph for the next launch on another stream while this one runs
        import ctypes
        rt_ = ctypes.CDLL('libcudart.so.12')
        up = torch.cuda.Stream()
        ex = ctypes.c_void_p(ss.graphs['step'].raw_cuda_graph_exec())
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(NS):
            if mode != 'nocopy':          # nocopy: graph launches back to back (is the ramp at the head of a step the stream -> graph hand-over?)
                ss.load_step(pp[i % 2])
            if mode == 'upload':
                rc = rt_.cudaGraphUpload(ex, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert rc == 0, rc
            ss.step()
        torch.cuda.synchronize()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        if rank != 0:
            os._exit(0)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # one step = from one 'Memcpy DtoD' group to the next: split at the first copy of each step
    t0 = evs[0].time_range.start
    rows = [(e.time_range.start - t0, e.time_range.end - e.time_range.start, e.name) for e in evs]
    # find step boundaries: the first memcpy after a kernel
    starts = [i for i, r in enumerate(rows) if ('emcpy' in r[2] or 'k_copy' in r[2]) and (i == 0 or not ('emcpy' in rows[i - 1][2] or 'k_copy' in rows[i - 1][2]))]
    # a step has two loads (dev + train) back to back -> boundaries every group; keep groups that start a step
    if mode == 'nocopy':
        starts = [i for i, r in enumerate(rows) if 'k_wprep' in r[2]][::2]
    print('# events', len(rows), 'memcpy groups', len(starts))
    per = len(starts) // NS if NS else 1
    b = starts[per * (NS - 2)] if len(starts) >= per * (NS - 1) else 0
    e = starts[per * (NS - 1)] if len(starts) > per * (NS - 1) else len(rows)
    step = rows[b:e]
    base = step[0][0]
    end_prev = base
    print('# one step: %d events, span %.1f us' % (len(step), step[-1][0] + step[-1][1] - base))
    tot = {}
    busy_end = base
    for s, d, n in step:
        gap = s - busy_end
        short = n.split('(')[0].replace('void ', '').replace('bmnas::', '')[:60]
        print('%9.2f  dur %7.2f  gap %6.2f  %s' % (s - base, d, gap, short))
        busy_end = max(busy_end, s + d)
        a = tot.setdefault(short, [0, 0.0])
        a[0] += 1; a[1] += d
    print('# per kernel (one step)')
    for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%8.1f us  n=%3d  avg %6.2f  %s' % (d, n, d / n, k))
    print('# sum of durations %.1f us' % sum(v[1] for v in tot.values()))


if __name__ == '__main__':
    main()
    if int(os.environ.get('WORLD_SIZE', 1)) > 1:
        sys.stdout.flush()
        os._exit(0)
