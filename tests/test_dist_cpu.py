"""world_size-2 gloo tests of the data-parallel host logic: one flat gradient bucket, one
all-reduce per half step, 1/world folded into Adam, replicas bit-identical, per-replica BatchNorm statistics
(= nn.DataParallel semantics) checked against the CPU oracle run chunk-by-chunk."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import O, ROOT
import gpu_util as U


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
            if p not in sys.path:
                sys.path.insert(0, p)
        os.environ['MASTER_ADDR'] = '127.0.0.1'
        os.environ['MASTER_PORT'] = str(port)
        dist.init_process_group('gloo', rank=rank, world_size=world)
        torch.set_num_threads(2)
        from bmnas import native as N
        from bmnas.nn import CrossEntropyLoss
        from bmnas.search import SearchStep
        N.set_validate_only(True)

        cfg = O.Cfg(16, 8, 4, 2, 2, 2, 2, 0.0)
        ncls, Bg = 5, 8
        Bl = Bg // world
        P = O.init_params(cfg, ncls, seed=1, prefix='cell')
        arch = O.init_arch(cfg, seed=1, scale=0.3)
        head = U.build_head(cfg, ncls, P, arch, device=torch.device('cpu'))
        if rank != 0:                     # a differently initialised replica: SearchStep must pull rank 0's state
            with torch.no_grad():
                for t in list(head.parameters()) + head.arch_parameters():
                    t.add_(0.5)
        from bmnas import runtime as rt
        ss = SearchStep(head, CrossEntropyLoss(), Bl, ncls, group=dist.group.WORLD, use_graphs=False)
        assert ss.world == world and abs(ss.w_opt.grad_scale - 1.0 / world) < 1e-12
        assert rt.SAMPLE_OFFSET[0] == rank * Bl          # world-size-invariant dropout streams
        for k, v in head.state_dict().items():
            assert torch.equal(v, P[k]), k               # sync_replicas(): rank 0's weights everywhere
        for a_, b_ in zip(head.arch_parameters(), arch):
            assert torch.equal(a_.detach(), b_)
        feats, labels = O.synthetic_batch(cfg, Bg, ncls, seed=2)

        # host path end to end in validate-only mode: plan build, backward wiring, all-reduce, fused Adam call
        ss.load('dev', torch.stack([f[rank * Bl:(rank + 1) * Bl] for f in feats]), labels[rank * Bl:(rank + 1) * Bl])
        ss.load('train', torch.stack([f[rank * Bl:(rank + 1) * Bl] for f in feats]), labels[rank * Bl:(rank + 1) * Bl])
        ss._half('dev')
        arena = head._joint_arena(torch.device('cpu'))
        weights = list(head.fusion_net.parameters()) + list(head.central_classifier.parameters())
        archs = head.arch_parameters()
        # the arch half produces alpha/beta/gamma gradients only (the reference zeroes the weight gradients unread)
        assert all(t.grad is not None and t.grad.data_ptr() == arena.view(t).data_ptr() for t in archs)
        assert all(t.grad is None for t in weights)
        ss._half('train')
        assert all(t.grad is not None and t.grad.data_ptr() == arena.view(t).data_ptr() for t in weights)
        # the two NCCL buckets: [alpha, beta, gamma] and [weights] are disjoint contiguous spans of one arena
        sa, sw = ss.grad_span('dev'), ss.grad_span('train')
        assert sa.data_ptr() == arena.flat.data_ptr() and sa.numel() == sum((a_.numel() + 3) // 4 * 4 for a_ in archs)
        assert sw.data_ptr() == sa.data_ptr() + 4 * sa.numel() and sa.numel() + sw.numel() == arena.flat.numel()

        # per-shard oracle gradients (per-replica BN statistics), written into the arena views
        def shard_grads(r):
            Pc = {k: v.clone() for k, v in P.items()}
            lv, _, gw, ga = O.loss_and_grads([f[r * Bl:(r + 1) * Bl] for f in feats], labels[r * Bl:(r + 1) * Bl],
                                             arch, Pc, None, cfg)
            return gw, ga
        gw, ga = shard_grads(rank)
        names = dict(head.named_parameters())
        with torch.no_grad():
            arena.flat.zero_()
            for k, p_ in names.items():
                arena.view(p_).copy_(gw[k])
            for a_, g_ in zip(head.arch_parameters(), ga):
                arena.view(a_).copy_(g_)
        ss.allreduce_grads('train')                            # ONE collective over the weight span ...
        # expected: sum over shards (the 1/world factor is Adam's grad_scale)
        exp_w = {k: sum(shard_grads(r)[0][k] for r in range(world)) for k in names}
        for k, p_ in names.items():
            assert torch.allclose(arena.view(p_), exp_w[k], rtol=1e-6, atol=1e-8), k
        for a_, g_ in zip(head.arch_parameters(), ga):         # ... which leaves the arch span alone
            assert torch.equal(arena.view(a_), g_)
        ss.allreduce_grads('dev')                              # ... and one over the 70-float arch span
        exp_a = [sum(shard_grads(r)[1][i] for r in range(world)) for i in range(len(ga))]
        for a_, g_ in zip(head.arch_parameters(), exp_a):
            assert torch.allclose(arena.view(a_), g_, rtol=1e-6, atol=1e-8)
        # every rank applies the identical update -> replicas stay bit-identical without a broadcast
        st = {}
        params = [p_.detach() for p_ in names.values()]
        O.adam_step(params, [arena.view(p_) * ss.w_opt.grad_scale for p_ in names.values()], st, 1e-3, (0.9, 0.999), 3e-4)
        flat = torch.cat([p_.reshape(-1) for p_ in params])
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], g_) for g_ in gathered)
        # BatchNorm running statistics follow rank 0 (nn.DataParallel semantics)
        for b in head.buffers():
            if b.dtype.is_floating_point:
                b.add_(float(rank))
        ss.sync_buffers()
        ref = [b.clone() for b in head.buffers()]
        for b in ref:
            dist.broadcast(b, 0)
        assert all(torch.equal(a_, b_) for a_, b_ in zip(head.buffers(), ref))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, 'FAIL: ' + traceback.format_exc()))


@pytest.mark.timeout(300)
def test_dp_world2_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 'ok' for r in res), res
