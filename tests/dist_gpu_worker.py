"""2+-GPU worker (launched by tests/test_gpu_dist.py through torch.distributed.run): the peer-memory fused optimiser step
(bmnas_dp_adam_step: reduce-scatter + Adam on the shard + all-gather of parameters) against (a) the NCCL all-reduce +
FusedAdam path on the same shards and (b) the CPU oracle run on the GLOBAL batch chunk by chunk (per-replica BatchNorm
statistics = nn.DataParallel semantics, summed gradients / world, one Adam step per half)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'bm-nas_b200'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch
import torch.distributed as dist

from helpers import O
import gpu_util as U


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from bmnas.nn import CrossEntropyLoss
    from bmnas.search import SearchStep
    cfg = O.Cfg(32, 8, 4, 2, 2, 2, 2, 0.0)
    ncls, Bl, K = 7, 8, 4
    Bg = Bl * world
    P = O.init_params(cfg, ncls, seed=11, prefix='cell')
    arch = O.init_arch(cfg, seed=11, scale=0.3)
    batches = [O.synthetic_batch(cfg, Bg, ncls, seed=100 + i) for i in range(2 * K)]
    shard = lambda b: (torch.stack([f[rank * Bl:(rank + 1) * Bl] for f in b[0]]), b[1][rank * Bl:(rank + 1) * Bl])

    def run(peer, graphs):
        head = U.build_head(cfg, ncls, P, arch, device=dev)
        head.train()
        for m in head.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        ss = SearchStep(head, CrossEntropyLoss(), Bl, ncls, use_graphs=graphs, group=dist.group.WORLD,
                        peer_step=None if peer else False, nbpe=3.0)
        assert (ss.peer is not None) == peer, 'peer-memory step could not be set up'
        ss.load('dev', *shard(batches[0])); ss.load('train', *shard(batches[1]))
        ss.prepare(warmup=2, restore=True)
        for i in range(K):
            ss.load('dev', *shard(batches[2 * i])); ss.load('train', *shard(batches[2 * i + 1]))
            ss.step()
        torch.cuda.synchronize()
        flat = torch.cat([p.detach().reshape(-1) for p in head.parameters()] + [a.detach().reshape(-1) for a in head.arch_parameters()])
        return flat.clone(), [a.detach().cpu().clone() for a in head.arch_parameters()], head

    ref_flat, ref_arch, _ = run(False, False)
    for graphs in (False, True):
        flat, archs, head = run(True, graphs)
        # (1) replicas bit-identical
        g = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(g, flat)
        assert all(torch.equal(g[0], x) for x in g), 'replicas diverged under the peer-memory step'
        # (2) same trajectory as NCCL all-reduce + FusedAdam (different summation order across ranks only)
        err = (flat - ref_flat).abs().max().item()
        # Adam turns a coordinate whose gradient is rounding noise (BN-fed conv biases) into +-lr steps: bound the gap by 5 % of
        # the distance a coordinate can travel in K steps (K * lr = 4e-3)
        assert err < 0.05 * K * 1e-3, f'peer-memory step vs NCCL path: max abs parameter difference {err}'
    # (3) against the CPU oracle: per-replica BatchNorm statistics, gradients summed over the chunks, / world, Adam
    st_arch = [a.clone() for a in arch]
    Pc = {k: v.clone() for k, v in P.items()}
    names = O.trainable_names(Pc)
    a_state, w_state = {}, {}
    sched = O.CosineRestartLR(1e-3, 1e-6, 1, 2, 3.0)
    zero_cfg = cfg
    for i in range(K):
        for which, b in (('dev', batches[2 * i]), ('train', batches[2 * i + 1])):
            gw_sum, ga_sum = None, None
            bufs = None
            for r in range(world):
                Pr = {k: v.clone() for k, v in Pc.items()}
                lv, _, gw, ga = O.loss_and_grads([f[r * Bl:(r + 1) * Bl] for f in b[0]], b[1][r * Bl:(r + 1) * Bl], st_arch, Pr,
                                                 None, zero_cfg, training=True)
                if r == 0:
                    bufs = {k: v for k, v in Pr.items() if 'running' in k or 'num_batches' in k}
                gw_sum = gw if gw_sum is None else {k: gw_sum[k] + gw[k] for k in gw}
                ga_sum = ga if ga_sum is None else [x + y for x, y in zip(ga_sum, ga)]
            Pc.update(bufs)                      # BatchNorm buffers follow rank 0 (each replica keeps its own on the GPUs)
            if which == 'dev':
                O.adam_step(st_arch, [x / world for x in ga_sum], a_state, 3e-4, (0.5, 0.999), 1e-3)
            else:
                lr = sched.step()
                O.adam_step([Pc[k] for k in names], [gw_sum[k] / world for k in names], w_state, lr, (0.9, 0.999), 3e-4)
    for x, y in zip(archs, st_arch):
        e = (x - y).abs().max().item()
        assert e < 0.1 * K * 3e-4, f'architecture after {K} data-parallel steps vs the chunked CPU oracle: {e}'
    if rank == 0:
        print('DIST_OK world', world, 'peer-vs-nccl max diff', err, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == '__main__':
    main()
