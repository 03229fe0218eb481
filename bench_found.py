#!/usr/bin/env python
"""bench_found.py -- BASELINE.json configs[4]: found (fixed-genotype) NTU fusion network, train and inference
throughput over batch size on one B200.   python bench_found.py [--batches 96,1024,8192] [--steps 30]

The genotype is the NTU genotype the reference ships (visualize.ipynb:618; models/search/darts/model.py:92-190,
node.py:8-91 build the network from it).  Train step = forward + CE loss + backward + fused Adam; inference =
eval-mode forward (BatchNorm on running statistics, no dropout).  Each is captured into one CUDA graph and replayed;
time = CUDA events around the replays.  Prints one JSON line per batch size (NOT the driver's bench line: that is
bench.py).  Synthetic unit-normal features, random-init weights."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'bm-nas_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def golden_genotype():
    from models.search.darts.genotypes import Genotype, StepGenotype
    return Genotype(edges=[('skip', 2), ('skip', 7), ('skip', 2), ('skip', 3)],
                    steps=[StepGenotype(inner_edges=[('skip', 0), ('skip', 1), ('skip', 2), ('skip', 0)],
                                        inner_steps=['LinearGLU', 'LinearGLU'], inner_concat=[2, 3]),
                           StepGenotype(inner_edges=[('skip', 0), ('skip', 1), ('skip', 2), ('skip', 0)],
                                        inner_steps=['ScaleDotAttn', 'ScaleDotAttn'], inner_concat=[2, 3])],
                    concat=[8, 9])


def graph_ms(fn, steps, warm=3):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def sweep(dev, batches, steps, emit=None):
    from bmnas.nn import SearchHead, CrossEntropyLoss
    from bmnas.optim import FusedAdam
    rows = {}
    a = types.SimpleNamespace(C=128, L=8, num_input_nodes=8, steps=2, multiplier=2, node_steps=2, node_multiplier=2,
                              drpt=0.2, weight_decay=1e-4)
    for B in batches:
        torch.manual_seed(2)
        crit = CrossEntropyLoss()
        head = SearchHead(a, 60, criterion=crit, genotype=golden_genotype()).to(dev)
        opt = FusedAdam(head.central_params(), lr=3e-4, weight_decay=1e-4)
        g = torch.Generator().manual_seed(3)
        feats = [torch.randn(B, a.C, a.L, generator=g).to(dev) for _ in range(a.num_input_nodes)]
        labels = torch.randint(0, 60, (B,), generator=g).to(dev)

        def train_step():
            loss = crit(head(feats), labels)
            loss.backward()
            opt.step()
            return loss.detach()
        head.train()
        t_train = graph_ms(train_step, steps)
        head.eval()

        def infer():
            with torch.no_grad():
                return head(feats)
        t_inf = graph_ms(infer, steps)
        n_w = sum(p.numel() for p in head.parameters())
        row = {'workload': 'found NTU fusion network (golden genotype), synthetic features', 'B': B,
               'weights': n_w, 'train_ms': round(t_train, 4), 'train_samples_per_s': round(B / t_train * 1e3, 1),
               'infer_ms': round(t_inf, 4), 'infer_samples_per_s': round(B / t_inf * 1e3, 1),
               'dtype': 'f32', 'cuda_graphs': True}
        rows[f'B{B}'] = row
        if emit:
            emit(row)
        del head, opt, feats
        torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='96,1024,8192,32768')
    ap.add_argument('--steps', type=int, default=30)
    args = ap.parse_args()
    sweep(torch.device('cuda:0'), [int(b) for b in args.batches.split(',')], args.steps,
          emit=lambda row: print(json.dumps(row), flush=True))


if __name__ == '__main__':
    main()
