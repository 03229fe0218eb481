"""Helpers for the -m gpu parity tests: build the CUDA-path modules from golden fixtures / oracle parameters."""
import types

import torch
import torch.nn as nn

from helpers import O, sub, cfg_of, arch_of, unpickle_genotype

DEV = torch.device('cuda:0') if torch.cuda.is_available() else None


def args_of(cfg, weight_decay=3e-4):
    return types.SimpleNamespace(C=cfg.C, L=cfg.L, drpt=cfg.drpt, num_input_nodes=cfg.num_input_nodes,
                                 steps=cfg.steps, multiplier=cfg.multiplier, node_steps=cfg.node_steps,
                                 node_multiplier=cfg.node_multiplier, weight_decay=weight_decay, parallel=False)


def to_product_genotype(g):
    from models.search.darts.genotypes import Genotype, StepGenotype
    if g is None:
        return None
    return Genotype(edges=list(g.edges), concat=list(g.concat),
                    steps=[StepGenotype(list(s.inner_edges), list(s.inner_steps), list(s.inner_concat)) for s in g.steps])


def build_head(cfg, num_classes, P, arch=None, genotype=None, device=None):
    from bmnas.nn import SearchHead
    device = device or DEV
    head = SearchHead(args_of(cfg), num_classes, genotype=to_product_genotype(genotype))
    missing = head.load_state_dict({k: v.clone() for k, v in P.items()}, strict=True)
    head.to(device)
    if arch is not None:
        with torch.no_grad():
            for a, b in zip(head.arch_parameters(), arch):
                a.copy_(b.to(device))
    return head


def inject_masks(root, masks, device=None):
    device = device or DEV
    n = 0
    for name, m in root.named_modules():
        if isinstance(m, nn.Dropout):
            if masks is not None and name in masks:
                m.injected_mask = masks[name].to(device=device, dtype=torch.uint8).contiguous()
                n += 1
            elif hasattr(m, 'injected_mask'):
                del m.injected_mask
    return n


def random_masks(root, B, C, L, seed, drpt):
    """keep-masks for every live dropout site of a SearchHead (names as in named_modules)"""
    g = torch.Generator().manual_seed(seed)
    masks = {}
    for name, m in root.named_modules():
        if isinstance(m, nn.Dropout) and not name.endswith('node_cell.dropout') and m.p > 0:
            masks[name] = (torch.rand(B, C, L, generator=g) >= m.p).to(torch.uint8)
    return masks


def grads_by_name(head):
    return {n: (p.grad.detach().cpu().clone() if p.grad is not None else None) for n, p in head.named_parameters()}


def training_program(head, mode=None):
    """the launch plan the last training-mode forward of head.fusion_net ran (mode: runtime.GRAD_MODE of that plan)"""
    progs = [(k, r.prog) for k, r in head.fusion_net._bm_cache.items()
             if r.prog.training and not r.prog.use_masks and r.prog.want_backward]
    if mode is not None:
        progs = [(k, p) for k, p in progs if mode in k]
    assert len(progs) == 1, [k for k, _ in progs]
    return progs[0][1]


def philox_masks(head, B, prog):
    """keep-masks the kernels drew in the LAST forward/backward of `prog` (Philox mode), read back through the C-ABI
    test hook bmnas_philox_keep_mask with the plan's own (seed, step); keyed like named_modules() so the oracle can
    consume them (prefix 'fusion_net.')."""
    import ctypes
    from bmnas import native as N
    from bmnas.program import uid_of
    assert prog.rng_state is not None, 'plan has no Philox dropout site'
    C, L = prog.C, prog.L
    masks = {}
    for name, m in head.named_modules():
        if not isinstance(m, nn.Dropout) or name.endswith('node_cell.dropout') or m.p <= 0:
            continue
        site = name[len('fusion_net.'):] if name.startswith('fusion_net.') else name
        key = site if site.endswith('out_dropout') else site[:-len('.dropout')]
        out = torch.empty(B, C, L, dtype=torch.uint8, device=prog.device)
        N.launch('bmnas_philox_keep_mask', ctypes.c_void_p(prog.rng_state.data_ptr()), ctypes.c_uint(uid_of(key)),
                 ctypes.c_float(m.p), ctypes.c_longlong(prog.sample_offset), ctypes.c_longlong(C * L),
                 ctypes.c_longlong(B), ctypes.c_void_p(out.data_ptr()), N.current_stream())
        masks[name] = out.cpu()
    torch.cuda.synchronize()
    return masks
