"""The NTU search / evaluation loop -- drop-in for models/search/train_searchable/ntu.py
(train_ntu_track_acc :12-178, test_ntu_track_acc :180-227), same signatures, phases, bookkeeping and files:

  per epoch   status == 'search': phases train (weight steps) then dev (Architect.step + metrics forward)
              otherwise:          phases train then test
  per phase   epoch loss / accuracy, parameter count of reshape layers + fusion net, the derived genotype
  on a new best dev (test) accuracy: best/best_model.pt (best_test_model.pt) = state_dict and
              best/best_genotype.pkl (best_test_genotype.pkl) = pickled Genotype
  per epoch   plotter.plot(genotype, architectures/epoch_<n>)

What changes is where the bookkeeping runs: the reference reads loss.item() and a correct-count back to the host
every batch (two synchronisations per iteration, ntu.py:98-104); here the running sums stay on the device and
are read ONCE per phase, so the host keeps launching while the GPU computes.  tqdm's per-batch postfix is kept
only under ``verbose`` (it needs those host values).  ``plotter`` may be None (graphviz is optional)."""
import copy
import os

import torch

import models.auxiliary.scheduler as sc
from models.search.darts.utils import count_parameters, save, save_pickle


def _net(model, parallel):
    return model.module if parallel else model


def _fusion_params(model, parallel):
    net = _net(model, parallel)
    n = 0
    for reshape_layer in getattr(net, 'reshape_layers', []):
        n += count_parameters(reshape_layer)
    return n + count_parameters(net.fusion_net)


def _batches(loader, verbose):
    if not verbose:
        return loader, None
    from tqdm import tqdm
    t = tqdm(loader)
    return t, t


def train_ntu_track_acc(model, architect, criterion, optimizer, scheduler, dataloaders, dataset_sizes,
                        device=None, num_epochs=200, verbose=False, parallel=False, logger=None,
                        plotter=None, args=None, status='search'):
    best_genotype = None
    best_acc = 0
    best_epoch = 0
    best_test_genotype = None
    best_test_acc = 0
    best_test_epoch = 0
    cosine = isinstance(scheduler, sc.LRCosineAnnealingScheduler)

    for epoch in range(num_epochs):
        logger.info("Epoch: {}".format(epoch))
        logger.info("EXP: {}".format(args.save))
        phases = ['train', 'dev'] if status == 'search' else ['train', 'test']
        genotype = None
        for phase in phases:
            if phase == 'train':
                if not cosine:
                    scheduler.step()
                model.train()
            elif phase == 'dev':
                model.train()
            else:
                model.eval()

            running_loss = torch.zeros((), dtype=torch.float64, device=device)
            running_corrects = torch.zeros((), dtype=torch.int64, device=device)
            it, bar = _batches(dataloaders[phase], verbose)
            for data in it:
                rgbs = data['rgb'].to(device, non_blocking=True)
                skes = data['ske'].to(device, non_blocking=True)
                labels = data['label'].to(device, non_blocking=True)
                input_features = (rgbs, skes)
                if status == 'search' and (phase == 'dev' or phase == 'test'):
                    if architect is not None:
                        architect.step(input_features, labels, logger)
                optimizer.zero_grad()
                grad_phase = phase == 'train' or (phase == 'dev' and status == 'eval')
                with torch.set_grad_enabled(grad_phase):
                    output = model(input_features)
                    _, preds = torch.max(output, 1)
                    loss = criterion(output, labels)
                    if grad_phase:
                        if cosine:
                            scheduler.step()
                            scheduler.update_optimizer(optimizer)
                        loss.backward()
                        optimizer.step()
                n = rgbs.size(0)
                running_loss += loss.detach().double() * n          # stays on the device: no host sync here
                correct = torch.sum(preds == labels)
                running_corrects += correct
                if bar is not None:
                    bar.set_postfix_str('batch_loss: {:.03f}, batch_acc: {:.03f}'.format(loss.item(), correct.item() / n))

            epoch_loss = running_loss.item() / dataset_sizes[phase]           # the phase's one synchronisation
            epoch_acc = running_corrects.double().cpu() / dataset_sizes[phase]
            logger.info('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, epoch_loss, epoch_acc))
            logger.info("Fusion Model Params: {}".format(_fusion_params(model, parallel)))
            genotype = _net(model, parallel).genotype()
            logger.info(str(genotype))

            if phase == 'dev' and epoch_acc >= best_acc:
                best_acc = epoch_acc
                best_genotype = copy.deepcopy(genotype)
                best_epoch = epoch
                save(_net(model, parallel), os.path.join(args.save, 'best', 'best_model.pt'))
                save_pickle(best_genotype, os.path.join(args.save, 'best', 'best_genotype.pkl'))
            if phase == 'test' and epoch_acc >= best_test_acc:
                best_test_acc = epoch_acc
                best_test_genotype = copy.deepcopy(genotype)
                best_test_epoch = epoch
                save(_net(model, parallel), os.path.join(args.save, 'best', 'best_test_model.pt'))
                save_pickle(best_test_genotype, os.path.join(args.save, 'best', 'best_test_genotype.pkl'))

        if plotter is not None:
            plotter.plot(genotype, os.path.join(args.save, "architectures", "epoch_{}".format(epoch)))
        logger.info("Current best dev accuracy: {}, at training epoch: {}".format(best_acc, best_epoch))
        logger.info("Current best test accuracy: {}, at training epoch: {}".format(best_test_acc, best_test_epoch))

    if status == 'search':
        return best_acc, best_genotype
    return best_test_acc, best_genotype


def test_ntu_track_acc(model, dataloaders, criterion, genotype, dataset_sizes, device, logger, args):
    model.eval()
    logger.info("EXP: {}".format(args.save))
    phase = 'test'
    running_loss = torch.zeros((), dtype=torch.float64, device=device)
    running_corrects = torch.zeros((), dtype=torch.int64, device=device)
    with torch.no_grad():
        for data in dataloaders[phase]:
            rgbs = data['rgb'].to(device, non_blocking=True)
            skes = data['ske'].to(device, non_blocking=True)
            labels = data['label'].to(device, non_blocking=True)
            output = model((rgbs, skes))
            _, preds = torch.max(output, 1)
            loss = criterion(output, labels)
            running_loss += loss.detach().double() * rgbs.size(0)
            running_corrects += torch.sum(preds == labels)
    test_loss = running_loss.item() / dataset_sizes[phase]
    test_acc = running_corrects.double().cpu() / dataset_sizes[phase]
    logger.info(str(genotype))
    logger.info('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, test_loss, test_acc))
    return test_acc


test_ntu_track_acc.__test__ = False      # not a pytest test
