"""The EgoGesture search / evaluation loop -- drop-in for models/search/train_searchable/ego.py
(train_ego_track_acc :13-181, test_ego_track_acc :183-223): same signatures, phases, log lines and files as the NTU
loop (train_searchable/ntu.py) with the Ego specifics --
  * a batch is (inputs, labels) with inputs (B, 3 + D, T, H, W): channels 0-2 are RGB, the rest depth (ego.py:61-71);
  * the architecture learning rate is logged at the start of every dev phase (ego.py:43-45), the weight learning rate
    at the start of every phase (ego.py:53-55);
  * "dataset_size: n" is printed per phase (ego.py:113), plotter.plot(..., task='ego').
Running loss / correct counts stay on the device and are read once per phase (the reference syncs twice per batch)."""
import copy
import os

import torch

import models.auxiliary.scheduler as sc
from models.search.darts.utils import save, save_pickle

from .ntu import _net, _fusion_params, _batches


def _split(inputs, device):
    inputs = inputs.to(device, non_blocking=True)
    return inputs[:, 0:3], inputs[:, 3:]


def train_ego_track_acc(model, architect, criterion, optimizer, scheduler, dataloaders, dataset_sizes,
                        device=None, num_epochs=200, parallel=False, logger=None, plotter=None, args=None,
                        status='search', verbose=False):
    best_genotype, best_acc, best_epoch = None, 0, 0
    best_test_genotype, best_test_acc, best_test_epoch = None, 0, 0
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
                if architect is not None:
                    architect.log_learning_rate(logger)
                model.train()
            else:
                model.eval()
            for group in optimizer.param_groups:
                logger.info("Learning Rate: {}".format(group['lr']))
                break
            running_loss = torch.zeros((), dtype=torch.float64, device=device)
            running_corrects = torch.zeros((), dtype=torch.int64, device=device)
            it, bar = _batches(dataloaders[phase], verbose)
            for inputs, labels in it:
                rgbs, depths = _split(inputs, device)
                labels = labels.to(device, non_blocking=True)
                input_features = (rgbs, depths)
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
                running_loss += loss.detach().double() * n
                correct = torch.sum(preds == labels)
                running_corrects += correct
                if bar is not None:
                    bar.set_postfix_str('batch_loss: {:.03f}, batch_acc: {:.03f}'.format(loss.item(), correct.item() / n))
            epoch_loss = running_loss.item() / dataset_sizes[phase]
            epoch_acc = running_corrects.double().cpu() / dataset_sizes[phase]
            print('dataset_size:', dataset_sizes[phase])
            logger.info('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, epoch_loss, epoch_acc))
            logger.info("Fusion Model Params: {}".format(_fusion_params(model, parallel)))
            genotype = _net(model, parallel).genotype()
            logger.info(str(genotype))
            if phase == 'dev' and epoch_acc >= best_acc:
                best_acc, best_genotype, best_epoch = epoch_acc, copy.deepcopy(genotype), epoch
                save(_net(model, parallel), os.path.join(args.save, 'best', 'best_model.pt'))
                save_pickle(best_genotype, os.path.join(args.save, 'best', 'best_genotype.pkl'))
            if phase == 'test' and epoch_acc >= best_test_acc:
                best_test_acc, best_test_genotype, best_test_epoch = epoch_acc, copy.deepcopy(genotype), epoch
                save(_net(model, parallel), os.path.join(args.save, 'best', 'best_test_model.pt'))
                save_pickle(best_test_genotype, os.path.join(args.save, 'best', 'best_test_genotype.pkl'))
        if plotter is not None:
            plotter.plot(genotype, os.path.join(args.save, "architectures", "epoch_{}".format(epoch)), task='ego')
        logger.info("Current best dev accuracy: {}, at training epoch: {}".format(best_acc, best_epoch))
        logger.info("Current best test accuracy: {}, at training epoch: {}".format(best_test_acc, best_test_epoch))
    if status == 'search':
        return best_acc, best_genotype
    return best_test_acc, best_genotype


def test_ego_track_acc(model, dataloaders, criterion, genotype, dataset_sizes, device, logger, args):
    model.eval()
    logger.info("EXP: {}".format(args.save))
    phase = 'test'
    running_loss = torch.zeros((), dtype=torch.float64, device=device)
    running_corrects = torch.zeros((), dtype=torch.int64, device=device)
    with torch.no_grad():
        for inputs, labels in dataloaders[phase]:
            rgbs, depths = _split(inputs, device)
            labels = labels.to(device, non_blocking=True)
            output = model((rgbs, depths))
            _, preds = torch.max(output, 1)
            loss = criterion(output, labels)
            running_loss += loss.detach().double() * rgbs.size(0)
            running_corrects += torch.sum(preds == labels)
    test_loss = running_loss.item() / dataset_sizes[phase]
    test_acc = running_corrects.double().cpu() / dataset_sizes[phase]
    logger.info(str(genotype))
    logger.info('{} Loss: {:.4f} Acc: {:.4f}'.format(phase, test_loss, test_acc))
    return test_acc


test_ego_track_acc.__test__ = False      # not a pytest test
