"""The MM-IMDB search / evaluation loop -- drop-in for models/search/train_searchable/mmimdb.py
(train_mmimdb_track_f1 :10-209, test_mmimdb_track_f1 :211-284): same signatures, phases, log lines, files and failsafes:

  phases      search: train, dev (Architect.step + no-grad metrics forward)     eval: train, dev (trained on), test
  metric      multi-label F1 (sklearn f1_score, average=f1_type, zero_division=1) of sigmoid(output) > th_fscore over the
              whole phase; best dev F1 -> best/best_model.pt + best_genotype.pkl, best test F1 -> best_test_*
  failsafes   NaN epoch loss in the train phase -> model.eval() and return best_f1 (mmimdb.py:150-153);
              NaN best F1 after a single-epoch run -> one more pass over the epochs, once (mmimdb.py:191-200)

Bookkeeping moved off the critical path: the reference copies predictions and labels to the host and calls sklearn on
EVERY batch (for a progress-bar string, mmimdb.py:99-111); here thresholded predictions and labels stay on the device
as bool tensors, the running loss is a device scalar, and each phase synchronises once, for its one F1 call.  The
per-batch F1 postfix is computed only under ``verbose``."""
import copy
import os

import torch

import models.auxiliary.scheduler as sc
from models.search.darts.utils import save, save_pickle

from .ntu import _net, _fusion_params, _batches


def _f1(y_true, y_pred, average, **kw):
    from sklearn.metrics import f1_score
    return f1_score(y_true, y_pred, average=average, **kw)


def train_mmimdb_track_f1(model, architect, criterion, optimizer, scheduler, dataloaders, dataset_sizes, device,
                          num_epochs, parallel, logger, plotter, args, f1_type='weighted', init_f1=0.0, th_fscore=0.3,
                          status='search', verbose=False):
    best_genotype, best_f1, best_epoch = None, init_f1, 0
    best_test_genotype, best_test_f1, best_test_epoch = None, init_f1, 0
    cosine = isinstance(scheduler, sc.LRCosineAnnealingScheduler)
    failsafe, cont_overloop = True, 0
    while failsafe:
        for epoch in range(num_epochs):
            logger.info('Epoch: {}'.format(epoch))
            logger.info("EXP: {}".format(args.save))
            phases = ['train', 'dev'] if status == 'search' else ['train', 'dev', 'test']
            genotype = None
            for phase in phases:
                if phase == 'train':
                    if not cosine:
                        scheduler.step()
                    if architect is not None:
                        architect.log_learning_rate(logger)
                    model.train()
                elif phase == 'dev':
                    if status == 'eval' and not cosine:
                        scheduler.step()
                    model.train()
                else:
                    model.eval()
                list_preds, list_label = [], []
                running_loss = torch.zeros((), dtype=torch.float64, device=device)
                it, bar = _batches(dataloaders[phase], verbose)
                for data in it:
                    image = data['image'].to(device, non_blocking=True)
                    text = data['text'].to(device, non_blocking=True)
                    label = data['label'].to(device, non_blocking=True)
                    if status == 'search' and (phase == 'dev' or phase == 'test'):
                        architect.step((text, image), label, logger)
                    optimizer.zero_grad()
                    grad_phase = phase == 'train' or (phase == 'dev' and status == 'eval')
                    with torch.set_grad_enabled(grad_phase):
                        output = model((text, image))
                        if isinstance(output, tuple):
                            output = output[-1]
                        loss = criterion(output, label)
                        preds_th = torch.sigmoid(output.detach()) > th_fscore
                        if grad_phase:
                            if cosine:
                                scheduler.step()
                                scheduler.update_optimizer(optimizer)
                            loss.backward()
                            optimizer.step()
                    list_preds.append(preds_th)                       # device bool tensors: gathered once per phase
                    list_label.append(label.detach())
                    running_loss += loss.detach().double() * image.size(0)
                    if bar is not None:
                        bf1 = _f1(preds_th.cpu().numpy(), label.cpu().numpy(), f1_type, zero_division=1)
                        bar.set_postfix_str('batch_loss: {:.03f}, batch_f1: {:.03f}'.format(loss.item(), bf1))
                epoch_loss = running_loss.item() / dataset_sizes[phase]   # the phase's synchronisation point
                y_pred = torch.cat(list_preds, dim=0).cpu().numpy()
                y_true = torch.cat(list_label, dim=0).cpu().numpy()
                epoch_f1 = _f1(y_true, y_pred, f1_type, zero_division=1)
                logger.info('{} Loss: {:.4f}, {} F1: {:.4f}'.format(phase, epoch_loss, f1_type, epoch_f1))
                logger.info("Fusion Model Params: {}".format(_fusion_params(model, parallel)))
                genotype = _net(model, parallel).genotype()
                logger.info(str(genotype))
                if phase == 'train' and epoch_loss != epoch_loss:
                    logger.info("Nan loss during training, escaping")
                    model.eval()
                    return best_f1
                if phase == 'dev' and status == 'search' and epoch_f1 > best_f1:
                    best_f1, best_genotype, best_epoch = epoch_f1, copy.deepcopy(genotype), epoch
                    save(_net(model, parallel), os.path.join(args.save, 'best', 'best_model.pt'))
                    save_pickle(best_genotype, os.path.join(args.save, 'best', 'best_genotype.pkl'))
                if phase == 'test' and epoch_f1 > best_test_f1:
                    best_test_f1, best_test_genotype, best_test_epoch = epoch_f1, copy.deepcopy(genotype), epoch
                    save(_net(model, parallel), os.path.join(args.save, 'best', 'best_test_model.pt'))
                    save_pickle(best_test_genotype, os.path.join(args.save, 'best', 'best_test_genotype.pkl'))
            if plotter is not None:
                plotter.plot(genotype, os.path.join(args.save, "architectures", "epoch_{}".format(epoch)), task='mmimdb')
            logger.info("Current best dev {} F1: {}, at training epoch: {}".format(f1_type, best_f1, best_epoch))
            logger.info("Current best test {} F1: {}, at training epoch: {}".format(f1_type, best_test_f1, best_test_epoch))
        if best_f1 != best_f1 and num_epochs == 1 and cont_overloop < 1:
            failsafe = True
            logger.info('Recording a NaN F1, training for one more epoch.')
        else:
            failsafe = False
        cont_overloop += 1
    if best_f1 != best_f1:
        best_f1 = 0.0
    if status == 'search':
        return best_f1, best_genotype
    return best_test_f1, best_test_genotype


def test_mmimdb_track_f1(model, criterion, dataloaders, dataset_sizes, device, parallel, logger, args,
                         f1_type='weighted', init_f1=0.0, th_fscore=0.3):
    model.eval()
    list_preds, list_label = [], []
    running_loss = torch.zeros((), dtype=torch.float64, device=device)
    phase = 'test'
    with torch.no_grad():
        for data in dataloaders[phase]:
            image = data['image'].to(device, non_blocking=True)
            text = data['text'].to(device, non_blocking=True)
            label = data['label'].to(device, non_blocking=True)
            output = model((text, image))
            if isinstance(output, tuple):
                output = output[-1]
            loss = criterion(output, label)
            list_preds.append(torch.sigmoid(output) > th_fscore)
            list_label.append(label)
            running_loss += loss.double() * image.size(0)
    epoch_loss = running_loss.item() / dataset_sizes[phase]
    y_pred = torch.cat(list_preds, dim=0).cpu().numpy()
    y_true = torch.cat(list_label, dim=0).cpu().numpy()
    epoch_f1 = _f1(y_true, y_pred, f1_type)
    logger.info('{} Loss: {:.4f}, {} F1: {:.4f}'.format(phase, epoch_loss, f1_type, epoch_f1))
    logger.info("Fusion Model Params: {}".format(_fusion_params(model, parallel)))
    genotype = _net(model, parallel).genotype()
    logger.info(str(genotype))
    return epoch_f1


test_mmimdb_track_f1.__test__ = False      # not a pytest test
