"""Host-side helpers of the search loop -- drop-in for the parts of models/search/darts/utils.py the path uses
(AvgrageMeter :9-21, accuracy :23-35, count_parameters[_in_MB] :77-81, save / load :90-94, save_pickle /
load_pickle :96-105, create_exp_dir :115-128).  The CIFAR transforms, Cutout and drop_path of that file belong to
the image-classification DARTS it was copied from and are not part of the multimodal search path."""
import os
import pickle
import shutil

import numpy as np
import torch


class AvgrageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.avg = 0
        self.sum = 0
        self.cnt = 0

    def update(self, val, n=1):
        self.sum += val * n
        self.cnt += n
        self.avg = self.sum / self.cnt


def accuracy(output, target, topk=(1,)):
    maxk = max(topk)
    batch_size = target.size(0)
    _, pred = output.topk(maxk, 1, True, True)
    pred = pred.t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [correct[:k].reshape(-1).float().sum(0).mul_(100.0 / batch_size) for k in topk]


def count_parameters(model):
    return int(sum(int(np.prod(v.size())) for name, v in model.named_parameters() if 'auxiliary' not in name))


def count_parameters_in_MB(model):
    return count_parameters(model) / 1e6


def save(model, model_path):
    torch.save(model.state_dict(), model_path)


def load(model, model_path):
    model.load_state_dict(torch.load(model_path))


def save_pickle(obj, obj_path):
    with open(obj_path, 'wb') as f:
        pickle.dump(obj, f)


def load_pickle(obj_path):
    with open(obj_path, 'rb') as f:
        return pickle.load(f)


def create_exp_dir(path, scripts_to_save=None):
    if not os.path.exists(path):
        os.makedirs(path)
    print('Experiment dir : {}'.format(path))
    if scripts_to_save is not None:
        os.mkdir(os.path.join(path, 'scripts'))
        for script in scripts_to_save:
            shutil.copyfile(script, os.path.join(path, 'scripts', os.path.basename(script)))
    os.mkdir(os.path.join(path, 'architectures'))
    os.mkdir(os.path.join(path, 'best'))
