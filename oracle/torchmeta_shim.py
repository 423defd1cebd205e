"""Restatement of the parts of torchmeta==1.7.0 the reference imports (TEST INFRASTRUCTURE).

torchmeta is a pinned third-party dependency of the reference (requirements.txt:10) whose
source is NOT under /root/reference and which is not installed here.  This module restates,
from its published algorithm, exactly the symbols the reference binds:

  fumi/models/fumi.py:5-6   MetaSequential, MetaLinear, gradient_update_parameters
  fumi/models/maml.py:8-9   MetaModule, MetaSequential, MetaLinear, gradient_update_parameters
  fumi/dataset/data.py:13,17-19  datasets.helpers, Categorical, ClassSplitter,
                                 BatchMetaDataLoader, ClassDataset, CombinationMetaDataset, Dataset

PARITY UNPINNED against the real package (no copy is available offline); the behaviour
restated here is the one SURVEY.md Appendix B / C.7 records.
"""
import random
import sys
import types
import warnings
from collections import OrderedDict, defaultdict
from copy import deepcopy
from itertools import combinations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.data import ConcatDataset, DataLoader
from torch.utils.data import Dataset as TorchDataset
from torch.utils.data.dataloader import default_collate
from torch.utils.data.sampler import RandomSampler


# ----------------------------------------------------------------------------- modules
class MetaModule(nn.Module):
    """nn.Module whose forward accepts an explicit ``params`` OrderedDict."""

    def meta_named_parameters(self, prefix="", recurse=True):
        gen = self._named_members(
            lambda module: module._parameters.items() if isinstance(module, MetaModule) else [],
            prefix=prefix, recurse=recurse)
        for elem in gen:
            yield elem

    def meta_parameters(self, recurse=True):
        for _, param in self.meta_named_parameters(recurse=recurse):
            yield param

    def get_subdict(self, params, key=None):
        if params is None:
            return None
        if key is None:
            return params
        pre = key + "."
        sub = OrderedDict((k[len(pre):], v) for k, v in params.items() if k.startswith(pre))
        return sub if len(sub) else None


class MetaLinear(nn.Linear, MetaModule):
    def forward(self, input, params=None):
        if params is None:
            params = OrderedDict(self.named_parameters())
        bias = params.get("bias", None)
        return F.linear(input, params["weight"], bias)


class MetaSequential(nn.Sequential, MetaModule):
    def forward(self, input, params=None):
        for name, module in self._modules.items():
            if isinstance(module, MetaModule):
                input = module(input, params=self.get_subdict(params, name))
            elif isinstance(module, nn.Module):
                input = module(input)
            else:
                raise TypeError(type(module))
        return input


def gradient_update_parameters(model, loss, params=None, step_size=0.5, first_order=False):
    if not isinstance(model, MetaModule):
        raise ValueError("model must be a MetaModule")
    if params is None:
        params = OrderedDict(model.meta_named_parameters())
    grads = torch.autograd.grad(loss, params.values(), create_graph=not first_order)
    updated = OrderedDict()
    if isinstance(step_size, (dict, OrderedDict)):
        for (name, param), grad in zip(params.items(), grads):
            updated[name] = param - step_size[name] * grad
    else:
        for (name, param), grad in zip(params.items(), grads):
            updated[name] = param - step_size * grad
    return updated


# ----------------------------------------------------------------------------- data
class Categorical(object):
    """Lazy label permutation: k-th distinct raw target seen -> torch.randperm(N)[k]."""

    def __init__(self, num_classes=None):
        self.num_classes = num_classes
        self._classes = None
        self._labels = None

    def reset(self):
        self._classes = None
        self._labels = None

    @property
    def classes(self):
        if self._classes is None:
            self._classes = defaultdict(None)
            if self.num_classes is None:
                self._classes.default_factory = lambda: len(self._classes)
            else:
                self._classes.default_factory = lambda: self.labels[len(self._classes)]
        if (self.num_classes is not None) and (len(self._classes) > self.num_classes):
            raise ValueError("more classes than num_classes")
        return self._classes

    @property
    def labels(self):
        if (self._labels is None) and (self.num_classes is not None):
            self._labels = torch.randperm(self.num_classes).tolist()
        return self._labels

    def __call__(self, target):
        return self.classes[target]


class _Compose(object):
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


class Dataset(TorchDataset):
    def __init__(self, index, transform=None, target_transform=None):
        self.index = index
        self.transform = transform
        self.target_transform = target_transform

    def target_transform_append(self, transform):
        if transform is None:
            return
        if self.target_transform is None:
            self.target_transform = transform
        else:
            self.target_transform = _Compose([self.target_transform, transform])

    def __hash__(self):
        return hash(self.index)


class ClassDataset(object):
    def __init__(self, meta_train=False, meta_val=False, meta_test=False, meta_split=None,
                 class_augmentations=None):
        if meta_train + meta_val + meta_test == 0 and meta_split is None:
            raise ValueError("one of meta_train/meta_val/meta_test must be set")
        self.meta_train, self.meta_val, self.meta_test = meta_train, meta_val, meta_test
        self.class_augmentations = class_augmentations or []

    def get_class_augmentation(self, index):
        return None

    def get_transform(self, index, transform=None):
        return transform

    def get_target_transform(self, index):
        return self.get_class_augmentation(index)

    def __getitem__(self, index):
        raise NotImplementedError()

    @property
    def num_classes(self):
        raise NotImplementedError()

    def __len__(self):
        return self.num_classes + len(self.class_augmentations) * self.num_classes


class Task(Dataset):
    def __init__(self, index, num_classes, transform=None, target_transform=None):
        super().__init__(index, transform=transform, target_transform=target_transform)
        self.num_classes = num_classes


class ConcatTask(Task, ConcatDataset):
    def __init__(self, datasets, num_classes, target_transform=None):
        index = tuple(task.index for task in datasets)
        Task.__init__(self, index, num_classes)
        ConcatDataset.__init__(self, datasets)
        for task in self.datasets:
            task.target_transform_append(target_transform)

    def __getitem__(self, index):
        return ConcatDataset.__getitem__(self, index)


class SubsetTask(Task):
    def __init__(self, dataset, indices, num_classes=None, target_transform=None):
        if num_classes is None:
            num_classes = dataset.num_classes
        super().__init__(dataset.index, num_classes, target_transform=target_transform)
        self.dataset = dataset
        self.indices = indices

    def __getitem__(self, index):
        return self.dataset[self.indices[index]]

    def __len__(self):
        return len(self.indices)


class CombinationMetaDataset(object):
    def __init__(self, dataset, num_classes_per_task, target_transform=None, dataset_transform=None):
        if not isinstance(num_classes_per_task, int):
            raise TypeError("num_classes_per_task must be int")
        self.dataset = dataset
        self.num_classes_per_task = num_classes_per_task
        self.target_transform = target_transform
        self.dataset_transform = dataset_transform
        self.seed()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed=seed)
        if hasattr(self.dataset_transform, "seed"):
            self.dataset_transform.seed(seed)

    def __iter__(self):
        for index in combinations(range(len(self.dataset)), self.num_classes_per_task):
            yield self[index]

    def __getitem__(self, index):
        if isinstance(index, int):
            raise ValueError("index of a CombinationMetaDataset must be a tuple")
        assert len(index) == self.num_classes_per_task
        datasets = [self.dataset[i] for i in index]
        tt = self.target_transform
        if isinstance(tt, Categorical):
            tt.reset()
            if tt.num_classes is None:
                tt.num_classes = self.num_classes_per_task
            tt = deepcopy(tt)
        task = ConcatTask(datasets, self.num_classes_per_task, target_transform=tt)
        if self.dataset_transform is not None:
            task = self.dataset_transform(task)
        return task

    def __len__(self):
        num_classes, length = len(self.dataset), 1
        for i in range(1, self.num_classes_per_task + 1):
            length *= (num_classes - i + 1) / i
        if length > sys.maxsize:
            length = sys.maxsize
        return int(length)


class ClassSplitter_(object):
    def __init__(self, shuffle=True, num_samples_per_class=None, num_train_per_class=None,
                 num_test_per_class=None, random_state_seed=0):
        self.shuffle = shuffle
        if num_samples_per_class is None:
            num_samples_per_class = OrderedDict()
            if num_train_per_class is not None:
                num_samples_per_class["train"] = num_train_per_class
            if num_test_per_class is not None:
                num_samples_per_class["test"] = num_test_per_class
        self.splits = num_samples_per_class
        self._min_samples_per_class = sum(num_samples_per_class.values())
        self.random_state_seed = random_state_seed
        self.seed(random_state_seed)

    def seed(self, seed):
        self.np_random = np.random.RandomState(seed=seed)

    def get_indices_concattask(self, task):
        indices = OrderedDict([(split, []) for split in self.splits])
        cum_size = 0
        for dataset in task.datasets:
            num_samples = len(dataset)
            if num_samples < self._min_samples_per_class:
                raise ValueError("The number of samples for one class ({0}) is smaller than the "
                                 "minimum number of samples per class required ({1})."
                                 .format(num_samples, self._min_samples_per_class))
            if self.shuffle:
                seed = (hash(task) + hash(dataset) + self.random_state_seed) % (2 ** 32)
                dataset_indices = np.random.RandomState(seed).permutation(num_samples)
            else:
                dataset_indices = np.arange(num_samples)
            ptr = 0
            for split, num_split in self.splits.items():
                split_indices = dataset_indices[ptr:ptr + num_split]
                if self.shuffle:
                    self.np_random.shuffle(split_indices)
                indices[split].extend(split_indices + cum_size)
                ptr += num_split
            cum_size += num_samples
        return indices

    def __call__(self, task):
        if not isinstance(task, ConcatTask):
            raise ValueError("only ConcatTask is restated")
        indices = self.get_indices_concattask(task)
        return OrderedDict([(split, SubsetTask(task, indices[split])) for split in self.splits])


def ClassSplitter(task=None, *args, **kwargs):
    splitter = ClassSplitter_(*args, **kwargs)
    if task is None:
        return splitter
    if isinstance(task, CombinationMetaDataset):
        task.dataset_transform = splitter
        return task
    return splitter(task)


class CombinationRandomSampler(RandomSampler):
    def __init__(self, data_source):
        if not isinstance(data_source, CombinationMetaDataset):
            raise TypeError("expected a CombinationMetaDataset")
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            super().__init__(data_source, replacement=True)

    def __iter__(self):
        num_classes = len(self.data_source.dataset)
        n = self.data_source.num_classes_per_task
        for _ in combinations(range(num_classes), n):
            yield tuple(random.sample(range(num_classes), n))


class BatchMetaCollate(object):
    def __init__(self, collate_fn):
        self.collate_fn = collate_fn

    def collate_task(self, task):
        if isinstance(task, TorchDataset):
            return self.collate_fn([task[idx] for idx in range(len(task))])
        elif isinstance(task, OrderedDict):
            return OrderedDict([(k, self.collate_task(sub)) for (k, sub) in task.items()])
        raise NotImplementedError()

    def __call__(self, batch):
        return self.collate_fn([self.collate_task(task) for task in batch])


class BatchMetaDataLoader(DataLoader):
    def __init__(self, dataset, batch_size=1, shuffle=True, sampler=None, num_workers=0,
                 pin_memory=False, drop_last=False, timeout=0, worker_init_fn=None):
        if isinstance(dataset, CombinationMetaDataset) and sampler is None:
            if not shuffle:
                raise NotImplementedError("sequential combination sampler not restated")
            sampler = CombinationRandomSampler(dataset)
            shuffle = False
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, sampler=sampler,
                         batch_sampler=None, num_workers=num_workers,
                         collate_fn=BatchMetaCollate(default_collate), pin_memory=pin_memory,
                         drop_last=drop_last, timeout=timeout, worker_init_fn=worker_init_fn)


# ----------------------------------------------------------------------------- install
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = __import__("importlib.machinery").machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Register the restated symbols under the import names the reference uses."""
    if "torchmeta" in sys.modules and getattr(sys.modules["torchmeta"], "_fumi_shim", False):
        return
    tm = _mod("torchmeta", _fumi_shim=True)
    tm.modules = _mod("torchmeta.modules", MetaModule=MetaModule, MetaLinear=MetaLinear,
                      MetaSequential=MetaSequential)
    tm.utils = _mod("torchmeta.utils")
    tm.utils.gradient_based = _mod("torchmeta.utils.gradient_based",
                                   gradient_update_parameters=gradient_update_parameters)
    tm.utils.data = _mod("torchmeta.utils.data", BatchMetaDataLoader=BatchMetaDataLoader,
                         ClassDataset=ClassDataset, CombinationMetaDataset=CombinationMetaDataset,
                         Dataset=Dataset)
    tm.transforms = _mod("torchmeta.transforms", Categorical=Categorical, ClassSplitter=ClassSplitter)
    tm.datasets = _mod("torchmeta.datasets")
    tm.datasets.helpers = _mod("torchmeta.datasets.helpers")
