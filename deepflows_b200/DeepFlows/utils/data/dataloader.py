"""Host-side batch sampler / loader over in-memory numpy arrays
(reference: DeepFlows/utils/data/dataloader.py:5-139). Indexing a dataset with a list of indices
yields a whole batch; `prefetch_size` batches are materialised ahead of the consumer."""
from collections import deque

import numpy as np
from numpy.random import permutation

from .dataset import Dataset


class Sampler:
    def __init__(self, dataset):
        self.dataset = dataset

    def __iter__(self):
        raise NotImplementedError

    def __len__(self):
        return len(self.dataset)


class SequentialSampler(Sampler):
    def __iter__(self):
        return iter(range(len(self.dataset)))


class RandomSampler(Sampler):
    def __iter__(self):
        return iter(permutation(len(self.dataset)).tolist())


class BatchSampler(Sampler):
    def __init__(self, sampler, batch_size, drop_last):
        super().__init__(sampler)
        self.sampler, self.batch_size, self.drop_last = sampler, batch_size, drop_last

    def __iter__(self):
        batch = []
        for idx in self.sampler:
            batch.append(idx)
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        if batch and not self.drop_last:
            yield batch

    def __len__(self):
        n = len(self.sampler)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


class _DataLoaderIter:
    def __init__(self, loader):
        self.loader = loader
        self.indices = iter(loader.batch_sampler)
        self.ready = deque()
        self._prefetch()

    def _load(self, index):
        x, y = self.loader.dataset[index]
        if self.loader.as_contiguous:
            x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        return x, y

    def _prefetch(self):
        while len(self.ready) < self.loader.prefetch_size:
            try:
                self.ready.append(self._load(next(self.indices)))
            except StopIteration:
                break

    def __iter__(self):
        return self

    def __next__(self):
        if self.ready:
            batch = self.ready.popleft()
            self._prefetch()
            return batch
        return self._load(next(self.indices))


class DataLoader:
    def __init__(self, dataset, batch_size=1, shuffle=False, drop_last=False, prefetch_size: int = 0,
                 as_contiguous: bool = True):
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.prefetch_size = max(0, int(prefetch_size))
        self.as_contiguous = as_contiguous
        self.sampler = RandomSampler(dataset) if shuffle else SequentialSampler(dataset)
        self.batch_sampler = BatchSampler(self.sampler, batch_size, drop_last)

    def __iter__(self):
        return _DataLoaderIter(self)

    def __len__(self):
        return len(self.batch_sampler)


class _ArrayPairDataset(Dataset):
    def __init__(self, X, y):
        self.data, self.target = X, y

    def __getitem__(self, index):
        return self.data[index], self.target[index]

    def __len__(self):
        return len(self.data)


def data_loader(X, y, batch_size, shuffle=False, prefetch_size: int = 0, as_contiguous: bool = True):
    return DataLoader(_ArrayPairDataset(X, y), batch_size, shuffle, prefetch_size=prefetch_size,
                      as_contiguous=as_contiguous)
