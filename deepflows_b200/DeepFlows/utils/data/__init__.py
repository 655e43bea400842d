from .dataset import Dataset  # noqa: F401
from .dataloader import DataLoader, data_loader, Sampler, SequentialSampler, RandomSampler, BatchSampler  # noqa: F401
from .prefetcher import DevicePrefetcher  # noqa: F401
from .augment import BatchAugment, smooth_one_hot  # noqa: F401
