from .data import *  # noqa: F401,F403
from .data import Dataset, DataLoader, data_loader  # noqa: F401
