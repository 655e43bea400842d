"""Map-style dataset protocol (reference: DeepFlows/utils/data/dataset.py): subclasses provide indexing and a length;
`DataLoader` needs nothing else."""


class Dataset:
    def __len__(self):
        raise NotImplementedError("%s must define __len__" % type(self).__name__)

    def __getitem__(self, index):
        raise NotImplementedError("%s must define __getitem__" % type(self).__name__)
