class Dataset:
    """Map-style dataset protocol (reference: DeepFlows/utils/data/dataset.py)."""

    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError
