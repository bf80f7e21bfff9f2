"""Drop-in for the reference's src/multi_patch_gan_dataloader.py (same names, signatures and batch tuples):
MultiPatchGANDataset [:9-48], split_data_train_test [:51], split_data [:77], standardize [:105], min_max [:114],
seed_worker [:123], dataloader_multi_patch_conditional_gan [:129-262]. Implementation: gemmgan_b200/datasets.py."""
from gemmgan_b200.datasets import (MultiPatchGANDataset, min_max, seed_worker, split_data,  # noqa: F401
                                   split_data_train_test, standardize)
from gemmgan_b200.datasets import multi_patch_loaders as dataloader_multi_patch_conditional_gan  # noqa: F401
