"""Drop-in for the reference's src/benchmark_gan_dataloader.py (same names, signatures and batch tuples):
BenchmarkGANDataset [:10-37], split_data_train_test [:39], standardize [:65], min_max [:74], seed_worker [:83],
dataloader_benchmark_conditional_gan [:89-199] — the loaders of benchmark_generative_model.py.
Implementation: gemmgan_b200/datasets.py."""
from gemmgan_b200.datasets import (BenchmarkGANDataset, min_max, seed_worker, split_data,  # noqa: F401
                                   split_data_train_test, standardize)
from gemmgan_b200.datasets import benchmark_loaders as dataloader_benchmark_conditional_gan  # noqa: F401
