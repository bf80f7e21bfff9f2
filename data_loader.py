"""Drop-in for the reference's src/data_loader.py (same names and signatures): split_data [:11], split_data_train_test
[:38], standardize [:64], min_max [:73], seed_worker [:82], dataloader_tcga [:87-174] — the gene-expression-only
loaders of vanilla_gan_unconditional.py. dataloader_tcga_cond [:177-263] serves no script of the reference and is not
provided. Implementation: gemmgan_b200/datasets.py."""
from gemmgan_b200.datasets import (min_max, seed_worker, split_data, split_data_train_test,  # noqa: F401
                                   standardize)
from gemmgan_b200.datasets import tcga_loaders as dataloader_tcga  # noqa: F401
