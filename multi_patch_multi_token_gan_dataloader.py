"""Drop-in for the reference's src/multi_patch_multi_token_gan_dataloader.py (same names, signatures and batch
tuples): MultiPatchMultiTokenGANDataset [:11-55], dataloader_multi_patch_conditional_gan [:58-187] (and the helpers
it imports from multi_patch_gan_dataloader [:8]). Implementation: gemmgan_b200/datasets.py."""
from gemmgan_b200.datasets import (MultiPatchMultiTokenGANDataset, min_max, seed_worker, split_data,  # noqa: F401
                                   standardize)
from gemmgan_b200.datasets import multi_patch_multi_token_loaders as dataloader_multi_patch_conditional_gan  # noqa: F401
