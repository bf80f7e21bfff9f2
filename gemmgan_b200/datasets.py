"""On-disk dataset readers that yield the batch tuples the WGAN-GP step consumes.

Behavioural mirror of the reference's two loaders (same file layout on disk, same splits under the same seed,
same normalisation, same tuple layouts), written once and shared by the two drop-in modules at the repo root:
  multi_patch_gan_dataloader.py             src/multi_patch_gan_dataloader.py:9-262   (film / concat / img variants)
  multi_patch_multi_token_gan_dataloader.py src/multi_patch_multi_token_gan_dataloader.py:11-187 (paper / cross)

Dataset directory (reference :153-166 / :84-97):
  rna_seq.parquet                       [cases x genes] expression table, index = case id
  case_ids.txt                          one case id per line
  <text_embedding_file>.parquet         [cases x Dt] text embedding table (index = case id)
  <patch_embeddings_folder>/<case>.npy  [n_patches_of_case, Dp] patch embeddings
  <token_embeddings_folder>/<case>.npy  [1, T, Dt] token embeddings + <case>_attention_mask.npy [1, T] (1 = token)
  metainfos.pkl                         {case: {'disease_type': str, 'primary_site': str}}
Batch tuples (True in a mask = padding; the reference never sets a patch mask bit, see MASK_ZERO_PADDING):
  multi-patch       (text[Dt] f32, genes[G] f32, patches[P,Dp] f32, patch_pad[P] bool, disease i64, site i64)
  multi-patch-token (tokens[T,Dt] f32, token_pad[T] bool, genes[G] f32, patches[P,Dp] f32, patch_pad[P] bool,
                     disease i64, site i64)
Everything here is host-side preparation; the tensors are cast / staged on the device by gg_engine_set_batch.
"""
from __future__ import annotations

import pickle
import random
from pathlib import Path

import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset


# ------------------------------------------------------------------------------------- helpers
def split_data_train_test(n_samples, train_rate=0.80, seed=42, shuffle=True):
    """Index split train / test (reference :51-73): seeded permutation, first train_rate of it trains."""
    random.seed(seed)
    np.random.seed(seed)
    order = np.arange(n_samples)
    if shuffle:
        np.random.shuffle(order)
    cut = int(train_rate * n_samples)
    return order[:cut], order[cut:]


def split_data(n_samples, train_rate=0.80, validation_rate=0.20, seed=42, shuffle=True):
    """Index split train / validation / test (reference :77-102): the validation set is carved out of the
    training share, everything after it is the test set."""
    random.seed(seed)
    np.random.seed(seed)
    order = np.arange(n_samples)
    if shuffle:
        np.random.shuffle(order)
    n_train = int(train_rate * (1 - validation_rate) * n_samples)
    n_val = int(train_rate * validation_rate * n_samples)
    return order[:n_train], order[n_train:n_train + n_val], order[n_train + n_val:]


def standardize(x, mean=None, std=None):
    """Per-gene z-score; statistics default to those of x itself (reference :105-111)."""
    mean = np.mean(x, axis=0) if mean is None else mean
    std = np.std(x, axis=0) if std is None else std
    return (x - mean) / std


def min_max(x, max=None, min=None):  # noqa: A002 - the reference's argument names
    """Per-gene min-max scaling (reference :114-120; its `min is None` branch assigns the wrong variable and
    cannot work without an explicit `min` — both bounds default to the data's here)."""
    max = np.max(x, axis=0) if max is None else max  # noqa: A001
    min = np.min(x, axis=0) if min is None else min  # noqa: A001
    return (x - min) / (max - min)


def seed_worker(worker_id):
    """DataLoader worker_init_fn: numpy / random follow torch's per-worker seed (reference :123-126)."""
    s = torch.initial_seed() % 2 ** 32
    np.random.seed(s)
    random.seed(s)


# Reference quirk, kept by default for drop-in parity: both reference datasets build the padding mask AFTER they
# have replaced `patches` by the zero-padded array (multi_patch_gan_dataloader.py:37-39,
# multi_patch_multi_token_gan_dataloader.py:37-39), so `self.num_patches - patches.shape[0]` is 0 and the mask is
# all False: the zero rows are attended like real patches. Set MASK_ZERO_PADDING = True to mark them as padding.
MASK_ZERO_PADDING = False


def _fit_patches(patches: np.ndarray, num_patches: int):
    """Exactly num_patches rows: a random subset when the case has more (np.random.choice without replacement,
    as the reference :33-36), zero rows when it has fewer (:37-38); mask: see MASK_ZERO_PADDING."""
    n = patches.shape[0]
    if n > num_patches:
        keep = np.random.choice(n, num_patches, replace=False)
        return patches[keep], np.zeros(num_patches, dtype=bool)
    pad = np.zeros((num_patches - n, patches.shape[1]), dtype=patches.dtype)
    mask = np.arange(num_patches) >= n if MASK_ZERO_PADDING else np.zeros(num_patches, dtype=bool)
    return np.concatenate((patches, pad), axis=0), mask


# ------------------------------------------------------------------------------------ datasets
class MultiPatchGANDataset(Dataset):
    """One text vector + a fixed number of patch embeddings per case (reference :9-48)."""

    def __init__(self, case_ids, text_embeddings, patches_path, gene_expressions, disease_types, primary_site,
                 num_patches=256):
        self.case_ids = case_ids
        self.text_embeddings = text_embeddings
        self.patches_path = Path(patches_path)
        self.gene_expressions = gene_expressions
        self.disease_types = disease_types
        self.primary_site = primary_site
        self.num_patches = num_patches

    def __len__(self):
        return self.text_embeddings.shape[0]

    def __getitem__(self, idx):
        patches, mask = _fit_patches(np.load(self.patches_path / f"{self.case_ids[idx]}.npy"), self.num_patches)
        return (torch.tensor(self.text_embeddings[idx], dtype=torch.float32),
                torch.tensor(self.gene_expressions[idx], dtype=torch.float32),
                torch.tensor(patches, dtype=torch.float32),
                torch.tensor(mask, dtype=torch.bool),
                torch.tensor(self.disease_types[idx], dtype=torch.long),
                torch.tensor(self.primary_site[idx], dtype=torch.long))


class MultiPatchMultiTokenGANDataset(Dataset):
    """Token-level text embeddings (+ padding mask) + patch embeddings per case (reference :11-55)."""

    def __init__(self, case_ids, tokens_path, patches_path, gene_expressions, disease_types, primary_site,
                 num_patches=256):
        self.case_ids = case_ids
        self.tokens_path = Path(tokens_path)
        self.patches_path = Path(patches_path)
        self.gene_expressions = gene_expressions
        self.disease_types = disease_types
        self.primary_site = primary_site
        self.num_patches = num_patches

    def __len__(self):
        return self.gene_expressions.shape[0]

    def __getitem__(self, idx):
        case = self.case_ids[idx]
        patches, mask = _fit_patches(np.load(self.patches_path / f"{case}.npy"), self.num_patches)
        tokens = torch.tensor(np.load(self.tokens_path / f"{case}.npy"), dtype=torch.float32).squeeze(0)
        attend = torch.tensor(np.load(self.tokens_path / f"{case}_attention_mask.npy"), dtype=torch.bool).squeeze(0)
        return (tokens,
                ~attend,  # the tokenizer marks real tokens with 1; attention masks mark PADDING with True (:46-47)
                torch.tensor(self.gene_expressions[idx], dtype=torch.float32),
                torch.tensor(patches, dtype=torch.float32),
                torch.tensor(mask, dtype=torch.bool),
                torch.tensor(self.disease_types[idx], dtype=torch.long),
                torch.tensor(self.primary_site[idx], dtype=torch.long))


# ------------------------------------------------------------------------------------- loaders
class _Prepared:
    """Everything the two loader functions share: case intersection, gene filter, split, normalisation, labels."""

    def __init__(self, dataset_path, text_embedding_file, patch_embeddings_folder, normalize, percentage_to_remove,
                 norm_type, need_text_table=True):
        dataset_path = Path(dataset_path)
        expr = pd.read_parquet(dataset_path / "rna_seq.parquet")
        listed = {c.strip() for c in (dataset_path / "case_ids.txt").read_text().splitlines()}
        self.text_table = pd.read_parquet(dataset_path / text_embedding_file)
        with_patches = {p.stem for p in (dataset_path / patch_embeddings_folder).glob("*.npy")}
        cases = sorted(listed & with_patches & set(self.text_table.index) & set(expr.index))
        # genes that are zero in more than `percentage_to_remove` % of ALL rows are dropped (:171-172)
        zero_pct = (expr == 0).sum() / len(expr) * 100
        expr = expr.loc[:, zero_pct <= percentage_to_remove]
        self.n_genes = expr.shape[1]
        self.gene_names = expr.columns
        parts = split_data(len(cases))
        self.case_ids = [[cases[i] for i in part] for part in parts]
        frames = [expr.loc[ids] for ids in self.case_ids]
        if normalize and norm_type == "standardize":
            mu, sd = np.mean(frames[0], axis=0), np.std(frames[0], axis=0)
            frames = [standardize(f, mean=mu, std=sd).fillna(0) for f in frames]
        elif normalize and norm_type == "min-max":
            hi, lo = np.max(frames[0], axis=0), np.min(frames[0], axis=0)
            frames = [min_max(f, max=hi, min=lo).fillna(0) for f in frames]
        self.genes = [f.values for f in frames]
        with open(dataset_path / "metainfos.pkl", "rb") as f:
            meta = pickle.load(f)
        self.disease = self._encode(meta, "disease_type")
        self.site = self._encode(meta, "primary_site")

    def _encode(self, meta, key):
        raw = [[meta[c][key] for c in ids] for ids in self.case_ids]
        code = {name: i for i, name in enumerate(sorted({v for part in raw for v in part}))}
        return [[code[v] for v in part] for part in raw]


def _loaders(datasets, batch_size, num_workers, g):
    mk = lambda ds, shuffle: DataLoader(ds, batch_size=batch_size, shuffle=shuffle, worker_init_fn=seed_worker,  # noqa: E731
                                        generator=g, num_workers=num_workers)
    return mk(datasets[0], True), mk(datasets[1], True), mk(datasets[2], False)


def multi_patch_loaders(dataset_path, normalize=True, percentage_to_remove=90, norm_type="standardize",
                        num_patches=256, batch_size=8, seed=42, num_workers=4, embedding_dim=256,
                        text_embedding_file=None, patch_embeddings_folder=None):
    """dataloader_multi_patch_conditional_gan of src/multi_patch_gan_dataloader.py:129-262."""
    text_embedding_file = text_embedding_file or f"text_embeddings_contrastive_{embedding_dim}.parquet"
    patch_embeddings_folder = patch_embeddings_folder or f"patch_embeddings_contrastive_{embedding_dim}"
    g = torch.Generator()
    g.manual_seed(seed)
    torch.manual_seed(seed)
    dataset_path = Path(dataset_path)
    p = _Prepared(dataset_path, text_embedding_file, patch_embeddings_folder, normalize, percentage_to_remove, norm_type)
    np.save("gene_names.npy", p.gene_names)  # the reference drops the kept gene names next to the script (:187)
    ds = [MultiPatchGANDataset(p.case_ids[i], p.text_table.loc[p.case_ids[i]].values,
                               dataset_path / patch_embeddings_folder, p.genes[i], p.disease[i], p.site[i],
                               num_patches=num_patches) for i in range(3)]
    return (*_loaders(ds, batch_size, num_workers, g), p.n_genes)


def multi_patch_multi_token_loaders(dataset_path, normalize=True, percentage_to_remove=90, norm_type="standardize",
                                    num_patches=256, batch_size=8, seed=42, num_workers=4, embedding_dim=256,
                                    text_embedding_file=None, patch_embeddings_folder=None,
                                    token_embeddings_folder=None):
    """dataloader_multi_patch_conditional_gan of src/multi_patch_multi_token_gan_dataloader.py:58-187."""
    text_embedding_file = text_embedding_file or f"text_embeddings_contrastive_{embedding_dim}.parquet"
    patch_embeddings_folder = patch_embeddings_folder or f"patch_embeddings_contrastive_{embedding_dim}"
    token_embeddings_folder = token_embeddings_folder or f"../text_embeddings_contrastive_{embedding_dim}"
    g = torch.Generator()
    g.manual_seed(seed)
    torch.manual_seed(seed)
    dataset_path = Path(dataset_path)
    p = _Prepared(dataset_path, text_embedding_file, patch_embeddings_folder, normalize, percentage_to_remove, norm_type)
    ds = [MultiPatchMultiTokenGANDataset(p.case_ids[i], dataset_path / token_embeddings_folder,
                                         dataset_path / patch_embeddings_folder, p.genes[i], p.disease[i], p.site[i],
                                         num_patches=num_patches) for i in range(3)]
    return (*_loaders(ds, batch_size, num_workers, g), p.n_genes)
