"""On-disk dataset readers that yield the batch tuples the WGAN-GP step consumes.

Behavioural mirror of the reference's two loaders (same file layout on disk, same splits under the same seed,
same normalisation, same tuple layouts), written once and shared by the drop-in modules at the repo root:
  multi_patch_gan_dataloader.py             src/multi_patch_gan_dataloader.py:9-262   (film / concat / img variants)
  multi_patch_multi_token_gan_dataloader.py src/multi_patch_multi_token_gan_dataloader.py:11-187 (paper / cross)
  data_loader.py                            src/data_loader.py:11-174 (dataloader_tcga: the unconditional script)
  benchmark_gan_dataloader.py               src/benchmark_gan_dataloader.py:10-199 (label-conditioned baseline)

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
        self.meta = meta
        self.disease = self._encode(meta, "disease_type")
        self.site = self._encode(meta, "primary_site")

    def _encode(self, meta, key, parts=(0, 1, 2)):
        """Label codes = rank of the name among the sorted names seen in the splits `parts`."""
        raw = [[meta[c][key] for c in ids] for ids in self.case_ids]
        code = {name: i for i, name in enumerate(sorted({v for k in parts for v in raw[k]}))}
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


# --------------------------------------------------- gene-only and label-conditioned loaders (vanilla / benchmark scripts)
class BenchmarkGANDataset(Dataset):
    """(gene_expression[G] f32, disease_type i64, primary_site i64) per case: src/benchmark_gan_dataloader.py:10-37."""

    def __init__(self, case_ids, gene_expressions, disease_types, primary_sites):
        self.case_ids = case_ids
        self.gene_expressions = gene_expressions
        self.disease_types = disease_types
        self.primary_sites = primary_sites

    def __len__(self):
        return self.gene_expressions.shape[0]

    def __getitem__(self, idx):
        return (torch.tensor(self.gene_expressions[idx], dtype=torch.float32),
                torch.tensor(self.disease_types[idx], dtype=torch.long),
                torch.tensor(self.primary_sites[idx], dtype=torch.long))


# both scripts intersect the cases with these two tables "to reproduce the same split as in the contrastive training
# and conditional gan training" (src/data_loader.py:109-116, src/benchmark_gan_dataloader.py:112-119)
_SPLIT_TEXT_TABLE = "text_embeddings_contrastive_256.parquet"
_SPLIT_PATCH_FOLDER = "patch_embeddings_contrastive_256"


def tcga_loaders(dataset_path, normalize=True, percentage_to_remove=90, norm_type="standardize", batch_size=8, seed=42,
                 num_workers=4):
    """dataloader_tcga of src/data_loader.py:87-174 (vanilla_gan_unconditional.py:778): TensorDatasets of the
    normalised expression tables, float64 as pandas hands them out (the trainer casts, vanilla :424); training and
    validation loaders shuffle, the test loader does not."""
    g = torch.Generator()
    g.manual_seed(seed)
    torch.manual_seed(seed)
    p = _Prepared(Path(dataset_path), _SPLIT_TEXT_TABLE, _SPLIT_PATCH_FOLDER, normalize, percentage_to_remove, norm_type)
    from torch.utils.data import TensorDataset

    mk = lambda x, shuffle: DataLoader(TensorDataset(torch.from_numpy(np.array(x))), batch_size=batch_size,  # noqa: E731
                                       shuffle=shuffle, worker_init_fn=seed_worker, generator=g, num_workers=num_workers)
    return mk(p.genes[0], True), mk(p.genes[1], True), mk(p.genes[2], False), p.n_genes


def benchmark_loaders(dataset_path, normalize=True, percentage_to_remove=90, norm_type="standardize", num_patches=256,
                      batch_size=8, seed=42, num_workers=4):
    """dataloader_benchmark_conditional_gan of src/benchmark_gan_dataloader.py:89-199. Disease-type codes rank the
    names of the training and test cases only (:160; a type seen only in the validation split raises KeyError there
    as well), primary-site codes those of all three splits (:171); only the training loader shuffles (:190-195)."""
    g = torch.Generator()
    g.manual_seed(seed)
    torch.manual_seed(seed)
    p = _Prepared(Path(dataset_path), _SPLIT_TEXT_TABLE, _SPLIT_PATCH_FOLDER, normalize, percentage_to_remove, norm_type)
    disease = p._encode(p.meta, "disease_type", parts=(0, 2))
    ds = [BenchmarkGANDataset(p.case_ids[i], p.genes[i], disease[i], p.site[i]) for i in range(3)]
    mk = lambda d, shuffle: DataLoader(d, batch_size=batch_size, shuffle=shuffle, worker_init_fn=seed_worker,  # noqa: E731
                                       generator=g, num_workers=num_workers)
    return mk(ds[0], True), mk(ds[1], False), mk(ds[2], False), p.n_genes


# ------------------------------------------------------------------- device-resident datasets
def _default_device():
    return torch.device("cuda", torch.cuda.current_device())


def _current_stream():
    import ctypes as C

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceResidentLoader:
    """The batches of MultiPatchMultiTokenGANDataset / MultiPatchGANDataset + DataLoader, assembled ON THE GPU.

    The reference's loader np.load()s two or three files per sample in worker processes, collates and copies ~110 KB
    per sample to the device every step (src/multi_patch_multi_token_gan_dataloader.py:25-55, :178-185): at a few
    milliseconds per training step that loader, not the step, bounds fit(). Here every case's patch embeddings (one
    ragged [sum n_i, 1024] matrix), token embeddings / text vectors, masks, gene profiles and labels are uploaded
    ONCE; per batch the host only draws the indices — the same np.random.choice(n, num_patches, replace=False) per
    over-long case, in batch order, so that a seeded run picks the same patches as the reference dataset read with
    num_workers=0 — and three gg_gather_rows launches build the fp32 batch tuple in HBM (padding rows are zero rows,
    mask True = padding, as the reference :37-38, :46-47).

    Iterating yields the reference's tuples with CUDA tensors: multi-token datasets
    (tokens [B,T,768], token_pad [B,T] bool, genes [B,G], patches [B,P,1024], patch_pad [B,P] bool, disease, site),
    single-vector datasets (text [B,768], genes, patches, patch_pad, disease, site). No drop_last, like the reference.
    """

    def __init__(self, dataset, batch_size, shuffle=False, generator=None, device=None):
        import ctypes as C

        from . import _lib

        self._C, self._lib = C, _lib.lib()
        self.device = device or _default_device()
        _lib.require_device(self.device.index or 0)
        self.batch_size, self.shuffle, self.generator = batch_size, shuffle, generator
        self.num_patches = dataset.num_patches
        self.multi_token = isinstance(dataset, MultiPatchMultiTokenGANDataset)
        n = len(dataset)
        self.n = n
        counts, chunks = [], []
        for i in range(n):
            pch = np.load(dataset.patches_path / f"{dataset.case_ids[i]}.npy").astype(np.float32, copy=False)
            counts.append(pch.shape[0])
            chunks.append(pch)
        self.counts = np.asarray(counts, dtype=np.int64)
        self.offsets = np.concatenate(([0], np.cumsum(self.counts)))
        dev = self.device
        self.patch_store = torch.from_numpy(np.ascontiguousarray(np.concatenate(chunks, axis=0))).to(dev)   # [sum n_i, Dp]
        self.genes = torch.from_numpy(np.ascontiguousarray(dataset.gene_expressions, dtype=np.float32)).to(dev)  # (C order: pandas hands out F-ordered blocks)
        self.disease = torch.as_tensor(np.asarray(dataset.disease_types), dtype=torch.long).to(dev)
        self.site = torch.as_tensor(np.asarray(dataset.primary_site), dtype=torch.long).to(dev)
        if self.multi_token:
            toks, pads = [], []
            for i in range(n):
                case = dataset.case_ids[i]
                tk = np.load(dataset.tokens_path / f"{case}.npy").astype(np.float32, copy=False)
                toks.append(tk.reshape(-1, tk.shape[-1]))                                  # [1, T, Dt] -> [T, Dt]
                pads.append(~np.load(dataset.tokens_path / f"{case}_attention_mask.npy").astype(bool).reshape(-1))
            self.text = torch.from_numpy(np.stack(toks)).to(dev)                          # [n, T, Dt]
            self.text_pad = torch.from_numpy(np.stack(pads)).to(dev)                      # [n, T] True = padding
        else:
            self.text = torch.from_numpy(np.ascontiguousarray(dataset.text_embeddings, dtype=np.float32)).to(dev)
            self.text_pad = None

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def _gather(self, src2d, index):
        rows, cols = index.numel(), src2d.shape[1]
        assert src2d.stride(1) == 1 and src2d.stride(0) >= cols, "device stores are row-major"
        out = torch.empty(rows, cols, device=self.device, dtype=torch.float32)
        C = self._C
        from . import _lib
        _lib.check(self._lib.gg_gather_rows(C.c_void_p(src2d.data_ptr()), src2d.stride(0), C.c_void_p(index.data_ptr()),
                                            C.c_void_p(out.data_ptr()), cols, rows, cols, _current_stream()))
        return out

    def __iter__(self):
        order = torch.randperm(self.n, generator=self.generator).numpy() if self.shuffle else np.arange(self.n)
        P = self.num_patches
        for b0 in range(0, self.n, self.batch_size):
            ids = order[b0:b0 + self.batch_size]
            B = len(ids)
            rows = np.full((B, P), -1, dtype=np.int64)
            pad = np.zeros((B, P), dtype=bool)
            for j, i in enumerate(ids):       # the reference's per-sample logic (:32-40), indices only
                cnt, off = int(self.counts[i]), int(self.offsets[i])
                if cnt > P:
                    rows[j] = off + np.random.choice(cnt, P, replace=False)
                else:
                    rows[j, :cnt] = off + np.arange(cnt)
                    pad[j, cnt:] = MASK_ZERO_PADDING
            dev = self.device
            sample = torch.from_numpy(np.ascontiguousarray(ids)).to(dev)
            patches = self._gather(self.patch_store, torch.from_numpy(rows.reshape(-1)).to(dev)).view(B, P, -1)
            genes = self._gather(self.genes, sample)
            ppad = torch.from_numpy(pad).to(dev)
            if self.multi_token:
                T, Dt = self.text.shape[1], self.text.shape[2]
                text = self._gather(self.text.view(self.n, T * Dt), sample).view(B, T, Dt)
                yield (text, self.text_pad[sample], genes, patches, ppad, self.disease[sample], self.site[sample])
            else:
                text = self._gather(self.text, sample)
                yield (text, genes, patches, ppad, self.disease[sample], self.site[sample])
