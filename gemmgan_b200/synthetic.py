"""Synthetic batches in the reference's dataloader tuple layouts (device-resident or pinned host).

Layouts (reference file:line):
  paper   (token_embeddings[T,Dt] f32, token_pad[T] bool, gene_expression[G] f32, patches[P,Dp] f32,
           patch_pad[P] bool, disease_type i64, primary_site i64)
          src/multi_patch_multi_token_gan_dataloader.py:55
  film    (text_embedding[Dt], gene_expression[G], patches[P,Dp], padding_mask[P], disease_type,
           primary_site)   src/multi_patch_gan_dataloader.py:48
  vanilla (gene_expression[G],)   src/data_loader.py:160 (TensorDataset)
  label   (gene_expression[G] f32, disease_type i64, primary_site i64)   src/benchmark_gan_dataloader.py:37
Masks: True = padding; token 0 is never padded. Data ~ N(0,1) (the real genes are z-scored log2(TPM+1)).
"""
from __future__ import annotations

import torch
from torch.utils.data import DataLoader, TensorDataset


def synthetic_tensors(variant, n, n_genes, n_patches=8, n_tokens=1, text_dim=768, patch_dim=1024, seed=42,
                      ragged=False):
    g = torch.Generator().manual_seed(seed)
    genes = torch.randn(n, n_genes, generator=g)
    if variant == "vanilla":
        return (genes,)
    if variant == "label":
        return (genes, torch.randint(0, 10, (n,), generator=g), torch.randint(0, 10, (n,), generator=g))
    patches = torch.randn(n, n_patches, patch_dim, generator=g)
    ppad = torch.zeros(n, n_patches, dtype=torch.bool)
    if ragged and n_patches > 1:
        k = torch.randint(0, n_patches, (n,), generator=g)
        ppad = torch.arange(n_patches)[None, :] >= (n_patches - k)[:, None]
    dtype_ = torch.randint(0, 10, (n,), generator=g)
    psite = torch.randint(0, 10, (n,), generator=g)
    if variant == "film":
        text = torch.randn(n, text_dim, generator=g)
        return (text, genes, patches, ppad, dtype_, psite)
    text = torch.randn(n, n_tokens, text_dim, generator=g)
    tpad = torch.zeros(n, n_tokens, dtype=torch.bool)
    if ragged and n_tokens > 1:
        k = torch.randint(0, n_tokens, (n,), generator=g)
        tpad = torch.arange(n_tokens)[None, :] >= (n_tokens - k)[:, None]
    return (text, tpad, genes, patches, ppad, dtype_, psite)


def synthetic_loader(variant, n_samples, batch_size, n_genes, n_patches=8, n_tokens=1, seed=42, **kw):
    tensors = synthetic_tensors(variant, n_samples, n_genes, n_patches, n_tokens, seed=seed, **kw)
    return DataLoader(TensorDataset(*tensors), batch_size=batch_size, shuffle=False, drop_last=True)
