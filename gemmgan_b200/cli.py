"""`python <drop-in script>.py ...`: the command line of the reference's training scripts on the sm_100a engine.

The reference scripts end in the same block (src/conditional_gan_cross_attention_with_film.py:902-995, and
conditional_gan_film.py:1014-1100, conditional_gan_cross_attention.py:869-960, conditional_gan_img_transformer.py:994-1080,
conditional_gan_concat.py:1023-1110, conditional_gan_attention.py:889-945): parse `--seed --num_epochs --batch_size
--latent_dim --hidden_dim --embedding_dim --num_patches --dataset_path --output_path --num_workers --freq_compute_test
--freq_plot_images [--optimizer] [--condition_type]`, build the three loaders with
`dataloader_multi_patch_conditional_gan`, construct `WGAN_GP`, `fit(train, validation, test, epochs)`, then the utility
evaluators and the DCR / NNDR privacy block over the `test_*` folders fit() left behind. `main()` is that block:

  * same flags and defaults, except `--dataset_path`, whose reference default is a path on the authors' cluster: here
    it defaults to '' = synthetic batches in the loader's tuple layout (gemmgan_b200/synthetic.py; no files needed);
  * the loaders are the drop-in loader modules (gemmgan_b200/datasets.py, equal to the reference's tensor for tensor);
    `--device_loader` uploads each split once and assembles the batches on the GPU (DeviceResidentLoader);
  * the file names the reference hard-codes in the call (`clinical_modernbert_embeddings.parquet`,
    `patch_embeddings_uni`, `../clinical_modernbert_embeddings`) and the feature widths (768 / 1024) are flags with
    those defaults;
  * after fit(): `loss_dict`, the best epochs, and the privacy block on the GPU kernels
    (`trainer.privacy_report`, printed in the reference's format). The sklearn / lightgbm utility evaluators
    (UtilityEvaluator, UtilityEvaluatorPrimary) are host code outside the hot path and are not run.
"""
from __future__ import annotations

import argparse
import importlib
import os
from glob import glob
from pathlib import Path

# script key -> (drop-in module, loader module, batch tuple layout of gemmgan_b200/synthetic.py)
SCRIPTS = {
    "paper": ("conditional_gan_cross_attention_with_film", "multi_patch_multi_token_gan_dataloader", "paper"),
    "cross": ("conditional_gan_cross_attention", "multi_patch_multi_token_gan_dataloader", "paper"),
    "film": ("conditional_gan_film", "multi_patch_gan_dataloader", "film"),
    "img": ("conditional_gan_img_transformer", "multi_patch_gan_dataloader", "film"),
    "attn": ("conditional_gan_attention", "multi_patch_gan_dataloader", "film"),
    "concat": ("conditional_gan_concat", "multi_patch_gan_dataloader", "film"),
    # gene-expression-only and label-conditioned scripts: other flags (vanilla_gan_unconditional.py:765-799,
    # benchmark_generative_model.py:906-962), fit(train, test, epochs)
    "vanilla": ("vanilla_gan_unconditional", "data_loader", "vanilla"),
    "label": ("benchmark_generative_model", "benchmark_gan_dataloader", "label"),
}


def _build_simple_parser() -> argparse.ArgumentParser:
    """Flags of the unconditional / label-conditioned scripts (vanilla_gan_unconditional.py:765-775)."""
    p = argparse.ArgumentParser(description='WGAN-GP training on the sm_100a engine (flags of the unconditional / label scripts)')
    p.add_argument('--dataset_path', type=str, default='', help="path to dataset ('' = synthetic batches)")
    p.add_argument('--output_path', type=str, default='', help='results directory')
    p.add_argument('--batch_size', type=int, default=8, help='rows per batch')
    p.add_argument('--epochs', type=int, default=500, help='training epochs')
    p.add_argument('--latent_dim', type=int, default=256, help='width of z')
    p.add_argument('--num_workers', type=int, default=4, help='DataLoader worker processes')
    p.add_argument('--seed', type=int, default=42, help='seed of torch / numpy / the split')
    # this implementation
    p.add_argument('--optimizer', type=str, default='rms_prop')
    p.add_argument('--hidden_dim', type=int, default=256, help='the reference hard-codes 256')
    p.add_argument('--n_genes', type=int, default=18868, help='synthetic batches: number of genes')
    p.add_argument('--synthetic_batches', type=int, default=4, help='synthetic batches per epoch')
    return p


def build_parser(script: str) -> argparse.ArgumentParser:
    if script in ("vanilla", "label"):
        return _build_simple_parser()
    p = argparse.ArgumentParser(description="WGAN-GP training on the sm_100a engine (flags of the reference's conditional scripts)")
    # the reference's flags (…with_film.py:903-917)
    p.add_argument('--seed', type=int, default=42, help='seed of torch / numpy / the split')
    p.add_argument('--num_epochs', type=int, default=500, help='training epochs')
    p.add_argument('--batch_size', type=int, default=8, help='rows per batch')
    p.add_argument('--latent_dim', type=int, default=256, help='width of z')
    p.add_argument('--hidden_dim', type=int, default=256, help='width of the two trunk layers')
    p.add_argument('--embedding_dim', type=int, default=256, help='width of the conditioning vector')
    p.add_argument('--num_patches', type=int, default=256, help='patch tokens per sample (sub-sampled / zero-padded to this)')
    p.add_argument('--dataset_path', type=str, default='', help="Path to the dataset ('' = synthetic batches)")
    p.add_argument('--output_path', type=str, default='', help='results directory (checkpoints, .npy dumps)')
    p.add_argument('--num_workers', type=int, default=16, help='DataLoader worker processes')
    p.add_argument('--freq_compute_test', type=int, default=50, help='epochs between validation passes / checkpoints')
    p.add_argument('--freq_plot_images', type=int, default=16, help='Frequency for the plot (accepted, unused)')
    p.add_argument('--optimizer', type=str, default='rms_prop', help='rms_prop | adam | adamw')
    if script == "concat":
        p.add_argument('--condition_type', type=str, default='text', choices=['text', 'image'],
                       help='which embedding conditions the nets')
    # what the reference hard-codes in its loader / model calls (:925-949)
    p.add_argument('--text_embedding_file', type=str, default='clinical_modernbert_embeddings.parquet')
    p.add_argument('--patch_embeddings_folder', type=str, default='patch_embeddings_uni')
    p.add_argument('--token_embeddings_folder', type=str, default='../clinical_modernbert_embeddings')
    p.add_argument('--text_embedding_dims', type=int, default=768)
    p.add_argument('--patches_embedding_dims', type=int, default=1024)
    # this implementation
    p.add_argument('--device_loader', action='store_true',
                   help='upload each split once and assemble the batches on the GPU (DeviceResidentLoader)')
    p.add_argument('--n_genes', type=int, default=18868, help='synthetic batches: number of genes')
    p.add_argument('--num_text_tokens', type=int, default=1, help='synthetic batches: text tokens per sample')
    p.add_argument('--synthetic_batches', type=int, default=4, help='synthetic batches per epoch')
    return p


def _device_resident(loader):
    from torch.utils.data import RandomSampler

    from .datasets import DeviceResidentLoader

    return DeviceResidentLoader(loader.dataset, loader.batch_size, shuffle=isinstance(loader.sampler, RandomSampler),
                                generator=loader.generator)


def load_data(script: str, args):
    """(train, validation, test, n_genes): the reference's loader call (:925-937), or synthetic batches."""
    _, loader_module, layout = SCRIPTS[script]
    if not args.dataset_path:
        import torch

        from .synthetic import synthetic_loader

        torch.manual_seed(args.seed)
        train = synthetic_loader(layout, n_samples=args.batch_size * args.synthetic_batches, batch_size=args.batch_size,
                                 n_genes=args.n_genes, n_patches=args.num_patches, n_tokens=args.num_text_tokens,
                                 seed=args.seed, text_dim=args.text_embedding_dims, patch_dim=args.patches_embedding_dims)
        return train, None, None, args.n_genes
    lm = importlib.import_module(loader_module)
    kw = dict(normalize=True, percentage_to_remove=90, norm_type='standardize', num_patches=args.num_patches,
              batch_size=args.batch_size, seed=args.seed, num_workers=args.num_workers,
              text_embedding_file=args.text_embedding_file, patch_embeddings_folder=args.patch_embeddings_folder)
    if script != "concat":
        kw["embedding_dim"] = args.embedding_dim
    if layout == "paper":
        kw["token_embeddings_folder"] = args.token_embeddings_folder
    train, val, test, n_genes = lm.dataloader_multi_patch_conditional_gan(Path(args.dataset_path), **kw)
    if args.device_loader:
        train, val, test = _device_resident(train), _device_resident(val), _device_resident(test)
    return train, val, test, n_genes


def build_model(script: str, args, n_genes: int):
    """The reference's constructor call (:939-951)."""
    m = importlib.import_module(SCRIPTS[script][0])
    h = args.hidden_dim
    kw = dict(input_dims=n_genes, latent_dims=args.latent_dim, embedding_dims=args.embedding_dim,
              generator_dims=[h, h, n_genes], discriminator_dims=[h, h, 1], optimizer=args.optimizer,
              negative_slope=0.0, is_bn=False, lr_d=5e-4, lr_g=5e-4, gp_weight=10, p_aug=0, norm_scale=0.5,
              freq_compute_test=args.freq_compute_test, results_dire=args.output_path)
    if script == "concat":
        image = args.condition_type == 'image'
        kw.update(condition_on=args.condition_type,
                  input_embedding_dims=args.patches_embedding_dims if image else args.text_embedding_dims)
    else:
        kw.update(text_embedding_dims=args.text_embedding_dims, patches_embedding_dims=args.patches_embedding_dims)
    return m.WGAN_GP(**kw)


def _label_vocabularies(dataset_path):
    """Numbers of distinct disease types / primary sites among the listed cases (benchmark_generative_model.py:931-945)."""
    import pickle

    with open(os.path.join(dataset_path, 'metainfos.pkl'), 'rb') as f:
        meta = pickle.load(f)
    with open(os.path.join(dataset_path, 'case_ids.txt')) as f:
        cases = {c.strip() for c in f.read().splitlines()}
    return [len({m[key] for c, m in meta.items() if c in cases}) for key in ('disease_type', 'primary_site')]


def _main_simple(script: str, args):
    """vanilla_gan_unconditional.py:777-799 / benchmark_generative_model.py:917-962."""
    module, loader_module, layout = SCRIPTS[script]
    m = importlib.import_module(module)
    vocab = [10, 10]
    if args.dataset_path:
        lm = importlib.import_module(loader_module)
        fn = lm.dataloader_tcga if script == "vanilla" else lm.dataloader_benchmark_conditional_gan
        train, _, test, n_genes = fn(dataset_path=Path(args.dataset_path), batch_size=args.batch_size,
                                     num_workers=args.num_workers, seed=args.seed)
        if script == "label":
            vocab = _label_vocabularies(args.dataset_path)
    else:
        import torch

        from .synthetic import synthetic_loader

        torch.manual_seed(args.seed)
        train, test, n_genes = synthetic_loader(layout, n_samples=args.batch_size * args.synthetic_batches,
                                                batch_size=args.batch_size, n_genes=args.n_genes, seed=args.seed), None, args.n_genes
    h = args.hidden_dim
    kw = dict(input_dims=n_genes, latent_dims=args.latent_dim, generator_dims=[h, h, n_genes],
              discriminator_dims=[h, h, 1], negative_slope=0.0, is_bn=False, lr_d=5e-4, lr_g=5e-4, gp_weight=10, p_aug=0,
              norm_scale=0.5, optimizer=args.optimizer, results_dire=args.output_path)
    model = m.WGAN_GP_nocond(vocab_sizes=[], **kw) if script == "vanilla" else m.WGAN_GP_benchmark(vocab_sizes=vocab, **kw)
    model.fit(train, test, epochs=args.epochs)
    print(model.loss_dict)
    return model


def main(script: str, argv=None):
    args = build_parser(script).parse_args(argv)
    print(f'Arguments: {args.__dict__}')
    if script in ("vanilla", "label"):
        return _main_simple(script, args)
    train, val, test, n_genes = load_data(script, args)
    model = build_model(script, args, n_genes)
    if script == "attn":    # fit(train_data, test_data, epochs, val) in this script (conditional_gan_attention.py:523)
        model.fit(train, test, epochs=args.num_epochs)
    else:
        model.fit(train, val, test, epochs=args.num_epochs)
    print(model.loss_dict)
    for name, scores in (("correlation", model.corr_scores), ("precision", model.precision_scores),
                         ("recall", model.recall_scores)):
        if scores:
            model.print_best_epoch(scores, name=name)
    if args.output_path and glob(os.path.join(args.output_path, 'test_*')) and model._is_main_rank():
        from .trainer import privacy_report

        print()
        print("--------- Privacy Evaluation ----------")
        r = privacy_report(args.output_path)
        print(f"DCR {r['mean_dcr']:.4f}±{r['std_dcr']:.4f}, NNDR {r['mean_nndr']:.4f}±{r['std_nndr']:.4f}")
        model.privacy = r
    return model
