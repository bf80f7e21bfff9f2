"""Drop-in for the reference's src/benchmark_generative_model.py (the label-conditioned WGAN-GP baseline: disease
type and primary site enter through two nn.Embedding tables whose rows are concatenated to the trunk input) backed
by the sm_100a engine.

Same public names and signatures as the reference (file:line of the reference in brackets):
  categorical_embedding [:27-35], build_linear_block / build_discriminator / build_generator [:41-99],
  wasserstein_loss, G_loss, D_loss [:75-89], discriminator [:101-160], generator [:163-236],
  WGAN_GP_model_benchmark [:238-259], WGAN_GP_benchmark [:265-...] with init_train [:337], build_WGAN_GP [:352],
  gradient_penalty [:366], train_disc [:393], train_gen [:448], train [:486], generate_samples(_all) [:499-543], fit.
Model argument order: (x, categorical_covariates, categorical_covariates_2) [:138, :204]; batch tuple of
benchmark_gan_dataloader.py:37: (gene_expression, disease_type, primary_site). No gradient clipping, no dropout;
optimizers rms_prop / adam only [:339-347]. Evaluation / plotting inside the reference's fit() is out of scope.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from gemmgan_b200.models import (LabelDiscriminator, LabelGenerator, build_linear_block, build_stack,  # noqa: F401
                                 categorical_embedding)
from gemmgan_b200.trainer import D_loss, G_loss, TrainerBase, save_numpy, wasserstein_loss  # noqa: F401


def save_numpy(file, data):
    with open(file, 'wb') as f:
        np.save(f, data)


def build_generator(input_dims, generator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, generator_dims, negative_slope, is_bn)


def build_discriminator(input_dims, dicriminator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, dicriminator_dims, negative_slope, is_bn)


class discriminator(LabelDiscriminator):
    pass


class generator(LabelGenerator):
    pass


def WGAN_GP_model_benchmark(latent_dims, vector_dims, numerical_dims, vocab_sizes, generator_dims, discriminator_dims,
                            negative_slope=0.0, is_bn=False):
    gen = generator(latent_dims, numerical_dims, vocab_sizes, generator_dims, negative_slope, is_bn)
    disc = discriminator(vector_dims, numerical_dims, vocab_sizes, discriminator_dims, negative_slope, is_bn)
    return gen, disc


class WGAN_GP_benchmark(TrainerBase):
    variant = "label"

    def __init__(self, input_dims, latent_dims, vocab_sizes, generator_dims, discriminator_dims,
                 negative_slope=0.0, is_bn=False, numerical_dims=[], lr_d=5e-4, lr_g=5e-4, optimizer='rms_prop',
                 gp_weight=10, p_aug=0, norm_scale=0.5, train=True, n_critic=5, freq_print=2, freq_compute_test=10,
                 freq_visualize_test=100, patience=10, normalization='standardize', log2=False, rpm=False,
                 results_dire=''):
        if optimizer.lower() not in ('rms_prop', 'adam'):
            raise ValueError(f"unknown optimizer {optimizer!r} (the reference's init_train knows rms_prop and adam)")
        self.numerical_dims = numerical_dims
        self.vocab_sizes = list(vocab_sizes)
        self._init_common(input_dims, latent_dims, generator_dims, discriminator_dims, negative_slope, is_bn,
                          lr_d, lr_g, optimizer, gp_weight, p_aug, norm_scale, train, n_critic, freq_print,
                          freq_compute_test, freq_visualize_test, patience, normalization, log2, rpm,
                          results_dire)
        self.dropout_p = 0.0     # no dropout layer anywhere in this model

    def _shape_cfg(self):
        # label engines: E = width of the conditioning vector, Dt / Dp = the two vocabulary sizes (include/gemmgan.h)
        return dict(E=self.gen.categorical_embedded_dims, H=self.generator_dims[0], Dt=int(self.vocab_sizes[0]),
                    Dp=int(self.vocab_sizes[1]), P=1, T=1)

    def build_WGAN_GP(self):
        self.numerical_dims = []
        gen, disc = WGAN_GP_model_benchmark(self.latent_dims, self.input_dims, self.numerical_dims, self.vocab_sizes,
                                            self.generator_dims, self.discriminator_dims, self.negative_slope,
                                            self.is_bn)
        self._attach(gen, disc)

    def _check_labels(self, y, vocab):
        """nn.Embedding raises IndexError on an out-of-range index; host tensors are checked before the copy
        (device tensors are trusted: checking them would cost a synchronisation per step)."""
        if not y.is_cuda and y.numel() and (int(y.min()) < 0 or int(y.max()) >= vocab):
            raise IndexError("index out of range in self")

    def _stage(self, genes, cat_vars, cat_vars2):
        self._check_labels(cat_vars, self.vocab_sizes[0])
        self._check_labels(cat_vars2, self.vocab_sizes[1])
        eng = self._engine(cat_vars.shape[0])
        if genes is not None:
            eng.set_batch(genes=self._dev(genes))
        eng.set_labels(self._dev(cat_vars), self._dev(cat_vars2))
        return eng

    # ---- reference-signature entry points -------------------------------------------------
    def gradient_penalty(self, real_data, fake_data, cat_vars, cat_vars2, alpha=None):
        eng = self._stage(None, cat_vars, cat_vars2)
        if alpha is None:
            alpha = self._alpha(eng.B)
        return eng.gradient_penalty(real_data.to(self.device), fake_data.to(self.device), alpha,
                                    training=self.disc.training)

    def train_disc(self, x, z, cat_vars, cat_vars2, alpha=None):
        eng = self._stage(x, cat_vars, cat_vars2)
        self._train_disc_staged(eng, z.to(self.device), alpha)

    def train_gen(self, z, cat_vars, cat_vars_2):
        eng = self._stage(None, cat_vars, cat_vars_2)
        self._train_gen_staged(eng, z.to(self.device))

    def train(self, x_GE, x_cat, x_cat_2, zs=None, alphas=None, prefetch=None):
        eng = self._stage(x_GE, x_cat, x_cat_2)
        self._train_staged(eng, zs, alphas)
        if prefetch is not None:   # host tensors of the NEXT batch: their H2D copies overlap this step
            self.prefetch(*prefetch)

    def _module_forward(self, module, x, cat_vars, cat_vars2):
        eng = self._stage(None, cat_vars, cat_vars2)
        if module is self.gen:
            return eng.generate(x.to(self.device), training=module.training)
        return eng.critic(x.to(self.device), training=module.training)

    def generate_samples(self, x_GE, x_cat, x_cat_2):
        with torch.no_grad():
            self.gen.eval()
            x_real = x_GE.clone().to(torch.float32)
            z = torch.normal(0, 1, size=(x_cat.shape[0], self.latent_dims), device=self.device)
            x_gen = self.gen(z, x_cat, x_cat_2)
        return x_real, x_gen

    def generate_samples_all(self, data):
        """Returns (real, generated, labels, labels, primary sites, primary sites) like the reference (:499-528): the
        generated samples carry the labels they were conditioned on, so both label lists are the loader's."""
        real, gen, cats, sites = [], [], [], []
        for batch in data:
            x_real, x_gen = self.generate_samples(batch[0].to(self.device), batch[1], batch[2])
            real.append(x_real.cpu().numpy())
            gen.append(x_gen.cpu().numpy())
            cats.extend(x.cpu().numpy() for x in batch[1])
            sites.extend(x.cpu().numpy() for x in batch[2])
        return np.vstack(real), np.vstack(gen), cats, cats, sites, sites

    def fit(self, train_data, test_data=None, epochs=1, val=True):
        """Training loop of the reference fit() [:559-...] without its evaluation / plotting."""
        self.build_WGAN_GP()
        if self.isTrain:
            self.init_train()
        for epoch in range(epochs):
            self._epoch_lr_decay(epoch, 50)   # both learning rates halve every 50 epochs (:593-602)
            self.epoch = epoch
            d_sum, g_sum, n = 0.0, 0.0, 0
            for i, (data, nxt) in enumerate(self._lookahead(train_data)):
                self.train(data[0], data[1], data[2], prefetch=None if nxt is None else (nxt[0], nxt[1], nxt[2]))
                d_sum, g_sum, n = d_sum + self.d_batch_loss, g_sum + self.g_batch_loss, n + 1
                if (i + 1) % self.freq_print == 0:
                    print('[Epoch %d/%d] [Batch %d/%d] [D loss : %f] [G loss : %f]' %
                          (epoch + 1, epochs, i + 1, len(train_data), self.disc_loss.item(), self.gen_loss.item()))
            d_mean = d_sum / max(n, 1)
            self.loss_dict['d loss'].append(d_mean[0])
            self.loss_dict['d real loss'].append(d_mean[1])
            self.loss_dict['d fake loss'].append(d_mean[2])
            self.loss_dict["g loss"].append(np.atleast_1d(g_sum)[0])   # summed, not averaged, in the reference (:638)
            if self.result_dire and val and (epoch + 1) % self.freq_compute_test == 0 and epoch + 1 == epochs:
                self._save_checkpoints('last_epoch')


def parse_args(argv=None):
    """The reference's flags [:906-915]; see gemmgan_b200/cli.py."""
    from gemmgan_b200.cli import build_parser

    return build_parser('label').parse_args(argv)


if __name__ == '__main__':
    from gemmgan_b200.cli import main

    main('label')
