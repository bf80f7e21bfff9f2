"""Drop-in for the reference's src/vanilla_gan_unconditional.py (unconditional WGAN-GP on gene
expression profiles) backed by the sm_100a engine.

Same public names and signatures as the reference (file:line of the reference in brackets):
  wasserstein_loss, G_loss, D_loss [:33-47], build_linear_block / build_generator /
  build_discriminator [:50-90], discriminator_nocond [:93-132], generator_nocond [:135-184],
  WGAN_GP_model_nocond [:186-207], WGAN_GP_nocond [:211-431] with build_WGAN_GP_nocond,
  init_train, gradient_penalty, train_disc, train_gen, train, generate_samples, fit.
Evaluation / plotting done by the reference inside fit() (detection, PRDC, UMAP ...) is outside the
hot path and not reproduced here.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from gemmgan_b200.models import VanillaDiscriminator, VanillaGenerator, build_linear_block, build_stack
from gemmgan_b200.trainer import D_loss, G_loss, TrainerBase, save_numpy, wasserstein_loss  # noqa: F401


def build_generator(input_dims, generator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, generator_dims, negative_slope, is_bn)


def build_discriminator(input_dims, dicriminator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, dicriminator_dims, negative_slope, is_bn)


class discriminator_nocond(VanillaDiscriminator):
    pass


class generator_nocond(VanillaGenerator):
    pass


def WGAN_GP_model_nocond(latent_dims, vector_dims, numerical_dims, vocab_sizes, generator_dims,
                         discriminator_dims, negative_slope=0.0, is_bn=False):
    gen = generator_nocond(latent_dims, numerical_dims, vocab_sizes, generator_dims, negative_slope, is_bn)
    disc = discriminator_nocond(vector_dims, numerical_dims, vocab_sizes, discriminator_dims, negative_slope, is_bn)
    return gen, disc


def categorical_embedding(vocab_sizes):
    """One nn.Embedding(vs, int(sqrt(vs)) + 1) per categorical variable [:19-27]; unused by the unconditional nets
    (vocab_sizes=[] in the script), kept for the module's public surface."""
    embedder = torch.nn.ModuleList()
    for vs in vocab_sizes:
        embedder.append(torch.nn.Embedding(vs, int(vs ** 0.5) + 1))
    return embedder


class WGAN_GP_nocond(TrainerBase):
    variant = "vanilla"

    def __init__(self, input_dims, latent_dims, vocab_sizes, generator_dims, discriminator_dims,
                 negative_slope=0.0, is_bn=False, numerical_dims=[], lr_d=5e-4, lr_g=5e-4,
                 optimizer='rms_prop', gp_weight=10, p_aug=0, norm_scale=0.5, train=True, n_critic=5,
                 freq_print=2, freq_compute_test=10, freq_visualize_test=100, patience=10,
                 normalization='standardize', log2=False, rpm=False, results_dire=''):
        self.numerical_dims = numerical_dims
        self.vocab_sizes = vocab_sizes
        self._init_common(input_dims, latent_dims, generator_dims, discriminator_dims, negative_slope, is_bn,
                          lr_d, lr_g, optimizer, gp_weight, p_aug, norm_scale, train, n_critic, freq_print,
                          freq_compute_test, freq_visualize_test, patience, normalization, log2, rpm,
                          results_dire)
        self.dropout_p = 0.0

    def _shape_cfg(self):
        return dict(E=0, H=self.generator_dims[0], Dt=0, Dp=0, P=0, T=0)

    def build_WGAN_GP_nocond(self):
        self.numerical_dims = []
        gen, disc = WGAN_GP_model_nocond(self.latent_dims, self.input_dims, self.numerical_dims, self.vocab_sizes,
                                         self.generator_dims, self.discriminator_dims, self.negative_slope,
                                         self.is_bn)
        self._attach(gen, disc)

    # ---- reference-signature entry points -------------------------------------------------
    def gradient_penalty(self, real_data, fake_data, alpha=None):
        eng = self._engine(real_data.shape[0])
        if alpha is None:
            alpha = self._alpha(eng.B)
        return eng.gradient_penalty(real_data.to(self.device), fake_data.to(self.device), alpha, training=True)

    def train_disc(self, x, z, alpha=None):
        eng = self._engine(z.shape[0])
        eng.set_batch(genes=x.to(self.device))
        self._train_disc_staged(eng, z.to(self.device), alpha)

    def train_gen(self, z):
        eng = self._engine(z.shape[0])
        self._train_gen_staged(eng, z.to(self.device))

    def train(self, x_GE, zs=None, alphas=None, prefetch=None):
        x_real = self._dev(x_GE)
        eng = self._engine(x_real.shape[0])
        eng.set_batch(genes=x_real)
        self._train_staged(eng, zs, alphas)
        if prefetch is not None:   # host tensors of the NEXT batch: their H2D copies overlap this step
            self.prefetch(*prefetch)

    def _module_forward(self, module, x):
        eng = self._engine(x.shape[0])
        if module is self.gen:
            return eng.generate(x.to(self.device), training=module.training)
        return eng.critic(x.to(self.device), training=module.training)

    def generate_samples(self, x_GE):
        with torch.no_grad():
            self.gen.eval()
            x_real = x_GE.clone().to(torch.float32)
            z = torch.normal(0, 1, size=(x_real.shape[0], self.latent_dims), device=self.device)
            x_gen = self.gen(z)
        return x_real, x_gen

    def generate_samples_all(self, data):
        all_real, all_gen = [], []
        for batch in data:
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            x_real, x_gen = self.generate_samples(x.to(self.device))
            all_real.append(x_real.cpu().numpy())
            all_gen.append(x_gen.cpu().numpy())
        return np.vstack(all_real), np.vstack(all_gen)

    def fit(self, train_data, test_data=None, epochs=1, val=True):
        """Training loop of the reference fit(train_data, test_data, epochs, val) [:543-700] without its evaluation /
        plotting (the test loader is only read by that part)."""
        self.build_WGAN_GP_nocond()
        if self.isTrain:
            self.init_train()
        for epoch in range(epochs):
            self._epoch_lr_decay(epoch, 50)  # vanilla halves both LRs every 50 epochs (:558)
            self.epoch = epoch
            d_sum, g_sum, n = 0.0, 0.0, 0
            first = lambda d: d[0] if isinstance(d, (list, tuple)) else d  # noqa: E731
            for i, (data, nxt) in enumerate(self._lookahead(train_data)):
                self.train(first(data), prefetch=None if nxt is None else (first(nxt),))
                d_sum, g_sum, n = d_sum + self.d_batch_loss, g_sum + self.g_batch_loss, n + 1
                if (i + 1) % self.freq_print == 0:
                    print('[Epoch %d/%d] [Batch %d/%d] [D loss : %f] [G loss : %f]' %
                          (epoch + 1, epochs, i + 1, len(train_data), self.disc_loss.item(), self.gen_loss.item()))
            d_mean = d_sum / max(n, 1)
            self.loss_dict['d loss'].append(d_mean[0])
            self.loss_dict['d real loss'].append(d_mean[1])
            self.loss_dict['d fake loss'].append(d_mean[2])
            self.loss_dict['g loss'].append(np.atleast_1d(g_sum)[0])   # summed over the epoch's batches, not averaged, in the reference
            if self.result_dire and epoch == epochs - 1:   # reference :614-615
                self._save_checkpoints('last_epoch')


def parse_args(argv=None):
    """The reference's flags [:765-775]; see gemmgan_b200/cli.py."""
    from gemmgan_b200.cli import build_parser

    return build_parser('vanilla').parse_args(argv)


if __name__ == '__main__':
    from gemmgan_b200.cli import main

    main('vanilla')
