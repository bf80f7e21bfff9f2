"""Drop-in for the reference's src/conditional_gan_concat.py (WGAN-GP conditioned by concatenating ONE linear
encoding of the text embedding -- or of the mean patch embedding -- to the trunk input) backed by the sm_100a engine.

Same public names and signatures as the reference (file:line of the reference in brackets):
  losses [:32-46], build_* [:56-95], generator [:97-148], discriminator [:151-196], WGAN_GP_model [:196-220],
  WGAN_GP [:223-...] with init_train, build_WGAN_GP [:304], gradient_penalty [:319], train_disc [:345],
  train_gen [:398], train [:437], generate_samples(_all), fit. Model argument order: (x, text_embedding, patches,
  padding_mask) [:129]; train() order: (gene, text, patches, pad) [:437]. No gradient clipping, no dropout.
condition_on='image' [:137-138]: the reference encodes every patch and takes the masked mean; the encoder is affine,
so the engine takes the masked mean of the patch embeddings first (gg_masked_mean_rows) and encodes once per sample.
condition_on='both' passes the reference's assert but has no branch in its forward (UnboundLocalError) -- rejected.
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from gemmgan_b200.models import ConcatDiscriminator, ConcatGenerator, build_linear_block, build_stack  # noqa: F401
from gemmgan_b200.trainer import D_loss, G_loss, TrainerBase, save_numpy, wasserstein_loss  # noqa: F401


def build_generator(input_dims, generator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, generator_dims, negative_slope, is_bn)


def build_discriminator(input_dims, dicriminator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, dicriminator_dims, negative_slope, is_bn)


class generator(ConcatGenerator):
    pass


class discriminator(ConcatDiscriminator):
    pass


def WGAN_GP_model(latent_dims, vector_dims, input_embedding_dims, embedding_dims, generator_dims, discriminator_dims,
                  condition_type='text', negative_slope=0.0, is_bn=False):
    gen = generator(latent_dims, input_embedding_dims, embedding_dims, generator_dims, condition_type, negative_slope,
                    is_bn)
    disc = discriminator(vector_dims, input_embedding_dims, embedding_dims, discriminator_dims, condition_type,
                         negative_slope, is_bn)
    return gen, disc


class WGAN_GP(TrainerBase):
    variant = "concat"

    def __init__(self, input_dims, latent_dims, embedding_dims, generator_dims, discriminator_dims,
                 input_embedding_dims=256, condition_on='text', negative_slope=0.0, is_bn=False,
                 lr_d=5e-4, lr_g=5e-4, optimizer='rms_prop', gp_weight=10, p_aug=0, norm_scale=0.5, train=True,
                 n_critic=5, freq_print=2, freq_compute_test=100, freq_visualize_test=100, patience=10,
                 normalization='standardize', log2=False, rpm=False, results_dire=''):
        if condition_on not in ('text', 'image'):
            raise NotImplementedError("condition_on='both' has no forward branch in the reference (:135-138)")
        self.condition_on = condition_on
        self.embedding_dims = embedding_dims
        self.input_embedding_dims = input_embedding_dims
        self._init_common(input_dims, latent_dims, generator_dims, discriminator_dims, negative_slope, is_bn,
                          lr_d, lr_g, optimizer, gp_weight, p_aug, norm_scale, train, n_critic, freq_print,
                          freq_compute_test, freq_visualize_test, patience, normalization, log2, rpm,
                          results_dire)
        self.dropout_p = 0.0     # no dropout layer anywhere in this variant

    def _shape_cfg(self):
        return dict(E=self.embedding_dims, H=self.generator_dims[0], Dt=self.input_embedding_dims,
                    Dp=self.input_embedding_dims, P=1, T=1, tower_bias=True)

    def build_WGAN_GP(self):
        self.numerical_dims = []
        gen, disc = WGAN_GP_model(self.latent_dims, self.input_dims, self.input_embedding_dims, self.embedding_dims,
                                  self.generator_dims, self.discriminator_dims, self.condition_on,
                                  self.negative_slope, self.is_bn)
        self._attach(gen, disc)

    def _stage(self, genes, text_embedding, patches, padding_mask):
        dev = self.device
        B = (text_embedding if self.condition_on == 'text' else patches).shape[0]
        eng = self._engine(B)
        if self.condition_on == 'text':
            vec = self._dev(text_embedding)
        else:  # masked mean of the patch embeddings, then ONE encoder GEMM (the encoder is affine)
            vec = eng.masked_mean_rows(self._dev(patches), self._dev(padding_mask))
        eng.set_batch(genes=None if genes is None else self._dev(genes), text=vec)
        return eng

    # ---- reference-signature entry points -------------------------------------------------
    def gradient_penalty(self, real_data, fake_data, text_embedding, patches, padding_mask, alpha=None):
        eng = self._stage(None, text_embedding, patches, padding_mask)
        if alpha is None:
            alpha = self._alpha(eng.B)
        return eng.gradient_penalty(real_data.to(self.device), fake_data.to(self.device), alpha,
                                    training=self.disc.training)

    def train_disc(self, real_data, z, text_embedding, patches, padding_mask, alpha=None):
        eng = self._stage(real_data, text_embedding, patches, padding_mask)
        self._train_disc_staged(eng, z.to(self.device), alpha)

    def train_gen(self, z, text_embedding, patches, padding_mask):
        eng = self._stage(None, text_embedding, patches, padding_mask)
        self._train_gen_staged(eng, z.to(self.device))

    def train(self, gene_expression, text_embedding, patches, padding_mask, zs=None, alphas=None, prefetch=None):
        eng = self._stage(gene_expression, text_embedding, patches, padding_mask)
        self._train_staged(eng, zs, alphas)
        if prefetch is not None:   # host tensors of the NEXT batch: their H2D copies overlap this step
            self.prefetch(*prefetch)

    def _module_forward(self, module, x, text_embedding, patches, padding_mask):
        eng = self._stage(None, text_embedding, patches, padding_mask)
        if module is self.gen:
            return eng.generate(x.to(self.device), training=module.training)
        return eng.critic(x.to(self.device), training=module.training)

    def generate_samples(self, gene_expression, text_embedding, patches, padding_mask):
        with torch.no_grad():
            self.gen.eval()
            x_real = gene_expression.clone().to(torch.float32)
            z = torch.normal(0, 1, size=(x_real.shape[0], self.latent_dims), device=self.device)
            x_gen = self.gen(z, text_embedding, patches, padding_mask)
        return x_real, x_gen

    def generate_samples_all(self, data_loader, num_repeats=1, balanced=False, balanced_max_oversample=5):
        """(real, generated, disease types real, disease types generated) [:453-552]; balanced=True = the class-balanced
        branch [:455-526]. Batch tuple layout of multi_patch_gan_dataloader.py:48."""
        return self._generate_all_film_layout(data_loader, num_repeats, balanced, balanced_max_oversample, with_site=False)

    def fit(self, train_data, val_data=None, test_data=None, epochs=1, val=True):
        """Training loop of the reference fit() without its evaluation / plotting."""
        self.build_WGAN_GP()
        if self.isTrain:
            self.init_train()
        for epoch in range(epochs):
            self._epoch_lr_decay(epoch, 50)   # both learning rates halve every 50 epochs in this script [:605-613]
            self.epoch = epoch
            d_sum, g_sum, n = 0.0, 0.0, 0
            for i, (data, nxt) in enumerate(self._lookahead(train_data)):
                self.train(data[1], data[0], data[2], data[3], prefetch=None if nxt is None else (nxt[1], nxt[0], nxt[2], nxt[3]))
                d_sum, g_sum, n = d_sum + self.d_batch_loss, g_sum + self.g_batch_loss, n + 1
                if (i + 1) % self.freq_print == 0:
                    print('[Epoch %d/%d] [Batch %d/%d] [D loss : %f] [G loss : %f]' %
                          (epoch + 1, epochs, i + 1, len(train_data), self.disc_loss.item(), self.gen_loss.item()))
            d_mean = d_sum / max(n, 1)
            self.loss_dict['d loss'].append(d_mean[0])
            self.loss_dict['d real loss'].append(d_mean[1])
            self.loss_dict['d fake loss'].append(d_mean[2])
            self.loss_dict['g loss'].append(np.atleast_1d(g_sum)[0])   # summed over the epoch's batches, not averaged, in the reference
            last = epoch == epochs - 1
            if self.result_dire and ((epoch + 1) % self.freq_compute_test == 0 or last):
                tag = 'last_epoch' if last else f'epoch_{epoch + 1}'
                self._save_checkpoints(tag)
            self._fit_evaluation(epoch, epochs, train_data, val_data, test_data, val)


def parse_args(argv=None):
    """The reference's flags [see gemmgan_b200/cli.py]."""
    from gemmgan_b200.cli import build_parser

    return build_parser('concat').parse_args(argv)


if __name__ == '__main__':
    from gemmgan_b200.cli import main

    main('concat')
