"""Drop-in for the reference's src/conditional_gan_attention.py (WGAN-GP conditioned on ONE cross-attention: the
encoded text vector queries the encoded patches; the generator batch-normalises the attended vector) backed by the
sm_100a engine (GG_VARIANT_ATTN).

Same public names and signatures as the reference (file:line of the reference in brackets):
  losses [:26-40], build_* [:50-90], generator [:92-133], discriminator [:136-170], WGAN_GP_model [:172-190],
  WGAN_GP [:193-521] with init_train, build_WGAN_GP, gradient_penalty, train_disc, train_gen, train,
  generate_samples, generate_samples_all, fit(train_data, test_data, epochs, val). Model argument order:
  (x, text_embedding, patches, padding_mask) [:113]; train() order: (gene, text, patches, pad) [:392]. No gradient
  clipping, no dropout. The reference's forward prints the BatchNorm input / output of every call [:122-127]; the
  prints are not reproduced.

BatchNorm1d (generator.attn_bn) follows nn.BatchNorm1d: batch statistics in every training-mode generator forward (the
critic steps run the generator in training mode too, as the reference does: it never calls gen.eval() inside train()),
running_mean / running_var / num_batches_tracked updated per forward, running statistics in generate_samples (eval).
Its statistics are per process: like the reference's single-GPU script this variant is not offered data-parallel.
"""
from __future__ import annotations

import argparse

import torch

from conditional_gan_film import WGAN_GP as _FilmStyleTrainer
from gemmgan_b200.models import AttnDiscriminator, AttnGenerator, build_linear_block, build_stack  # noqa: F401
from gemmgan_b200.trainer import D_loss, G_loss, save_numpy, wasserstein_loss  # noqa: F401


def build_generator(input_dims, generator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, generator_dims, negative_slope, is_bn)


def build_discriminator(input_dims, dicriminator_dims, negative_slope=0.0, is_bn=False):
    return build_stack(input_dims, dicriminator_dims, negative_slope, is_bn)


class generator(AttnGenerator):
    pass


class discriminator(AttnDiscriminator):
    pass


def WGAN_GP_model(latent_dims, vector_dims, embedding_dims, generator_dims, discriminator_dims,
                  text_embedding_dims=768, patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
    gen = generator(latent_dims, embedding_dims, generator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)
    disc = discriminator(vector_dims, embedding_dims, discriminator_dims, text_embedding_dims,
                         patches_embedding_dims, negative_slope, is_bn)
    return gen, disc


class WGAN_GP(_FilmStyleTrainer):
    """Same constructor, staging and entry points as the film trainer (the two reference scripts share them: the batch
    tuple, the argument orders and the training loop are identical); what differs is the model and the BatchNorm
    bookkeeping."""

    variant = "attn"

    def _shape_cfg(self):
        return dict(E=self.embedding_dims, H=self.generator_dims[0], Dt=self.text_embedding_dims,
                    Dp=self.patches_embedding_dims, P=self._tokens, T=1, tower_bias=True)

    def build_WGAN_GP(self):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1:
            raise NotImplementedError("conditional_gan_attention normalises over the per-process batch (BatchNorm1d): "
                                      "data-parallel replicas would not be equivalent to the reference's single process")
        self.numerical_dims = []
        self.dropout_p = 0.0   # nn.MultiheadAttention(dropout=0.0) and no encoder layers: nothing is dropped
        gen, disc = WGAN_GP_model(self.latent_dims, self.input_dims, self.embedding_dims, self.generator_dims,
                                  self.discriminator_dims, self.text_embedding_dims, self.patches_embedding_dims,
                                  self.negative_slope, self.is_bn)
        self._attach(gen, disc)

    def _count_bn_batches(self, n):
        # nn.BatchNorm1d.num_batches_tracked (only read when momentum is None; kept for state_dict equality)
        self.gen.attn_bn.num_batches_tracked += n

    @staticmethod
    def _need_two_rows(t):
        # nn.BatchNorm1d in training mode refuses a single row (torch/nn/functional.py::_verify_batch_size); the
        # generator's BatchNorm sees every batch of train_disc / train_gen in training mode [:322, :371]
        if t.shape[0] == 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(t.shape)}")

    def train_disc(self, real_data, z, text_embedding, patches, padding_mask, alpha=None):
        counted = self.gen.training        # an eval-mode generator (after generate_samples) does not track batches
        if counted:
            self._need_two_rows(z)
        super().train_disc(real_data, z, text_embedding, patches, padding_mask, alpha)
        self._count_bn_batches(int(counted))

    def train_gen(self, z, text_embedding, patches, padding_mask):
        self._need_two_rows(z)
        super().train_gen(z, text_embedding, patches, padding_mask)
        self._count_bn_batches(1)

    def train(self, gene_expression, text_embedding, patches, padding_mask, zs=None, alphas=None, prefetch=None):
        self._need_two_rows(gene_expression)
        counted = self.n_critic if self.gen.training else 0   # critic steps with the generator in eval mode do not count
        super().train(gene_expression, text_embedding, patches, padding_mask, zs, alphas, prefetch)
        self._count_bn_batches(counted + 1)

    def _module_forward(self, module, x, text_embedding, patches, padding_mask):
        out = super()._module_forward(module, x, text_embedding, patches, padding_mask)
        if module is self.gen and module.training:
            self._count_bn_batches(1)
        return out

    def generate_samples_all(self, data_loader, num_repeats=1, balanced=False, balanced_max_oversample=5):
        """(real, generated, disease types real, disease types generated) [:407-506]; balanced=True = the class-balanced
        branch [:409-480]."""
        return self._generate_all_film_layout(data_loader, num_repeats, balanced, balanced_max_oversample, with_site=False)

    def fit(self, train_data, test_data=None, epochs=1, val=True):
        """Training loop of the reference fit() [:523-600] (learning rates halve every 50 epochs) without its
        evaluation / plotting."""
        self._lr_decay_every = 50
        return super().fit(train_data, None, test_data, epochs, val)


def parse_args(argv=None):
    """The reference's flags [see gemmgan_b200/cli.py]."""
    from gemmgan_b200.cli import build_parser

    return build_parser('attn').parse_args(argv)


if __name__ == '__main__':
    from gemmgan_b200.cli import main

    main('attn')
