"""Drop-in for the reference's src/conditional_gan_cross_attention.py — cross-attention fusion of patch and
text tokens WITHOUT FiLM (WGAN-GP) — backed by the sm_100a engine.

Same public names and signatures as the reference (file:line of the reference in brackets):
  wasserstein_loss / G_loss / D_loss [:32-46], build_linear_block / build_generator / build_discriminator
  [:56-95], generator [:97-150], discriminator [:153-206], WGAN_GP_model [:209-225], WGAN_GP [:228-...] with
  init_train, build_WGAN_GP, gradient_penalty [:323], train_disc [:348], train_gen [:396], train [:433],
  generate_samples_all [:449], generate_samples [:571], set_requires_grad, fit [:589], print_best_epoch.
Differences from the paper model (conditional_gan_cross_attention_with_film.py), as in the reference:
  * no film_generator; the encoder layers and both MultiheadAttentions are built with bias=False [:112-121];
  * every text token queries the patches (and the result queries the text tokens) [:136-139], but only row 0
    of either result reaches the conditioning vector [:140-142] and attention rows are independent, so the
    engine evaluates query row 0 only — the same single-query tail as the paper model, identical results;
  * no gradient clipping in train_disc / train_gen (the clip_grad_norm_ calls of the paper model are absent).
"""
from __future__ import annotations

import torch

import conditional_gan_cross_attention_with_film as _paper
from conditional_gan_cross_attention_with_film import (D_loss, G_loss, build_discriminator,  # noqa: F401
                                                        build_generator, build_linear_block, save_numpy,
                                                        wasserstein_loss)
from gemmgan_b200.models import CrossDiscriminator, CrossGenerator


class generator(CrossGenerator):
    pass


class discriminator(CrossDiscriminator):
    pass


def WGAN_GP_model(latent_dims, vector_dims, embedding_dims, generator_dims, discriminator_dims,
                  text_embedding_dims=768, patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
    gen = generator(latent_dims, embedding_dims, generator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)
    disc = discriminator(vector_dims, embedding_dims, discriminator_dims, text_embedding_dims,
                         patches_embedding_dims, negative_slope, is_bn)
    return gen, disc


class WGAN_GP(_paper.WGAN_GP):
    """Same constructor kwargs, entry points and batch tuple layout as the paper model's trainer."""

    variant = "cross"
    clip_d = clip_g = None

    def _shape_cfg(self):
        cfg = super()._shape_cfg()
        cfg["tower_bias"] = False
        return cfg

    def build_WGAN_GP(self):
        self.numerical_dims = []
        gen, disc = WGAN_GP_model(self.latent_dims, self.input_dims, self.embedding_dims, self.generator_dims,
                                  self.discriminator_dims, self.text_embedding_dims, self.patches_embedding_dims,
                                  self.negative_slope, self.is_bn)
        self._attach(gen, disc)


def parse_args(argv=None):
    """The reference's flags [see gemmgan_b200/cli.py]."""
    from gemmgan_b200.cli import build_parser

    return build_parser('cross').parse_args(argv)


if __name__ == '__main__':
    from gemmgan_b200.cli import main

    main('cross')
