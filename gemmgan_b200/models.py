"""Model classes behind the drop-in modules at the repo root.

The parameters are ordinary nn.Parameters created by the same torch constructors, in the same order
and under the same attribute names as the reference classes, so torch.manual_seed(s) reproduces the
reference's initial weights and state_dict()s are interchangeable with reference checkpoints
(including the never-used prototype `patches_transformer_layer.*` entries):
  paper   generator/discriminator  src/conditional_gan_cross_attention_with_film.py:97-233
  film    generator/discriminator  src/conditional_gan_film.py:97-204
  cross   generator/discriminator  src/conditional_gan_cross_attention.py:97-206
  concat  generator/discriminator  src/conditional_gan_concat.py:97-196
  img     generator/discriminator  src/conditional_gan_img_transformer.py:97-190
  vanilla generator_nocond/discriminator_nocond  src/vanilla_gan_unconditional.py:93-184
  label   generator/discriminator  src/benchmark_generative_model.py:101-236 (label-conditioned baseline)
  attn    generator/discriminator  src/conditional_gan_attention.py:92-170 (one MultiheadAttention, BatchNorm1d in G)

forward() does not run torch kernels: it calls the engine (libgemmgan_sm100a.so). Inside a drop-in trainer it is the
trainer's engine (an inference forward: the training step uses the engine's fused, hand-written backward instead,
gemmgan_b200/trainer.py). A free-standing module (no trainer) is differentiable: its forward goes through one
torch.autograd.Function over the engine's forward / backward entry points (gemmgan_b200/standalone.py), first order.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _abi_decl as A

VARIANT_IDS = {"vanilla": A.VARIANT_VANILLA, "film": A.VARIANT_FILM, "paper": A.VARIANT_PAPER,
               "cross": A.VARIANT_CROSS, "concat": A.VARIANT_CONCAT, "img": A.VARIANT_IMG, "label": A.VARIANT_LABEL,
               "attn": A.VARIANT_ATTN}


def build_linear_block(input_dims, output_dims, negative_slope=0.0, is_bn=False):
    """Linear + LeakyReLU(negative_slope) (reference :56-72). BatchNorm (is_bn=True) is not on any
    script's path (every __main__ passes is_bn=False) and is rejected."""
    if is_bn:
        raise NotImplementedError("is_bn=True is not used by any reference script and is not data-parallel safe")
    return nn.Sequential(nn.Linear(input_dims, output_dims), nn.LeakyReLU(negative_slope=negative_slope))


def build_stack(input_dims, dims, negative_slope=0.0, is_bn=False):
    """nn.ModuleList of linear blocks (reference build_generator / build_discriminator :76-95)."""
    stack = nn.ModuleList()
    for i, d in enumerate(dims):
        stack.append(build_linear_block(input_dims if i == 0 else dims[i - 1], d, negative_slope, is_bn))
    return stack


class _Net(nn.Module):
    """Common body; subclasses fix role / variant and the reference constructor signature."""

    _role = "gen"
    _variant = "paper"

    def _build(self, first_dim, embedding_dims, dims, text_embedding_dims, patches_embedding_dims,
               negative_slope, is_bn):
        v = self._variant
        self.embedding_dims = embedding_dims
        self.text_embedding_dims = text_embedding_dims
        self.patches_embedding_dims = patches_embedding_dims
        self.negative_slope = negative_slope
        self.is_bn = is_bn
        E = embedding_dims
        if v in ("paper", "film", "cross", "img"):
            if v in ("paper", "film"):
                self.film_generator = nn.Linear(text_embedding_dims, patches_embedding_dims * 2)
            if v in ("paper", "cross"):
                self.text_encoder = nn.Linear(text_embedding_dims, E)
            if v == "img":  # conditional_gan_img_transformer.py:111-115
                self.patches_encoder = nn.Sequential(nn.Linear(patches_embedding_dims, E), nn.ReLU(), nn.LayerNorm(E))
            else:
                self.patches_encoder = nn.Linear(patches_embedding_dims, E)
            self.patches_transformer_layer = nn.TransformerEncoderLayer(
                d_model=E, nhead=4, dim_feedforward=E * 2, dropout=0.1, activation="relu", batch_first=True,
                bias=(v == "paper"))
            self.patches_cls_token = nn.Parameter(torch.empty(1, 1, E))
            torch.nn.init.trunc_normal_(self.patches_cls_token, std=0.02)
            self.patches_transformer = nn.TransformerEncoder(self.patches_transformer_layer, num_layers=2)
            if v in ("paper", "cross"):
                ab = v == "paper"  # conditional_gan_cross_attention.py:118-121 builds its attentions with bias=False
                self.patch2text_attention = nn.MultiheadAttention(embed_dim=E, num_heads=4, batch_first=True, bias=ab)
                self.text2patch_attention = nn.MultiheadAttention(embed_dim=E, num_heads=4, batch_first=True, bias=ab)
        self.input_dims = first_dim + (0 if v == "vanilla" else E)
        stack = build_stack(self.input_dims, dims[:-1], negative_slope, is_bn)
        setattr(self, "generator" if self._role == "gen" else "discriminator", stack)
        self.final_layer = nn.Linear(dims[-2], dims[-1])
        self._gg_owner = None  # set by the trainer: object with ._module_forward(module, *args)

    # -- engine plumbing ------------------------------------------------------------------
    def trunk_blocks(self):
        return self.generator if self._role == "gen" else self.discriminator

    def slot_table(self):
        """C-ABI parameter slot -> nn.Parameter (include/gemmgan.h enum gg_param_slot)."""
        t = {}
        v = self._variant
        if v in ("paper", "film", "cross", "img"):
            if v in ("paper", "film"):
                t[A.P_FILM_W], t[A.P_FILM_B] = self.film_generator.weight, self.film_generator.bias
            if v in ("paper", "cross"):
                t[A.P_TEXT_W], t[A.P_TEXT_B] = self.text_encoder.weight, self.text_encoder.bias
            if v == "img":
                lin, norm = self.patches_encoder[0], self.patches_encoder[2]
                t[A.P_PATCH_W], t[A.P_PATCH_B] = lin.weight, lin.bias
                t[A.P_PENC_LN_W], t[A.P_PENC_LN_B] = norm.weight, norm.bias
            else:
                t[A.P_PATCH_W], t[A.P_PATCH_B] = self.patches_encoder.weight, self.patches_encoder.bias
            t[A.P_CLS] = self.patches_cls_token
            for l, layer in enumerate(self.patches_transformer.layers):
                b = A.P_LAYER0 + A.L_COUNT * l
                t[b + A.L_IN_W], t[b + A.L_IN_B] = layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias
                t[b + A.L_OUT_W], t[b + A.L_OUT_B] = layer.self_attn.out_proj.weight, layer.self_attn.out_proj.bias
                t[b + A.L_FF1_W], t[b + A.L_FF1_B] = layer.linear1.weight, layer.linear1.bias
                t[b + A.L_FF2_W], t[b + A.L_FF2_B] = layer.linear2.weight, layer.linear2.bias
                t[b + A.L_N1_W], t[b + A.L_N1_B] = layer.norm1.weight, layer.norm1.bias
                t[b + A.L_N2_W], t[b + A.L_N2_B] = layer.norm2.weight, layer.norm2.bias
            if v in ("paper", "cross"):
                p2t, t2p = self.patch2text_attention, self.text2patch_attention
                t[A.P_P2T_IN_W], t[A.P_P2T_IN_B] = p2t.in_proj_weight, p2t.in_proj_bias
                t[A.P_P2T_OUT_W], t[A.P_P2T_OUT_B] = p2t.out_proj.weight, p2t.out_proj.bias
                t[A.P_T2P_IN_W], t[A.P_T2P_IN_B] = t2p.in_proj_weight, t2p.in_proj_bias
                t[A.P_T2P_OUT_W], t[A.P_T2P_OUT_B] = t2p.out_proj.weight, t2p.out_proj.bias
        blocks = self.trunk_blocks()
        if len(blocks) != 2:
            raise NotImplementedError("the engine implements the reference's 2-hidden-layer trunks "
                                      "(generator_dims=[h,h,G], discriminator_dims=[h,h,1])")
        t[A.P_TR0_W], t[A.P_TR0_B] = blocks[0][0].weight, blocks[0][0].bias
        t[A.P_TR1_W], t[A.P_TR1_B] = blocks[1][0].weight, blocks[1][0].bias
        t[A.P_FIN_W], t[A.P_FIN_B] = self.final_layer.weight, self.final_layer.bias
        return {k: p for k, p in t.items() if p is not None}

    def _engine_forward(self, *args):
        if self._gg_owner is None:   # free-standing module: autograd over the engine (raises off a CUDA device)
            from .standalone import owner_of
            return owner_of(self)._module_forward(self, *args)
        return self._gg_owner._module_forward(self, *args)


# ----------------------------------------------------------------------------- paper model
class PaperGenerator(_Net):
    _role, _variant = "gen", "paper"

    def __init__(self, latent_dims, embedding_dims, generator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.generator_dims = generator_dims
        self._build(latent_dims, embedding_dims, generator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)

    def forward(self, gene_expression, patches, patches_padding_mask, text_tokens, text_padding_mask):
        return self._engine_forward(gene_expression, patches, patches_padding_mask, text_tokens, text_padding_mask)


class PaperDiscriminator(_Net):
    _role, _variant = "disc", "paper"

    def __init__(self, vector_dims, embedding_dims, discriminator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.discriminator_dims = discriminator_dims
        self._build(vector_dims, embedding_dims, discriminator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)

    def forward(self, gene_expression, patches, patches_padding_mask, text_tokens, text_padding_mask):
        return self._engine_forward(gene_expression, patches, patches_padding_mask, text_tokens, text_padding_mask)


# ------------------------------------------------------------ cross-attention model (no FiLM)
class CrossGenerator(PaperGenerator):
    _role, _variant = "gen", "cross"


class CrossDiscriminator(PaperDiscriminator):
    _role, _variant = "disc", "cross"


# ------------------------------------------------------------------------------ film model
class FilmGenerator(_Net):
    _role, _variant = "gen", "film"

    def __init__(self, latent_dims, embedding_dims, generator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.generator_dims = generator_dims
        self.embeding_dims = embedding_dims
        self._build(latent_dims, embedding_dims, generator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)

    def forward(self, x, text_embedding, patches, padding_mask):
        return self._engine_forward(x, text_embedding, patches, padding_mask)


class FilmDiscriminator(_Net):
    _role, _variant = "disc", "film"

    def __init__(self, vector_dims, embedding_dims, discriminator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.discriminator_dims = discriminator_dims
        self._build(vector_dims, embedding_dims, discriminator_dims, text_embedding_dims, patches_embedding_dims,
                    negative_slope, is_bn)

    def forward(self, gene_expression, text_embedding, patches, padding_mask):
        return self._engine_forward(gene_expression, text_embedding, patches, padding_mask)


# ------------------------------------------------- image-transformer model (no text, no FiLM)
class ImgGenerator(FilmGenerator):
    _role, _variant = "gen", "img"


class ImgDiscriminator(FilmDiscriminator):
    _role, _variant = "disc", "img"


# ---------------------------------------------------------------------------- concat model
class _ConcatNet(_Net):
    """conditional_gan_concat.py: conditioning vector = encoder(text embedding) ('text') or the masked mean over
    the patches of encoder(patch embedding) ('image'). Construction order as in the reference: trunk blocks,
    final_layer, then the encoder (:119-124 / :172-176)."""

    _variant = "concat"

    def _build_concat(self, first_dim, input_embedding_dims, embedding_dims, dims, condition_type, negative_slope,
                      is_bn):
        assert condition_type in ['text', 'image', 'both'], \
            "Condition type must be either 'text' or 'image' or 'both'"
        self.embedding_dims = embedding_dims
        self.input_embedding_dims = input_embedding_dims
        self.negative_slope = negative_slope
        self.is_bn = is_bn
        self.input_dims = first_dim + embedding_dims
        stack = build_stack(self.input_dims, dims[:-1], negative_slope, is_bn)
        setattr(self, "generator" if self._role == "gen" else "discriminator", stack)
        self.final_layer = nn.Linear(dims[-2], dims[-1])
        self.condition_type = condition_type
        self._gg_owner = None

    def slot_table(self):
        blocks = self.trunk_blocks()
        if len(blocks) != 2:
            raise NotImplementedError("the engine implements the reference's 2-hidden-layer trunks")
        return {A.P_TEXT_W: self.encoder.weight, A.P_TEXT_B: self.encoder.bias,
                A.P_TR0_W: blocks[0][0].weight, A.P_TR0_B: blocks[0][0].bias,
                A.P_TR1_W: blocks[1][0].weight, A.P_TR1_B: blocks[1][0].bias,
                A.P_FIN_W: self.final_layer.weight, A.P_FIN_B: self.final_layer.bias}

    def forward(self, x, text_embedding, patches, padding_mask):
        return self._engine_forward(x, text_embedding, patches, padding_mask)


class ConcatGenerator(_ConcatNet):
    _role = "gen"

    def __init__(self, latent_dims, input_embedding_dims, embedding_dims, generator_dims, condition_type='text',
                 negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.generator_dims = generator_dims
        self._build_concat(latent_dims, input_embedding_dims, embedding_dims, generator_dims, condition_type,
                           negative_slope, is_bn)
        self.final_activation = nn.ReLU()
        self.relu = nn.ReLU()
        self.encoder = nn.Linear(input_embedding_dims, embedding_dims)


class ConcatDiscriminator(_ConcatNet):
    _role = "disc"

    def __init__(self, vector_dims, input_embedding_dims, embedding_dims, discriminator_dims, condition_type='text',
                 negative_slope=0.0, is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.discriminator_dims = discriminator_dims
        self._build_concat(vector_dims, input_embedding_dims, embedding_dims, discriminator_dims, condition_type,
                           negative_slope, is_bn)
        self.encoder = nn.Linear(input_embedding_dims, embedding_dims)

    def forward(self, gene_expression, text_embedding, patches, padding_mask):   # the reference's argument name (:178)
        return self._engine_forward(gene_expression, text_embedding, patches, padding_mask)


# ----------------------------------------------------------------- single-attention model
class _AttnNet(_Net):
    """conditional_gan_attention.py: conditioning vector = MultiheadAttention(query = text_encoder(text embedding),
    keys / values = patches_encoder(patches), key_padding_mask) (:113-120 / :154-159); the generator normalises it with
    BatchNorm1d over the batch (`attn_bn`, :108, :126). Construction order as in the reference: text_encoder,
    patches_encoder, attention, (attn_bn), trunk blocks, final_layer."""

    _variant = "attn"

    def _build_attn(self, first_dim, embedding_dims, dims, text_embedding_dims, patches_embedding_dims, negative_slope,
                    is_bn):
        E = embedding_dims
        self.embedding_dims = E
        self.text_embedding_dims = text_embedding_dims
        self.patches_embedding_dims = patches_embedding_dims
        self.is_bn = is_bn
        self.negative_slope = negative_slope
        self.text_encoder = nn.Linear(text_embedding_dims, E)
        self.patches_encoder = nn.Linear(patches_embedding_dims, E)
        self.attention = nn.MultiheadAttention(embed_dim=E, num_heads=4, batch_first=True)
        self.input_dims = first_dim + E
        if self._role == "gen":
            self.attn_bn = nn.BatchNorm1d(E)
        stack = build_stack(self.input_dims, dims[:-1], negative_slope, is_bn)
        setattr(self, "generator" if self._role == "gen" else "discriminator", stack)
        self.final_layer = nn.Linear(dims[-2], dims[-1])
        self._gg_owner = None

    def slot_table(self):
        blocks = self.trunk_blocks()
        if len(blocks) != 2:
            raise NotImplementedError("the engine implements the reference's 2-hidden-layer trunks")
        att = self.attention
        t = {A.P_TEXT_W: self.text_encoder.weight, A.P_TEXT_B: self.text_encoder.bias,
             A.P_PATCH_W: self.patches_encoder.weight, A.P_PATCH_B: self.patches_encoder.bias,
             A.P_P2T_IN_W: att.in_proj_weight, A.P_P2T_IN_B: att.in_proj_bias,
             A.P_P2T_OUT_W: att.out_proj.weight, A.P_P2T_OUT_B: att.out_proj.bias,
             A.P_TR0_W: blocks[0][0].weight, A.P_TR0_B: blocks[0][0].bias,
             A.P_TR1_W: blocks[1][0].weight, A.P_TR1_B: blocks[1][0].bias,
             A.P_FIN_W: self.final_layer.weight, A.P_FIN_B: self.final_layer.bias}
        if self._role == "gen":
            t[A.P_BN_W], t[A.P_BN_B] = self.attn_bn.weight, self.attn_bn.bias
        return t

    def forward(self, x, text_embedding, patches, padding_mask):
        return self._engine_forward(x, text_embedding, patches, padding_mask)


class AttnGenerator(_AttnNet):
    _role = "gen"

    def __init__(self, latent_dims, embedding_dims, generator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.generator_dims = generator_dims
        self._build_attn(latent_dims, embedding_dims, generator_dims, text_embedding_dims, patches_embedding_dims,
                         negative_slope, is_bn)


class AttnDiscriminator(_AttnNet):
    _role = "disc"

    def __init__(self, vector_dims, embedding_dims, discriminator_dims, text_embedding_dims=768,
                 patches_embedding_dims=1024, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.discriminator_dims = discriminator_dims
        self._build_attn(vector_dims, embedding_dims, discriminator_dims, text_embedding_dims, patches_embedding_dims,
                         negative_slope, is_bn)

    def forward(self, gene_expression, text_embedding, patches, padding_mask):   # the reference's argument name (:153)
        return self._engine_forward(gene_expression, text_embedding, patches, padding_mask)


# ------------------------------------------------------ label-conditioned baseline model
LABEL_EMBEDDING_DIMS = 128  # benchmark_generative_model.py:32 (hard-coded; 2 variables -> the 256 of :119 / :185)


def categorical_embedding(vocab_sizes):
    """One nn.Embedding(vs, 128) per categorical variable (reference :27-35)."""
    return nn.ModuleList(nn.Embedding(vs, LABEL_EMBEDDING_DIMS) for vs in vocab_sizes)


class _LabelNet(_Net):
    """benchmark_generative_model.py: conditioning vector = [emb0[disease type] | emb1[primary site]], concatenated
    to the trunk input. Construction order as in the reference: trunk blocks, final_layer, then the embedding
    tables (:122-126 / :190-196). `categorical_embedded_dims` is hard-coded to 256 there while every table is 128
    wide, so only two categorical variables give consistent shapes; anything else fails in the reference's first
    forward (mat1 / mat2 shape mismatch) and is rejected here at construction."""

    _variant = "label"

    def _build_label(self, first_dim, numerical_dims, vocab_sizes, dims, negative_slope, is_bn):
        self.numerical_dims = len(numerical_dims)
        self.vocab_sizes = vocab_sizes
        self.negative_slope = negative_slope
        self.n_cat_vars = len(vocab_sizes)
        self.categorical_embedded_dims = 256
        if self.n_cat_vars * LABEL_EMBEDDING_DIMS != self.categorical_embedded_dims or self.numerical_dims != 0:
            raise NotImplementedError("the reference's shapes are only consistent for two categorical variables and "
                                      "no numerical covariates (128-wide tables vs a hard-coded 256: :32, :119)")
        self.input_dims = first_dim + self.numerical_dims + self.categorical_embedded_dims
        stack = build_stack(self.input_dims, dims[:-1], negative_slope, is_bn)
        setattr(self, "generator" if self._role == "gen" else "discriminator", stack)
        self.final_layer = nn.Linear(dims[-2], dims[-1])
        self._gg_owner = None

    def slot_table(self):
        blocks = self.trunk_blocks()
        if len(blocks) != 2:
            raise NotImplementedError("the engine implements the reference's 2-hidden-layer trunks")
        return {A.P_EMB0: self.categorical_embedding[0].weight, A.P_EMB1: self.categorical_embedding[1].weight,
                A.P_TR0_W: blocks[0][0].weight, A.P_TR0_B: blocks[0][0].bias,
                A.P_TR1_W: blocks[1][0].weight, A.P_TR1_B: blocks[1][0].bias,
                A.P_FIN_W: self.final_layer.weight, A.P_FIN_B: self.final_layer.bias}

    def forward(self, x, categorical_covariates, categorical_covariates_2):
        return self._engine_forward(x, categorical_covariates, categorical_covariates_2)


class LabelGenerator(_LabelNet):
    _role = "gen"

    def __init__(self, latent_dims, numerical_dims, vocab_sizes, generator_dims, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.generator_dims = generator_dims
        self._build_label(latent_dims, numerical_dims, vocab_sizes, generator_dims, negative_slope, is_bn)
        self.final_activation = nn.ReLU()
        self.relu = nn.ReLU()
        self.categorical_embedding = categorical_embedding(vocab_sizes)


class LabelDiscriminator(_LabelNet):
    _role = "disc"

    def __init__(self, vector_dims, numerical_dims, vocab_sizes, discriminator_dims, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.discriminator_dims = discriminator_dims
        self._build_label(vector_dims, numerical_dims, vocab_sizes, discriminator_dims, negative_slope, is_bn)
        self.categorical_embedding = categorical_embedding(vocab_sizes)


# --------------------------------------------------------------------------- vanilla model
class VanillaGenerator(_Net):
    _role, _variant = "gen", "vanilla"

    def __init__(self, latent_dims, numerical_dims, vocab_sizes, generator_dims, negative_slope=0.0, is_bn=False):
        super().__init__()
        self.latent_dims = latent_dims
        self.numerical_dims = len(numerical_dims)
        self.vocab_sizes = vocab_sizes
        self.generator_dims = generator_dims
        self.n_cat_vars = len(vocab_sizes)
        self._build(latent_dims, 0, generator_dims, 0, 0, negative_slope, is_bn)
        self.final_activation = nn.ReLU()
        self.threshold = nn.Threshold(0, 0)

    def forward(self, x):
        return self._engine_forward(x)


class VanillaDiscriminator(_Net):
    _role, _variant = "disc", "vanilla"

    def __init__(self, vector_dims, numerical_dims, vocab_sizes, discriminator_dims, negative_slope=0.0,
                 is_bn=False):
        super().__init__()
        self.vector_dims = vector_dims
        self.numerical_dims = len(numerical_dims)
        self.vocab_sizes = vocab_sizes
        self.discriminator_dims = discriminator_dims
        self.n_cat_vars = len(vocab_sizes)
        self._build(vector_dims, 0, discriminator_dims, 0, 0, negative_slope, is_bn)

    def forward(self, x):
        return self._engine_forward(x)
