"""CPU restatement of the embed path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

What it restates: `model.encode_image` / `model.encode_text` followed by the L2-normalise and
cast of ClipMapper.__call__ (reference clip_retrieval/clip_inference/mapper.py:56-59,64-67) and of
KnnService.compute_query (clip_retrieval/clip_back.py:230-232,244-246).  The model arithmetic
lives in un-vendored dependencies reached through `all_clip.load_clip` (reference
requirements.txt:2,21,28: clip-anytorch>=2.5.0,<3, open-clip-torch>=2.0.0,<3.0.0,
all_clip>=1.3.0,<2), none of which is under /root/reference or installed here.  This file restates
the published CLIP architecture those packages implement (OpenAI CLIP `VisionTransformer` /
text `Transformer`, identical in open_clip), in plain fp32 tensor ops:
  vision: conv(patch p, stride p, no bias) == im2col GEMM with K index c*p*p + i*p + j ->
          prepend class_embedding -> + positional_embedding -> ln_pre ->
          L x { x += out_proj(MHA(ln_1 x)); x += c_proj(act(c_fc(ln_2 x))) } -> token 0 -> ln_post -> @ proj
  text:   token_embedding[tokens] + positional_embedding -> L x same block with a causal mask ->
          ln_final -> row at argmax(tokens) (EOT = largest id) -> @ text_projection
  MHA:    fused in_proj [3w, w] + bias, heads of w/h, scale (w/h)^-0.5, softmax over keys
  act:    QuickGELU x*sigmoid(1.702x) (OpenAI checkpoints) or exact erf GELU (LAION open_clip)
  LN eps 1e-5.
PIN: tests/golden/clip_*.npz hold embeddings computed by the independent in-container
implementation `transformers.models.clip.modeling_clip.CLIPModel` (transformers 5.5) on the same
seeded weights (tests/golden/make_clip_golden.py); tests/test_oracle_cpu.py checks this file
against them.  PARITY UNPINNED against the reference's own fixtures: the only golden vectors it
holds for this path, tests/test_clip_inference/test_embeddings/*.pkl (OpenAI ViT-B/32 weights),
need a checkpoint that is not on disk and cannot be downloaded.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch


@dataclass
class TowerCfg:
    width: int
    layers: int
    heads: int
    mlp: int


@dataclass
class ClipCfg:
    embed_dim: int
    image_size: int
    patch: int
    vision: TowerCfg
    text: TowerCfg
    context_length: int = 77
    vocab_size: int = 49408
    quick_gelu: bool = True
    name: str = ""

    @property
    def grid(self):
        return self.image_size // self.patch


CONFIGS = {
    # dims: SURVEY.md §8c (cross-checked there against HF random-init parameter counts)
    "ViT-B/32": ClipCfg(512, 224, 32, TowerCfg(768, 12, 12, 3072), TowerCfg(512, 12, 8, 2048), quick_gelu=True, name="ViT-B/32"),
    "ViT-B/16": ClipCfg(512, 224, 16, TowerCfg(768, 12, 12, 3072), TowerCfg(512, 12, 8, 2048), quick_gelu=True, name="ViT-B/16"),
    "ViT-L/14": ClipCfg(768, 224, 14, TowerCfg(1024, 24, 16, 4096), TowerCfg(768, 12, 12, 3072), quick_gelu=True, name="ViT-L/14"),
    "ViT-H/14": ClipCfg(1024, 224, 14, TowerCfg(1280, 32, 16, 5120), TowerCfg(1024, 24, 16, 4096), quick_gelu=False, name="ViT-H/14"),
    # small shapes for fast tests (same structure; odd sizes exercise tile tails)
    "tiny": ClipCfg(64, 64, 16, TowerCfg(128, 2, 2, 512), TowerCfg(64, 2, 1, 256), context_length=77, vocab_size=49408, quick_gelu=True, name="tiny"),
    "tiny-gelu": ClipCfg(96, 96, 32, TowerCfg(192, 3, 3, 768), TowerCfg(128, 2, 2, 512), context_length=77, vocab_size=49408, quick_gelu=False, name="tiny-gelu"),
}


def make_state_dict(cfg, seed=0, dtype=torch.float32):
    """Seeded random-init weights in the OpenAI/open_clip state_dict key layout (no checkpoint
    exists offline).  Scales follow the CLIP init recipe (attn std w^-0.5, proj std (2w)^-0.5*L^-0.5,
    fc std (2w)^-0.5, embeddings 0.02/0.01) with non-trivial LayerNorm affine and biases so that
    every term of the forward is exercised."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    sd = {}

    def tower(prefix, t):
        w, L = t.width, t.layers
        attn_std = w ** -0.5
        proj_std = (w ** -0.5) * ((2 * L) ** -0.5)
        fc_std = (2 * w) ** -0.5
        for i in range(L):
            p = "%stransformer.resblocks.%d." % (prefix, i)
            sd[p + "ln_1.weight"] = 1.0 + rn(w, std=0.1)
            sd[p + "ln_1.bias"] = rn(w, std=0.05)
            sd[p + "attn.in_proj_weight"] = rn(3 * w, w, std=attn_std)
            sd[p + "attn.in_proj_bias"] = rn(3 * w, std=0.02)
            sd[p + "attn.out_proj.weight"] = rn(w, w, std=proj_std)
            sd[p + "attn.out_proj.bias"] = rn(w, std=0.02)
            sd[p + "ln_2.weight"] = 1.0 + rn(w, std=0.1)
            sd[p + "ln_2.bias"] = rn(w, std=0.05)
            sd[p + "mlp.c_fc.weight"] = rn(t.mlp, w, std=fc_std)
            sd[p + "mlp.c_fc.bias"] = rn(t.mlp, std=0.02)
            sd[p + "mlp.c_proj.weight"] = rn(w, t.mlp, std=proj_std)
            sd[p + "mlp.c_proj.bias"] = rn(w, std=0.02)

    v, t = cfg.vision, cfg.text
    sd["visual.conv1.weight"] = rn(v.width, 3, cfg.patch, cfg.patch, std=(3 * cfg.patch * cfg.patch) ** -0.5)
    sd["visual.class_embedding"] = rn(v.width, std=v.width ** -0.5)
    sd["visual.positional_embedding"] = rn(cfg.grid ** 2 + 1, v.width, std=v.width ** -0.5)
    sd["visual.ln_pre.weight"] = 1.0 + rn(v.width, std=0.1)
    sd["visual.ln_pre.bias"] = rn(v.width, std=0.05)
    tower("visual.", v)
    sd["visual.ln_post.weight"] = 1.0 + rn(v.width, std=0.1)
    sd["visual.ln_post.bias"] = rn(v.width, std=0.05)
    sd["visual.proj"] = rn(v.width, cfg.embed_dim, std=v.width ** -0.5)
    sd["token_embedding.weight"] = rn(cfg.vocab_size, t.width, std=0.02)
    sd["positional_embedding"] = rn(cfg.context_length, t.width, std=0.01)
    tower("", t)
    sd["ln_final.weight"] = 1.0 + rn(t.width, std=0.1)
    sd["ln_final.bias"] = rn(t.width, std=0.05)
    sd["text_projection"] = rn(t.width, cfg.embed_dim, std=t.width ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07))
    return {k: v_.to(dtype) for k, v_ in sd.items()}


def synth_images(n, cfg, seed=0):
    """`image_tensor` batches as the reader produces them: fp32 NCHW in the range of the real
    preprocess output ([-1.80, 2.15])."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(n, 3, cfg.image_size, cfg.image_size, generator=g).clamp_(-1.80, 2.15)


def synth_tokens(n, cfg, seed=0):
    """`text_tokens` rows [n, 77] int64: SOT 49406, 1..75 random ids < 49406, EOT 49407, zero pad
    (no BPE vocabulary exists offline; EOT is the largest id, which is what pooling relies on)."""
    rng = np.random.default_rng(2000 + seed)
    T = cfg.context_length
    out = np.zeros((n, T), dtype=np.int64)
    for i in range(n):
        length = int(rng.integers(1, T - 1)) if i % 7 else T - 2  # every 7th row is full length
        out[i, 0] = cfg.vocab_size - 2
        out[i, 1:1 + length] = rng.integers(1, cfg.vocab_size - 2, size=length)
        out[i, 1 + length] = cfg.vocab_size - 1
    return torch.from_numpy(out)


def _ln(x, w, b):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _act(x, quick):
    return x * torch.sigmoid(1.702 * x) if quick else torch.nn.functional.gelu(x)


def _block(x, sd, p, heads, quick, mask):
    B, T, w = x.shape
    hd = w // heads
    h = _ln(x, sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
    qkv = h @ sd[p + "attn.in_proj_weight"].t() + sd[p + "attn.in_proj_bias"]
    q, k, v = qkv.split(w, dim=-1)
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, T, w)
    x = x + a @ sd[p + "attn.out_proj.weight"].t() + sd[p + "attn.out_proj.bias"]
    h = _ln(x, sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
    h = _act(h @ sd[p + "mlp.c_fc.weight"].t() + sd[p + "mlp.c_fc.bias"], quick)
    return x + h @ sd[p + "mlp.c_proj.weight"].t() + sd[p + "mlp.c_proj.bias"]


@torch.no_grad()
def encode_image(sd, cfg, pixels):
    """model.encode_image: fp32 NCHW [B,3,S,S] -> [B, D] (un-normalised)."""
    B = pixels.shape[0]
    p, g, w = cfg.patch, cfg.grid, cfg.vision.width
    cols = pixels.view(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * p * p)
    x = cols @ sd["visual.conv1.weight"].view(w, 3 * p * p).t()
    cls = sd["visual.class_embedding"].expand(B, 1, w)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]
    x = _ln(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    for i in range(cfg.vision.layers):
        x = _block(x, sd, "visual.transformer.resblocks.%d." % i, cfg.vision.heads, cfg.quick_gelu, None)
    x = _ln(x[:, 0], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return x @ sd["visual.proj"]


@torch.no_grad()
def encode_text(sd, cfg, tokens):
    """model.encode_text: int [B, 77] -> [B, D] (un-normalised)."""
    B, T = tokens.shape
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"][:T]
    mask = torch.full((T, T), float("-inf")).triu_(1)
    for i in range(cfg.text.layers):
        x = _block(x, sd, "transformer.resblocks.%d." % i, cfg.text.heads, cfg.quick_gelu, mask)
    x = _ln(x, sd["ln_final.weight"], sd["ln_final.bias"])
    x = x[torch.arange(B), tokens.argmax(dim=-1)]
    return x @ sd["text_projection"]


def mapper_image(sd, cfg, pixels):
    """mapper.py:57-59: encode, `/= norm(dim=-1, keepdim=True)` (no epsilon), fp16 numpy."""
    f = encode_image(sd, cfg, pixels)
    f = f / f.norm(dim=-1, keepdim=True)
    return f.to(torch.float16).numpy()


def mapper_text(sd, cfg, tokens):
    """mapper.py:65-67."""
    f = encode_text(sd, cfg, tokens)
    f = f / f.norm(dim=-1, keepdim=True)
    return f.to(torch.float16).numpy()


def query_embedding(sd, cfg, tokens=None, pixels=None):
    """clip_back.py:226-232 / :240-246: batch-1 embed, normalise, fp32 numpy [1, D]."""
    f = encode_text(sd, cfg, tokens) if tokens is not None else encode_image(sd, cfg, pixels)
    f = f / f.norm(dim=-1, keepdim=True)
    return f.to(torch.float32).numpy()


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))
