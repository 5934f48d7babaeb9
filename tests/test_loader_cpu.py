"""CPU tests of the a6 loader row: model-name resolution as all_clip's dispatcher reads it
(README.md:179,201,237; tests/test_clip_inference/test_mapper.py:11-15), checkpoint key-layout conversion,
and the tokenizer.  No GPU compute: nothing here touches a handle."""
import gzip
import json
import os

import numpy as np
import pytest
import torch

from oracle import clip_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_resolve_arch_accepts_the_names_the_reference_uses():
    import clip_retrieval_b200 as m
    from clip_retrieval_b200.model import resolve_arch

    # the three names of the reference's mapper test (test_mapper.py:11-15)
    a, _ = resolve_arch("ViT-B/32")
    assert a.quick_gelu and a.vision.width == 768 and a.patch == 32
    a, key = resolve_arch("open_clip:ViT-B-32/laion2b_s34b_b79k")
    assert not a.quick_gelu and key == "open_clip:ViT-B-32"
    a, _ = resolve_arch("hf_clip:patrickjohncyh/fashion-clip")
    assert a.quick_gelu and a.embed_dim == 512 and a.patch == 32
    # README.md:237 and open_clip's rule for OpenAI weights
    assert resolve_arch("open_clip:ViT-B-32-quickgelu")[0].quick_gelu
    assert resolve_arch("open_clip:ViT-B-32/openai")[0].quick_gelu
    assert resolve_arch("open_clip:ViT-L-14/openai")[0].quick_gelu
    assert not resolve_arch("open_clip:ViT-L-14/laion2b_s32b_b82k")[0].quick_gelu
    # the back ends of docs/laion5B_back.md:24 and docs/laion5B_h14_back.md:60
    assert resolve_arch("ViT-L/14")[0].embed_dim == 768
    h = resolve_arch("open_clip:ViT-H-14")[0]
    assert h.embed_dim == 1024 and h.vision.width == 1280 and not h.quick_gelu
    # README.md:201 (DeepSparse names): the architecture resolves; 256 px input for the DataComp export
    assert resolve_arch("nm:neuralmagic/CLIP-ViT-B-32-256x256-DataComp-s34B-b86K-quant-ds")[0].image_size == 256
    assert resolve_arch("synthetic:hf_clip:openai/clip-vit-large-patch14")[0].vision.layers == 24
    for bad in ("ViT-Z/99", "open_clip:nope", "hf_clip:someone/unknown"):
        with pytest.raises(ValueError):
            resolve_arch(bad)
    # every head dimension in the table is one the attention kernels implement
    for a in m.ARCHS.values():
        for t in (a.vision, a.text):
            assert t.width % t.heads == 0 and t.width // t.heads in (64, 80, 96, 128)


def test_arch_from_hf_config_matches_table():
    from clip_retrieval_b200.model import arch_from_hf_config, resolve_arch
    from transformers import CLIPConfig

    # transformers' default CLIPConfig is openai/clip-vit-base-patch32
    cfg = CLIPConfig().to_dict()
    cfg["text_config"]["vocab_size"] = 49408
    a = arch_from_hf_config(cfg)
    assert a == resolve_arch("hf_clip:openai/clip-vit-base-patch32")[0]


@pytest.mark.parametrize("name", ["tiny", "tiny-gelu"])
def test_convert_hf_state_dict_round_trips_bit_exactly(name):
    """tests/golden/make_clip_golden.py maps the oracle's weights INTO HuggingFace CLIPModel (the reference's
    `hf_clip:` backend); convert_hf_state_dict is the inverse and must give back every tensor bit for bit."""
    import importlib.util

    from clip_retrieval_b200.model import convert_hf_state_dict

    spec = importlib.util.spec_from_file_location("make_clip_golden", os.path.join(GOLDEN, "make_clip_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    cfg = clip_ref.CONFIGS[name]
    sd = clip_ref.make_state_dict(cfg, seed=0)
    hf = gen.to_hf(sd, cfg)
    back = convert_hf_state_dict(hf.state_dict(), cfg)
    for k, v in sd.items():
        if k == "logit_scale":
            continue
        assert k in back, k
        assert back[k].shape == v.shape and torch.equal(back[k], v), k
    assert set(back) == set(sd) - {"logit_scale"}


def test_read_checkpoint_layouts(tmp_path):
    """A state_dict saved plainly, wrapped in {"state_dict": ...} with a `module.` prefix, and as a
    HuggingFace CLIPModel state_dict all load to the same OpenAI-layout tensors."""
    import importlib.util

    from clip_retrieval_b200.model import read_checkpoint

    spec = importlib.util.spec_from_file_location("make_clip_golden", os.path.join(GOLDEN, "make_clip_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    cfg = clip_ref.CONFIGS["tiny"]
    sd = {k: v for k, v in clip_ref.make_state_dict(cfg, seed=0).items() if k != "logit_scale"}
    torch.save(sd, tmp_path / "a.pt")
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}, "epoch": 3}, tmp_path / "b.pt")
    torch.save(gen.to_hf(clip_ref.make_state_dict(cfg, seed=0), cfg).state_dict(), tmp_path / "c.bin")
    for f in ("a.pt", "b.pt", "c.bin"):
        got = read_checkpoint(str(tmp_path / f), cfg)
        assert set(sd) <= set(got), f
        for k, v in sd.items():
            assert torch.equal(got[k].float(), v), (f, k)


def test_tokenizer_matches_independent_bpe():
    """SimpleTokenizer (the object load_clip returns, used at reader.py:83,145 and clip_back.py:227) against
    ids produced by HuggingFace tokenizers' BPE in CLIPTokenizer's configuration on the same small merge table
    (tests/golden/make_tokenizer_golden.py), including non-ASCII letters and digits (one token per digit)."""
    from clip_retrieval_b200.model import SimpleTokenizer

    tk = SimpleTokenizer(os.path.join(GOLDEN, "bpe_tiny.txt.gz"))
    g = json.load(open(os.path.join(GOLDEN, "tokenizer_golden.json")))
    sot, eot = g["sot"], g["eot"]
    assert tk.encoder["<|startoftext|>"] == sot and tk.encoder["<|endoftext|>"] == eot
    for c in g["cases"]:
        assert [sot] + tk.encode(c["text"]) + [eot] == c["ids"], c["text"]
    # basic_clean of clip.simple_tokenizer: html entities are unescaped twice before tokenising
    assert tk.encode("&lt;b&gt; &amp;amp; ") == tk.encode("<b> &")
    # the callable contract: LongTensor [n, 77], zero padded, truncation keeps EOT last (HISTORY.md:47-49)
    out = tk(["a photo of a cat", "x" * 300])
    assert out.shape == (2, 77) and out.dtype == torch.long
    assert out[0, 0] == sot and out[0].max() == eot and out[0, int(out[0].argmax()) + 1:].sum() == 0
    assert out[1, 0] == sot and out[1, 76] == eot
    assert tk("one string").shape == (1, 77)
