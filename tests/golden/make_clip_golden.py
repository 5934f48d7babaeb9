"""Generate tests/golden/clip_*.npz with an implementation independent of oracle/clip_ref.py:
HuggingFace `transformers` CLIPModel (in-container, transformers 5.5), loaded with the same seeded
weights (oracle.clip_ref.make_state_dict) mapped to its parameter names.  Run from the repo root:
    python tests/golden/make_clip_golden.py
The .npz files hold un-normalised fp32 features for seeded synthetic inputs; tests regenerate the
weights and inputs from the seeds, so only the small outputs are committed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import clip_ref  # noqa: E402


def to_hf(sd, cfg):
    from transformers import CLIPConfig, CLIPModel, CLIPTextConfig, CLIPVisionConfig

    act = "quick_gelu" if cfg.quick_gelu else "gelu"
    vc = CLIPVisionConfig(hidden_size=cfg.vision.width, intermediate_size=cfg.vision.mlp, num_hidden_layers=cfg.vision.layers,
                          num_attention_heads=cfg.vision.heads, image_size=cfg.image_size, patch_size=cfg.patch,
                          hidden_act=act, projection_dim=cfg.embed_dim, layer_norm_eps=1e-5)
    tc = CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.text.width, intermediate_size=cfg.text.mlp,
                        num_hidden_layers=cfg.text.layers, num_attention_heads=cfg.text.heads,
                        max_position_embeddings=cfg.context_length, hidden_act=act, projection_dim=cfg.embed_dim,
                        layer_norm_eps=1e-5, eos_token_id=cfg.vocab_size - 1, bos_token_id=cfg.vocab_size - 2, pad_token_id=0)
    model = CLIPModel(CLIPConfig(text_config=tc.to_dict(), vision_config=vc.to_dict(), projection_dim=cfg.embed_dim))
    model = model.float().eval()
    new = {}

    def tower(src, dst, t):
        w = t.width
        for i in range(t.layers):
            s = "%stransformer.resblocks.%d." % (src, i)
            d = "%sencoder.layers.%d." % (dst, i)
            wq, wk, wv = sd[s + "attn.in_proj_weight"].split(w, dim=0)
            bq, bk, bv = sd[s + "attn.in_proj_bias"].split(w, dim=0)
            new[d + "self_attn.q_proj.weight"], new[d + "self_attn.q_proj.bias"] = wq, bq
            new[d + "self_attn.k_proj.weight"], new[d + "self_attn.k_proj.bias"] = wk, bk
            new[d + "self_attn.v_proj.weight"], new[d + "self_attn.v_proj.bias"] = wv, bv
            new[d + "self_attn.out_proj.weight"] = sd[s + "attn.out_proj.weight"]
            new[d + "self_attn.out_proj.bias"] = sd[s + "attn.out_proj.bias"]
            new[d + "layer_norm1.weight"], new[d + "layer_norm1.bias"] = sd[s + "ln_1.weight"], sd[s + "ln_1.bias"]
            new[d + "layer_norm2.weight"], new[d + "layer_norm2.bias"] = sd[s + "ln_2.weight"], sd[s + "ln_2.bias"]
            new[d + "mlp.fc1.weight"], new[d + "mlp.fc1.bias"] = sd[s + "mlp.c_fc.weight"], sd[s + "mlp.c_fc.bias"]
            new[d + "mlp.fc2.weight"], new[d + "mlp.fc2.bias"] = sd[s + "mlp.c_proj.weight"], sd[s + "mlp.c_proj.bias"]

    tower("visual.", "vision_model.", cfg.vision)
    tower("", "text_model.", cfg.text)
    new["vision_model.embeddings.patch_embedding.weight"] = sd["visual.conv1.weight"]
    new["vision_model.embeddings.class_embedding"] = sd["visual.class_embedding"]
    new["vision_model.embeddings.position_embedding.weight"] = sd["visual.positional_embedding"]
    new["vision_model.pre_layrnorm.weight"], new["vision_model.pre_layrnorm.bias"] = sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"]
    new["vision_model.post_layernorm.weight"], new["vision_model.post_layernorm.bias"] = sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]
    new["visual_projection.weight"] = sd["visual.proj"].t().contiguous()
    new["text_model.embeddings.token_embedding.weight"] = sd["token_embedding.weight"]
    new["text_model.embeddings.position_embedding.weight"] = sd["positional_embedding"]
    new["text_model.final_layer_norm.weight"], new["text_model.final_layer_norm.bias"] = sd["ln_final.weight"], sd["ln_final.bias"]
    new["text_projection.weight"] = sd["text_projection"].t().contiguous()
    new["logit_scale"] = sd["logit_scale"]
    missing, unexpected = model.load_state_dict(new, strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model


def feats(out):
    return out.pooler_output if hasattr(out, "pooler_output") else out


def main():
    torch.set_num_threads(8)
    for name, n in (("tiny", 5), ("tiny-gelu", 5), ("ViT-B/32", 4)):
        cfg = clip_ref.CONFIGS[name]
        sd = clip_ref.make_state_dict(cfg, seed=0)
        model = to_hf(sd, cfg)
        px = clip_ref.synth_images(n, cfg, seed=0)
        tk = clip_ref.synth_tokens(n, cfg, seed=0)
        with torch.no_grad():
            fi = feats(model.get_image_features(pixel_values=px)).float().numpy()
            ft = feats(model.get_text_features(input_ids=tk, attention_mask=torch.ones_like(tk))).float().numpy()
        out = os.path.join(ROOT, "tests", "golden", "clip_%s.npz" % name.replace("/", "-"))
        np.savez_compressed(out, image_features=fi, text_features=ft, n=n, seed=0, source="transformers CLIPModel fp32")
        oi = clip_ref.encode_image(sd, cfg, px).numpy()
        ot = clip_ref.encode_text(sd, cfg, tk).numpy()
        print(name, "HF vs oracle max abs diff: image %.3g text %.3g (|f| ~ %.3g)" % (
            np.abs(fi - oi).max(), np.abs(ft - ot).max(), np.abs(fi).mean()))


if __name__ == "__main__":
    main()
