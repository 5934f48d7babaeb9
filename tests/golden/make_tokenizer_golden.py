"""Generate tests/golden/bpe_tiny.txt.gz + tokenizer_golden.json.

The CLIP BPE vocabulary (bpe_simple_vocab_16e6.txt.gz) ships inside clip / open_clip, neither of which is
installed offline, so the tokenizer of `load_clip` is pinned on a SMALL merge table in the same file format
against an independent implementation: HuggingFace `tokenizers`' Rust BPE configured as transformers 5.5
`CLIPTokenizer` configures it (same pre-tokenisation regex, byte-level alphabet, `</w>` suffix).
    python tests/golden/make_tokenizer_golden.py
"""
import collections
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CORPUS = """a photo of a cat sitting on the sofa . a photo of two dogs playing in the park , 2023 .
the café sells crème brûlée and jalapeño poppers for 12 euros ! über straße naïve façade
an illustration of the eiffel tower at night ; it's the artist's best work , isn't it ?
日本語のテキスト と 中文文本 mixed with english words 1234567890 times
photo photography photographer photos cats dogs dog's cat's we're they've i'm you'll he'd
""" * 3

TEXTS = [
    "a photo of a cat",
    "A Photo of 2 Dogs playing, in the PARK!",
    "café crème brûlée — jalapeño 12€",
    "it's the artist's best work, isn't it?",
    "日本語のテキスト 123 and naïve façade",
    "   multiple   spaces\tand\nnewlines  ",
    "x" * 300,
    "",
]


def train_merges(corpus, n_merges):
    import regex
    from clip_retrieval_b200.model import _bytes_to_unicode

    be = _bytes_to_unicode()
    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", regex.IGNORECASE)
    words = collections.Counter()
    for tok in pat.findall(corpus.lower()):
        t = "".join(be[b] for b in tok.encode("utf-8"))
        words[tuple(t[:-1]) + (t[-1] + "</w>",)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w[:-1], w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])
        merges.append(best)
        nw = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1]); i += 2
                else:
                    out.append(w[i]); i += 1
            nw[tuple(out)] += c
        words = nw
    return merges


def main():
    from clip_retrieval_b200.model import _bytes_to_unicode
    from transformers import CLIPTokenizer

    merges = train_merges(CORPUS, 300)
    path = os.path.join(ROOT, "tests", "golden", "bpe_tiny.txt.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(('"bpe_tiny_vocab" - version: 0.1\n' + "\n".join(" ".join(m) for m in merges) + "\n").encode("utf-8"))
    base = list(_bytes_to_unicode().values())
    vocab = base + [v + "</w>" for v in base] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
    enc = {v: i for i, v in enumerate(vocab)}
    tok = CLIPTokenizer(vocab=enc, merges=[tuple(m) for m in merges])
    golden = {"n_merges": len(merges), "sot": enc["<|startoftext|>"], "eot": enc["<|endoftext|>"], "cases": []}
    for t in TEXTS:
        ids = tok(t, add_special_tokens=True)["input_ids"]
        golden["cases"].append({"text": t, "ids": ids})
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json"), "w") as f:
        json.dump(golden, f, ensure_ascii=True, indent=0)
    print("merges", len(merges), "cases", len(TEXTS))


if __name__ == "__main__":
    main()
