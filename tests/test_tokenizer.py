"""Native tokenizer of the serve host's text path against the HF `tokenizers` library, the
tokenizer the reference's serving images use: ids and decoded text must be identical for the two supported families —
Llama-2 style (SentencePiece-like BPE, Prepend/Replace normalizer, byte fallback, BOS template) and GPT-2/OPT style
(byte-level BPE with the GPT-2 split pattern).  No tokenizer files exist offline, so both are trained here with the
same component configuration the real checkpoints ship.  Unsupported components must be refused at load."""
import json
import random

import pytest

tokenizers = pytest.importorskip("tokenizers")


def _corpus():
    random.seed(0)
    words = ("the quick brown fox jumps over lazy dog Hello world Substratus serves models on Kubernetes with GPUs émigré "
             "naïve 你好 世界 token 12345 6789 foo_bar baz() {} [] :: -> == != <= >= tensor parallel decode prefill llama falcon").split()
    return [" ".join(random.choice(words) for _ in range(random.randint(3, 15))) for _ in range(2000)]


def _llama_like(path):
    from tokenizers import Tokenizer, decoders, models, normalizers, processors, trainers

    tok = Tokenizer(models.BPE(unk_token="<unk>", byte_fallback=True, fuse_unk=True))
    tok.normalizer = normalizers.Sequence([normalizers.Prepend("▁"), normalizers.Replace(" ", "▁")])
    tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.ByteFallback(), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    tr = trainers.BpeTrainer(vocab_size=700, special_tokens=["<unk>", "<s>", "</s>"] + [f"<0x{i:02X}>" for i in range(256)],
                             show_progress=False)
    tok.train_from_iterator(_corpus(), tr)
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", pair="<s> $A <s> $B", special_tokens=[("<s>", tok.token_to_id("<s>"))])
    tok.save(path)
    return tok


def _opt_like(path):
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors, trainers

    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True)
    tok.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=600, special_tokens=["<pad>", "</s>"], initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                             show_progress=False)
    tok.train_from_iterator(_corpus(), tr)
    tok.post_processor = processors.TemplateProcessing(single="</s> $A", special_tokens=[("</s>", tok.token_to_id("</s>"))])
    tok.save(path)
    return tok


def _falcon_like(path):
    """Falcon-40B's tokenizer.json: Sequence[Punctuation(Contiguous), ByteLevel(no prefix space), Digits, Split(3 digits)]"""
    from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers, trainers

    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Punctuation("contiguous"), pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=True),
        pre_tokenizers.Digits(individual_digits=False), pre_tokenizers.Split(Regex("[0-9][0-9][0-9]"), "isolated")])
    tok.decoder = decoders.ByteLevel()
    tr = trainers.BpeTrainer(vocab_size=600, special_tokens=[">>TITLE<<", "<|endoftext|>"],
                             initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), show_progress=False)
    tok.train_from_iterator(_corpus() + ["12345 678 90", "3.14159, 2024-09-21!!! (x) [y]... ¿qué? «hi» — ²³¼"], tr)
    tok.save(path)
    return tok


TEXTS = ["Hello world, naïve 你好!", "", " ", "  leading and  double  spaces ", "it's we're I'll don't 'quoted'",
         "tabs\tand\nnewlines\n\nx", "12345 67 8.9e-3 $100 €5 — “quotes” … emoji 🙂 ok", "<s>special</s> inside",
         "foo_bar baz() {} [] :: -> == != <= >=", "ÀÉÎÕÜ çñß Ω мир שלום مرحبا", "Who was the first president of the United States?"]


@pytest.mark.parametrize("family", ["llama", "opt", "falcon"])
def test_native_tokenizer_matches_hf_tokenizers(tmp_path, family, lib):
    from substratus_b200.engine import NativeTokenizer

    path = str(tmp_path / "tokenizer.json")
    ref = {"llama": _llama_like, "opt": _opt_like, "falcon": _falcon_like}[family](path)
    nat = NativeTokenizer(path)
    random.seed(1)
    alphabet = "abc XYZ 0123456789 .,'!?()-\t\né你🙂_-²¼٣"
    texts = TEXTS + ["".join(random.choice(alphabet) for _ in range(random.randint(1, 40))) for _ in range(400)]
    for t in texts:
        want = ref.encode(t).ids
        assert nat.encode(t) == want, repr(t)
        assert nat.decode(want) == ref.decode(want), repr(t)
        assert nat.encode(t, add_special=False) == ref.encode(t, add_special_tokens=False).ids, repr(t)
    rnd = [random.randrange(ref.get_vocab_size()) for _ in range(200)]  # arbitrary id streams (ill-formed byte runs included)
    if family == "llama":
        assert nat.decode(rnd) == ref.decode(rnd)


@pytest.mark.parametrize("family", ["opt", "falcon"])
def test_byte_level_splits_agree_on_every_unicode_plane(tmp_path, family, lib):
    """The character classes behind the GPT-2 pattern / Punctuation / Digits come from tables generated from the
    `tokenizers` library's own behaviour (tools/gen_unicode_tables.py): ids must agree on text drawn from all of
    Unicode, not just the scripts of the training corpus."""
    from substratus_b200.engine import NativeTokenizer

    path = str(tmp_path / "tokenizer.json")
    ref = (_opt_like if family == "opt" else _falcon_like)(path)
    nat = NativeTokenizer(path)
    rng = random.Random(5)

    def cp():
        r = rng.random()
        c = rng.randrange(0x20, 0x7F) if r < 0.35 else rng.randrange(0x80, 0x3000) if r < 0.7 else rng.randrange(0x3000, 0x110000)
        return " " if 0xD800 <= c <= 0xDFFF else chr(c)

    for _ in range(3000):
        t = "".join(cp() for _ in range(rng.randint(1, 24)))
        want = ref.encode(t).ids
        assert nat.encode(t) == want, [hex(ord(c)) for c in t]
        assert nat.decode(want) == ref.decode(want)


def test_unsupported_tokenizer_is_refused(tmp_path, lib):
    from substratus_b200 import SsbError
    from substratus_b200.engine import NativeTokenizer

    p = tmp_path / "tokenizer.json"
    p.write_text(json.dumps({"model": {"type": "WordPiece", "vocab": {}}}))
    with pytest.raises(SsbError):
        NativeTokenizer(str(p))
    p.write_text(json.dumps({"model": {"type": "BPE", "vocab": {"a": 0}, "merges": []}, "normalizer": {"type": "NFKC"},
                             "pre_tokenizer": None}))
    with pytest.raises(SsbError) as ei:
        NativeTokenizer(str(p))
    assert "normalizer" in str(ei.value)
    with pytest.raises(SsbError):
        NativeTokenizer(str(tmp_path / "missing.json"))
