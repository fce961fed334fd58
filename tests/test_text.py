import torch
import pytest

from mdi_llm_b200.text import prompts as P
from mdi_llm_b200.text.simple_tokenizers import BPETokenizer, CharacterTokenizer
from mdi_llm_b200.text.tokenizer import Tokenizer, write_bytes_tokenizer
from mdi_llm_b200.utils.misc import detect_stop_tokens, find_eot, get_lr, loading_bar


class _Tok:
    eos_id = 2

    def token_to_id(self, t):
        return {"<|eot_id|>": 128009}.get(t, 7)


def test_llama3_style_and_stops():
    s = P.model_name_to_prompt_style("Llama-3-8B-Instruct")
    assert type(s).__name__ == "Llama3"
    text = s.apply("Who are you?")
    assert text.startswith("<|begin_of_text|><|start_header_id|>system<|end_header_id|>")
    assert text.endswith("<|start_header_id|>assistant<|end_header_id|>\n\n") and "Who are you?<|eot_id|>" in text
    assert s.stop_tokens(_Tok()) == ([2], [128009])


@pytest.mark.parametrize("name,cls", [("tiny-llama-1.1b-chat", "TinyLlama"), ("Llama-2-7b-chat-hf", "Llama2"),
                                      ("CodeLlama-7b-Instruct-hf", "CodeLlama"), ("NanoLlama", "NoPrompt"),
                                      ("pythia-14m", "Default"), ("Gemma-7b-it", "Gemma"),
                                      ("Llama-2-7b-chat-hf-function-calling-v2", "Llama2FunctionCalling")])
def test_name_to_style(name, cls):
    assert type(P.model_name_to_prompt_style(name)).__name__ == cls


def test_prompt_style_yaml_roundtrip_and_reference_class_path(tmp_path):
    P.save_prompt_style("llama3", tmp_path)
    assert P.has_prompt_style(tmp_path) and type(P.load_prompt_style(tmp_path)).__name__ == "Llama3"
    (tmp_path / "prompt_style.yaml").write_text("class_path: sub.prompts.TinyLlama\n")
    assert type(P.load_prompt_style(tmp_path)).__name__ == "TinyLlama"


def test_get_user_prompt_file_paragraphs(tmp_path):
    f = tmp_path / "p.txt"
    f.write_text("first line\nsecond line\n\npara two\n\npara three\n")
    out = P.get_user_prompt(f"FILE:{f}", 2, P.Default())
    assert out == ["first line\nsecond line\n", "para two\n"]
    out = P.get_user_prompt(f"FILE:{f}", 5, P.Default())
    assert out[2] == "para three\n" and out[3:] == ["\n", "\n"]
    assert P.get_user_prompt("hello", 3, P.NoPrompt()) == ["\n"] * 3
    with pytest.raises(ValueError):
        P.get_user_prompt("FILE:x.pdf", 1)


def test_byte_tokenizer_backend(tmp_path):
    write_bytes_tokenizer(tmp_path)
    t = Tokenizer(tmp_path)
    ids = t.encode("héllo", eos=True)
    assert ids.dtype == torch.int32 and ids[0] == t.bos_id and ids[-1] == t.eos_id
    assert t.decode(ids) == "héllo" and t.vocab_size == 259
    assert t.encode("abcdef", bos=False, max_length=3).tolist() == [97, 98, 99]
    with pytest.raises(FileNotFoundError):
        Tokenizer(tmp_path, force_backend="sentencepiece")


def test_hf_backend_from_generated_tokenizer_json(tmp_path):
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import models, pre_tokenizers

    vocab = {"<s>": 0, "</s>": 1, "hello": 2, "world": 3, "[UNK]": 4}
    tk = tokenizers.Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tk.pre_tokenizer = pre_tokenizers.Whitespace()
    tk.save(str(tmp_path / "tokenizer.json"))
    (tmp_path / "tokenizer_config.json").write_text('{"bos_token": "<s>", "eos_token": "</s>", "add_bos_token": true}')
    t = Tokenizer(tmp_path)
    assert t.backend == "huggingface" and (t.bos_id, t.eos_id) == (0, 1) and t.use_bos
    assert t.encode("hello world").tolist() == [0, 2, 3]
    assert t.token_to_id("world") == 3
    with pytest.raises(ValueError):
        t.token_to_id("absent")


def test_char_and_bpe_tokenizers_roundtrip(tmp_path):
    text = "the quick brown fox jumps over the lazy dog. " * 30
    c = CharacterTokenizer()
    c.tokenize(text)
    assert c.decode(c.encode("lazy fox")) == "lazy fox"
    c.save(tmp_path / "c")
    assert Tokenizer(tmp_path / "c").backend == "char"
    b = BPETokenizer()
    b.tokenize(text, 300)
    assert b.trained() and 256 < b.vocab_size <= 300
    ids = b.encode("the lazy dog jumps")
    assert b.decode(ids) == "the lazy dog jumps" and len(ids) < len("the lazy dog jumps")
    b.store_tokenizer_info(tmp_path / "b")
    b2 = BPETokenizer()
    b2.load_tokenizer_info(tmp_path / "b")
    assert b2.encode("quick brown") == b.encode("quick brown")
    assert Tokenizer(tmp_path / "b").backend == "bpe"


def test_misc_helpers():
    toks = torch.tensor([[1, 2, 3, 9, 4, 5, 9, 7]])
    assert find_eot(toks, ([9],), 3).tolist() == [[1, 2, 3, 9]]
    assert find_eot(toks, ([5, 9],), 0).tolist() == [[1, 2, 3, 9, 4, 5, 9]]
    assert find_eot(toks, ([42],), 0) is toks
    assert detect_stop_tokens(toks[:, :7], ([5, 9],)) and not detect_stop_tokens(toks, ([5, 9],))
    assert get_lr(0) == 0 and abs(get_lr(2000) - 3e-4) < 1e-12 and get_lr(10 ** 7) == 6e-5
    assert loading_bar(5, 10, 10) == "[=====    ]"


def test_legacy_get_prompt_is_verbatim(tmp_path):
    from mdi_llm_b200.text.prompts import get_prompt

    assert get_prompt("hi", 2) == ["hi", "hi"]
    f = tmp_path / "p.txt"
    f.write_text("first\nparagraph\n\nsecond\n")
    assert get_prompt(f"FILE:{f}", 3) == ["first\nparagraph\n", "second\n", "\n"]
