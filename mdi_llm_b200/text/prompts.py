"""Prompt styles (chat templates), their stop sequences, and user-prompt parsing.

Parity: reference ``src/sub/prompts.py`` — ``PromptStyle`` API (``apply``, ``stop_tokens``,
``from_name``, ``from_config``), the 24 styles, ``model_name_to_prompt_style`` (:325-366),
``save/load/has_prompt_style`` through ``prompt_style.yaml`` (:369-389) and
``get_user_prompt`` with ``FILE:`` paragraph parsing (:392-447).

Implementation is table-driven: a style is a template plus a stop-sequence spec, and a class
per style is synthesised so that ``prompt_style.yaml`` files keep a resolvable ``class_path``
(class paths written by the reference, e.g. ``sub.prompts.Llama3``, resolve by class name).
"""
from __future__ import annotations

import importlib
import json
import re
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Type, Union

import yaml

from ..models.config import Config

StopSpec = Sequence[Sequence[Union[int, str]]]  # each entry: ids (int) or token strings (str)


class PromptStyle:
    """Base interface.  ``template`` uses ``{prompt}``; ``stops`` lists extra stop sequences
    (besides EOS) whose items are token strings (resolved via the tokenizer) or raw ids."""

    template: str = "{prompt}"
    stops: StopSpec = ()

    def apply(self, prompt: str, **kwargs: str) -> str:
        return self.template.replace("{prompt}", prompt)

    def stop_tokens(self, tokenizer: Any) -> Tuple[List[int], ...]:
        out: List[List[int]] = [[tokenizer.eos_id]]
        for seq in self.stops:
            out.append([t if isinstance(t, int) else tokenizer.token_to_id(t) for t in seq])
        return tuple(out)

    @classmethod
    def from_name(cls, name: str) -> "PromptStyle":
        return prompt_styles[name]()

    @classmethod
    def from_config(cls, config: Config) -> "PromptStyle":
        return model_name_to_prompt_style(config.name)


class Default(PromptStyle):
    pass


class NoPrompt(PromptStyle):
    """Ignore the user prompt: generation starts from a single newline."""

    def apply(self, prompt: str, **kwargs: str) -> str:
        return "\n"


_ALPACA_HEAD = "Below is an instruction that describes a task"
_ALPACA_TAIL = "Write a response that appropriately completes the request.\n\n"


class Alpaca(PromptStyle):
    def apply(self, prompt: str, **kwargs: str) -> str:
        if kwargs.get("input"):
            return (f"{_ALPACA_HEAD}, paired with an input that provides further context. {_ALPACA_TAIL}"
                    f"### Instruction:\n{prompt}\n\n### Input:\n{kwargs['input']}\n\n### Response:\n")
        return f"{_ALPACA_HEAD}. {_ALPACA_TAIL}### Instruction:\n{prompt}\n\n### Response:\n"


class Llama2FunctionCalling(PromptStyle):
    """Llama-2 chat wrapper preceded by a JSON function list (example: a Bing search tool)."""

    FUNCTION = {
        "function": "search_bing",
        "description": ("Search the web for content on Bing. This allows users to search online/the "
                        "internet/the web for content."),
        "arguments": [{"name": "query", "type": "string", "description": "The search query string"}],
    }
    SYSTEM = ("You are a helpful, respectful and honest assistant. Always answer as helpfully as"
              "possible. Your only response should be JSON formatted functions")

    def apply(self, prompt: str, **kwargs: str) -> str:
        funcs = json.dumps(self.FUNCTION).replace("{", "{{").replace("}", "}}").strip()
        return (f"<FUNCTIONS>{funcs}</FUNCTIONS>\n\n[INST]<<SYS>>\n{self.SYSTEM.strip()}"
                f"\n<</SYS>>\n\n{prompt}[/INST]\n\n")


_LLAMA2_SYSTEM = (
    "You are a helpful, respectful and honest assistant. Always answer as helpfully as"
    " possible, while being safe.  Your answers should not include any harmful, unethical, racist, sexist,"
    " toxic, dangerous, or illegal content. Please ensure that your responses are socially unbiased and"
    " positive in nature.\n\nIf a question does not make any sense, or is not factually coherent, explain why"
    " instead of answering something not correct. If you don't know the answer to a question, please don't"
    " share false information."
)
_STABLELM_SYSTEM = (
    "<|SYSTEM|># StableLM Tuned (Alpha version)\n- StableLM is a helpful and harmless open-source AI language"
    " model developed by StabilityAI.\n- StableLM is excited to be able to help the user, but will refuse to do"
    " anything that could be considered harmful to the user.\n- StableLM is more than just an information"
    " source, StableLM is also able to write poetry, short stories, and make jokes.\n- StableLM will refuse to"
    " participate in anything that could harm a human."
)

# name -> (class name, template, extra stop sequences)
_TABLE: Dict[str, Tuple[str, str, StopSpec]] = {
    "flan": ("FLAN", f"{_ALPACA_HEAD}. {_ALPACA_TAIL}### Instruction:\n{{prompt}}\n\n### Response:\n", ()),
    "longform": ("Longform", f"{_ALPACA_HEAD}, paired with an input that provides further context. "
                             f"{_ALPACA_TAIL}### Instruction:\n{{prompt}}\n\n### Response:\n", ()),
    "stablelm-alpha": ("StableLMAlpha", _STABLELM_SYSTEM + "<|USER|>{prompt}<|ASSISTANT|>",
                       (["<|SYSTEM|>"], ["<|ASSISTANT|>"], ["<|USER|>"])),
    "stablelm-zephyr": ("StableLMZephyr", "<|user|>\n{prompt}<|endoftext|>\n<|assistant|>\n", ()),
    "togethercomputer-chat": ("TogetherComputerChat", "<human>: {prompt}\n<bot>:",
                              (["<", "human", ">:"], ["<", "bot", ">:"])),
    "togethercomputer-instruct": ("TogetherComputerInstruct", "Q: {prompt}\nA:",
                                  (["Q", ":"], ["Question"], ["A", ":"], ["Label", ":"],
                                   [187, 187], [535], [2756])),
    "falcon": ("Falcon", "Do not prefix your replies with 'Bot: '\nUser: {prompt}\n",
               (["User", ":"], [193, "User"])),
    "vicuna": ("Vicuna", "A chat between a curious user and an artificial intelligence assistant. "
                         "The assistant gives helpful, detailed, and polite answers to the user's "
                         "questions. USER: {prompt} ASSISTANT:", ()),
    "llama2": ("Llama2", "[INST] <<SYS>>\n" + _LLAMA2_SYSTEM + "\n<</SYS>>\n\n {prompt} [/INST] ", ()),
    "llama3": ("Llama3", "<|begin_of_text|><|start_header_id|>system<|end_header_id|>\n\n"
                         "You are a helpful assistant.<|eot_id|>\n"
                         "<|start_header_id|>user<|end_header_id|>\n\n{prompt}<|eot_id|>\n"
                         "<|start_header_id|>assistant<|end_header_id|>\n\n", (["<|eot_id|>"],)),
    "freewilly2": ("FreeWilly2", "### System:\nThis is a system prompt, please behave and help the user."
                                 "\n\n### User:\n{prompt}\n\n### Assistant:\n", ()),
    "platypus": ("Platypus", "### Instruction:\n\n{prompt}\n\n### Response:\n", ()),
    "nous-research": ("NousResearch", "### Instruction:\n{prompt}\n\n### Response:\n", ()),
    "stablecode": ("StableCode", "###Instruction\n{prompt}###Response\n", ()),
    "codellama": ("CodeLlama", "<s>[INST] {prompt} [/INST]", ()),
    "phi-1": ("Phi1", "{prompt}\n\nAnswer:", (["Answer", ":"], [198, "Answer", ":"])),
    "phi-2": ("Phi2", "Instruct: {prompt}\nOutput:", ()),
    "tinyllama": ("TinyLlama", "<|system|>\nYou are a friendly chatbot who always gives helpful, "
                               "detailed, and polite answers.</s>\n<|user|>\n{prompt}</s>\n<|assistant|>\n", ()),
    "gemma": ("Gemma", "<start_of_turn>user\n{prompt}<end_of_turn>\n<start_of_turn>model\n", ()),
    "h2oai": ("H2Oai", "<|prompt|>{prompt}</s><|answer|>", ()),
}

prompt_styles: Dict[str, Type[PromptStyle]] = {
    "default": Default, "no-prompt": NoPrompt, "alpaca": Alpaca,
    "llama2-function-calling": Llama2FunctionCalling,
}
for _name, (_cls_name, _tmpl, _stops) in _TABLE.items():
    _cls = type(_cls_name, (PromptStyle,), {"template": _tmpl, "stops": _stops, "__module__": __name__})
    globals()[_cls_name] = _cls
    prompt_styles[_name] = _cls

# first matching pattern wins — order mirrors the reference's if-chain (prompts.py:325-366)
_NAME_RULES: List[Tuple[str, str]] = [
    (r"stablelm-tuned-alpha", "stablelm-alpha"), (r"stablelm-zephyr-3b", "stablelm-zephyr"),
    (r"stablecode-instruct", "stablecode"), (r"RedPajama-INCITE.*-Chat", "togethercomputer-chat"),
    (r"RedPajama-INCITE.*-Instruct", "togethercomputer-instruct"), (r"falcon.*-instruct", "falcon"),
    (r"vicuna|longchat", "vicuna"),
    (r"Llama-2-7b-chat-hf-function-calling-v2", "llama2-function-calling"),
    (r"Llama-2.*-chat*", "llama2"), (r"Llama-3.*-Instruct", "llama3"), (r"FreeWilly2", "freewilly2"),
    (r"Platypus", "platypus"), (r"Nous-Hermes", "nous-research"),
    (r"CodeLlama|Mistral.*Instruct", "codellama"), (r"phi-1", "phi-1"), (r"phi-2", "phi-2"),
    (r"tiny-llama.*chat", "tinyllama"), (r"(Code)?Gemma.*-it", "gemma"), (r"Danube2.*-chat", "h2oai"),
]


def model_name_to_prompt_style(model_name: str) -> PromptStyle:
    for pattern, style in _NAME_RULES:
        if re.search(pattern, model_name):
            return prompt_styles[style]()
    if re.search(r"nanollama*", model_name.lower()):
        return NoPrompt()
    return Default()


def save_prompt_style(style: Union[str, PromptStyle], checkpoint_dir: Union[str, Path]) -> None:
    style = PromptStyle.from_name(style) if isinstance(style, str) else style
    cls = type(style)
    with open(Path(checkpoint_dir) / "prompt_style.yaml", "w", encoding="utf-8") as fp:
        yaml.dump({"class_path": f"{cls.__module__}.{cls.__name__}"}, fp)


def load_prompt_style(checkpoint_dir: Union[str, Path]) -> PromptStyle:
    with open(Path(checkpoint_dir) / "prompt_style.yaml", "r", encoding="utf-8") as fp:
        class_path = yaml.safe_load(fp)["class_path"]
    module_path, cls_name = class_path.rsplit(".", 1)
    try:
        return getattr(importlib.import_module(module_path), cls_name)()
    except (ImportError, AttributeError):
        # files written by the reference / litGPT name their own modules: resolve by class name
        if cls_name in globals() and isinstance(globals()[cls_name], type):
            return globals()[cls_name]()
        raise


def has_prompt_style(checkpoint_dir: Union[str, Path]) -> bool:
    return (Path(checkpoint_dir) / "prompt_style.yaml").is_file()


def read_prompt_file(path: Union[str, Path], limit: Optional[int] = None) -> List[str]:
    """Paragraphs (blank-line separated, newlines kept) of a prompt file."""
    out: List[str] = []
    cur = ""
    with open(path, "r", encoding="utf-8") as fp:
        for line in fp:
            if line.strip():
                cur += line
            else:
                # the reference closes a paragraph on every blank line, even consecutive ones
                out.append(cur)
                cur = ""
            if limit is not None and len(out) == limit:
                break
    if cur:
        out.append(cur)
    return out


def get_user_prompt(prompt: str, n_samples: int = 1, prompt_style: Optional[PromptStyle] = None,
                    **kwargs: Any) -> List[str]:
    """Expand the CLI ``--prompt`` into ``n_samples`` styled prompts.

    A literal string is repeated for every sample; ``FILE:<path>`` (``.txt/.md/.tex``) yields
    one paragraph per sample, truncated to ``n_samples`` or padded with ``"\\n"``.
    """
    style = prompt_style if prompt_style is not None else NoPrompt()
    if isinstance(style, type):
        style = style()
    if not prompt.startswith("FILE:"):
        return [style.apply(prompt)] * n_samples
    if not prompt.endswith((".txt", ".md", ".tex")):
        raise ValueError(f"Unsupported file type for {prompt}\nSupported types are: '.txt', '.md', '.tex'")
    paragraphs = read_prompt_file(prompt[5:], limit=n_samples)[:n_samples]
    paragraphs += ["\n"] * (n_samples - len(paragraphs))
    out = [style.apply(p) for p in paragraphs]
    if kwargs.get("verb"):
        print(out)
    return out


def get_prompt(prompt: str, n_samples: int = 1, **kwargs: Any) -> List[str]:
    """Second-generation prompt expansion: no prompt style, the text is used verbatim
    (reference ``old/GPT2/sub/utils.py:478-531``)."""
    return get_user_prompt(prompt, n_samples, Default(), **kwargs)
