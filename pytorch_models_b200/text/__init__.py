from .bert import BERT
from .generator import DecoderGenerator
from .gpt import GPT
from .gpt2 import GPT2

__all__ = ["BERT", "DecoderGenerator", "GPT", "GPT2"]
