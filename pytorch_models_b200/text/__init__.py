from .bert import BERT
from .gpt import GPT
from .gpt2 import GPT2

__all__ = ["BERT", "GPT", "GPT2"]
