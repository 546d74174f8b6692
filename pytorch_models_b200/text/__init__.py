from .bert import BERT

__all__ = ["BERT"]
