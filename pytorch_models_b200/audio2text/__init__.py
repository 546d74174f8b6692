from .whisper import WhisperEncoder

__all__ = ["WhisperEncoder"]
