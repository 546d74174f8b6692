from .whisper import Whisper, WhisperDecoder, WhisperEncoder

__all__ = ["Whisper", "WhisperDecoder", "WhisperEncoder"]
