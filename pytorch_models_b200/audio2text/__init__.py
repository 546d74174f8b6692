from .whisper import Whisper, WhisperDecoder, WhisperEncoder, WhisperPreprocessor

__all__ = ["Whisper", "WhisperDecoder", "WhisperEncoder", "WhisperPreprocessor"]
