"""16-bit PCM WAV read / write with the standard library (the reference uses librosa.load and the
long-removed librosa.output.write_wav, generate.py:52-59,112-116; librosa is not a dependency here)."""
from __future__ import annotations

import wave

import numpy as np


def read_wav(path: str, sample_rate: int, offset=None, duration=None) -> np.ndarray:
    """Mono float32 in [-1, 1].  The file must already be at ``sample_rate`` (no resampler here)."""
    with wave.open(path, "rb") as w:
        sr, nch, sw, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        if sr != sample_rate:
            raise ValueError("{} is sampled at {} Hz, expected {} Hz (resample it first)".format(path, sr, sample_rate))
        raw = w.readframes(n)
    if sw == 2:
        x = np.frombuffer(raw, "<i2").astype(np.float32) / 32768.0
    elif sw == 1:
        x = (np.frombuffer(raw, np.uint8).astype(np.float32) - 128.0) / 128.0
    elif sw == 4:
        x = np.frombuffer(raw, "<i4").astype(np.float32) / 2147483648.0
    else:
        raise ValueError("unsupported sample width {}".format(sw))
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1)
    beg = int(round((offset or 0.0) * sr))
    end = len(x) if duration is None else min(len(x), beg + int(round(duration * sr)))
    return x[beg:end].astype(np.float32)


def write_wav(path: str, x: np.ndarray, sample_rate: int) -> None:
    pcm = np.clip(np.asarray(x, np.float64), -1.0, 1.0)
    pcm = np.round(pcm * 32767.0).astype("<i2")
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(pcm.tobytes())
