"""Audio front-end for speech models: container decoding, resampling and Whisper-style log-mel features.

Behavioural reference: vllm/multimodal/audio.py (resampling helpers), vllm/entrypoints/speech_to_text/* (30-second
chunking) and the HF ``WhisperFeatureExtractor`` the reference delegates to (n_fft=400, hop=160, slaney mel bank,
log10 → clamp to max-8 → (x+4)/4).  Implemented in torch so the STFT runs on the device when one is present.
"""
from __future__ import annotations

import io
import math
import wave
from typing import Any, Union

import numpy as np
import torch

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
CHUNK_S = 30


def load_audio(src: Union[str, bytes, io.IOBase, tuple, np.ndarray, torch.Tensor], target_sr: int = SAMPLE_RATE) -> np.ndarray:
    """→ mono float32 waveform at ``target_sr``.  Accepts a path / bytes / file object holding RIFF-WAV (PCM 8/16/24/32
    or float32), a ``(waveform, sample_rate)`` tuple, or a raw array already at the target rate."""
    if isinstance(src, tuple):
        wav, sr = src
        wav = np.asarray(wav, dtype=np.float32)
    elif isinstance(src, (np.ndarray, torch.Tensor, list)):
        return np.asarray(src, dtype=np.float32).reshape(-1)
    else:
        if isinstance(src, (bytes, bytearray)):
            src = io.BytesIO(src)
        wav, sr = _read_wav(src)
    if wav.ndim > 1:
        wav = wav.mean(axis=-1 if wav.shape[-1] <= 8 else 0)
    if sr != target_sr:
        wav = resample(wav, sr, target_sr)
    return np.ascontiguousarray(wav, dtype=np.float32)


def _read_wav(f: Any) -> tuple[np.ndarray, int]:
    try:
        with wave.open(f, "rb") as w:
            sr, ch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
            raw = w.readframes(n)
    except wave.Error:
        # IEEE-float WAVs (format tag 3) are rejected by the stdlib reader: parse the chunks by hand
        if hasattr(f, "seek"):
            f.seek(0)
            data = f.read()
        else:
            with open(f, "rb") as fh:
                data = fh.read()
        return _read_wav_float(data)
    if width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = v.astype(np.float32) / float(1 << 23)
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / float(1 << 31)
    else:
        raise ValueError(f"unsupported WAV sample width {width}")
    return (x.reshape(-1, ch) if ch > 1 else x), sr


def _read_wav_float(data: bytes) -> tuple[np.ndarray, int]:
    import struct
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("unsupported audio container (only RIFF/WAVE is decoded natively)")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("malformed WAV file")
    tag, ch, sr, _, _, bits = fmt
    if tag == 3 and bits == 32:
        x = np.frombuffer(pcm, dtype="<f4").astype(np.float32)
    elif tag == 3 and bits == 64:
        x = np.frombuffer(pcm, dtype="<f8").astype(np.float32)
    else:
        raise ValueError(f"unsupported WAV format tag {tag} / {bits} bits")
    return (x.reshape(-1, ch) if ch > 1 else x), sr


def resample(wav: np.ndarray, sr: int, target_sr: int) -> np.ndarray:
    if sr == target_sr:
        return wav
    from scipy.signal import resample_poly
    g = math.gcd(int(sr), int(target_sr))
    return resample_poly(wav, target_sr // g, sr // g).astype(np.float32)


def _hz_to_mel(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    log_region = f >= 1000.0
    mel = np.where(log_region, 15.0 + np.log(np.maximum(f, 1e-10) / 1000.0) * (27.0 / np.log(6.4)), mel)
    return mel


def _mel_to_hz(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    log_region = m >= 15.0
    return np.where(log_region, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), f)


def mel_filter_bank(n_mels: int, n_fft: int = N_FFT, sr: int = SAMPLE_RATE) -> torch.Tensor:
    """[n_fft/2+1, n_mels] slaney-scale, slaney-normalised triangular filters (what Whisper was trained with)."""
    n_freq = n_fft // 2 + 1
    fft_freqs = np.linspace(0, sr / 2, n_freq)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    hz_pts = _mel_to_hz(mel_pts)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (hz_pts[2:n_mels + 2] - hz_pts[:n_mels]))[None, :]
    return torch.from_numpy(fb.astype(np.float32))


_FB_CACHE: dict[tuple[int, str], torch.Tensor] = {}


def log_mel_spectrogram(wav: Union[np.ndarray, torch.Tensor], n_mels: int = 80, chunk_s: int = CHUNK_S,
                        device: Union[str, torch.device] = "cpu") -> torch.Tensor:
    """One 30-second window → float32 [n_mels, chunk_s*100].  Shorter input is zero-padded, longer input is cut
    (callers that transcribe long audio split it with :func:`split_chunks` first)."""
    x = torch.as_tensor(wav, dtype=torch.float32, device=device).reshape(-1)
    n = chunk_s * SAMPLE_RATE
    x = x[:n] if x.numel() >= n else torch.nn.functional.pad(x, (0, n - x.numel()))
    win = torch.hann_window(N_FFT, periodic=True, device=x.device)
    spec = torch.stft(x, N_FFT, HOP, window=win, center=True, pad_mode="reflect", return_complex=True)
    power = spec[:, :-1].abs().pow(2)                                    # [201, 3000]
    key = (n_mels, str(x.device))
    fb = _FB_CACHE.get(key)
    if fb is None:
        fb = _FB_CACHE[key] = mel_filter_bank(n_mels).to(x.device)
    mel = fb.T @ power
    log = torch.clamp(mel, min=1e-10).log10()
    log = torch.maximum(log, log.max() - 8.0)
    return (log + 4.0) / 4.0


def split_chunks(wav: np.ndarray, chunk_s: int = CHUNK_S, overlap_s: float = 1.0, min_energy_window_s: float = 0.1) -> list[np.ndarray]:
    """Cuts long audio into ≤ ``chunk_s`` pieces, preferring the quietest point of the last ``overlap_s`` seconds of
    each window as the cut (reference: speech_to_text ``_split_audio`` / ``_find_split_point``)."""
    n = chunk_s * SAMPLE_RATE
    if wav.shape[0] <= n:
        return [wav]
    out, pos = [], 0
    ov, win = int(overlap_s * SAMPLE_RATE), max(int(min_energy_window_s * SAMPLE_RATE), 1)
    while pos < wav.shape[0]:
        end = min(pos + n, wav.shape[0])
        if end < wav.shape[0]:
            seg = wav[end - ov:end]
            k = seg.shape[0] // win
            if k > 0:
                e = (seg[:k * win].reshape(k, win) ** 2).sum(axis=1)
                end = end - ov + int(e.argmin()) * win + win // 2
        out.append(wav[pos:end])
        pos = end
    return out
