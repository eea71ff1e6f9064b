"""Synthetic frame generation for tests and benchmarks (numpy, host side).

Mirrors what the reference harness feeds the decoder (main.cpp:450-538): arbitrary
random bit vectors x (not codewords), target syndrome s = H x, BPSK 1 -> +1 / 0 -> -1
(common.h:51-59) through an AWGN or BSC channel, punctured tail = 0.  The random
streams are numpy's, not the reference's ChaCha8 (that replay is SURVEY.md row f1).
"""
from __future__ import annotations

import numpy as np

from .alist import Code


def make_frames(code: Code, n_frames: int, channel: str, noise: float, seed: int = 0, dtype=np.float32):
    """Returns (inputs [N, n_frames] dtype, syndromes u32[n_frames, S], bits_packed u32[n_frames, N/32]).

    inputs follows the reference API layout p_input[v + n_frames*i] (ldpc_decoder_gpu.h:93-94)."""
    rng = np.random.default_rng(seed)
    N = code.n_vars
    n_tx = N - code.n_erased
    bits_packed = rng.integers(0, 2**32, size=(n_frames, N // 32), dtype=np.uint32)
    syndromes = code.syndrome_packed(bits_packed)
    inputs = np.zeros((N, n_frames), dtype=dtype)
    # generate in frame-major chunks to bound memory, then transpose into the API layout
    step = max(1, (1 << 24) // max(N, 1))
    for f0 in range(0, n_frames, step):
        f1 = min(n_frames, f0 + step)
        bits = np.unpackbits(bits_packed[f0:f1].view(np.uint8), axis=1, bitorder="little")[:, :n_tx]
        x = bits.astype(np.float32) * 2.0 - 1.0
        if channel == "awgn":
            x += rng.standard_normal(x.shape, dtype=np.float32) * np.float32(noise)
        elif channel == "bsc":
            flip = rng.random(x.shape, dtype=np.float32) < np.float32(noise)
            x = np.where(flip, -x, x)
        else:
            raise ValueError(channel)
        inputs[:n_tx, f0:f1] = x.T.astype(dtype)
    return inputs, syndromes, bits_packed


def channel_param(channel: str, noise: float) -> float:
    """BSC: ref_llr = ln((1-p)/p) (channel.cpp:6-9); AWGN: factor = 2/sigma^2 (channel.h:70-73)."""
    if channel == "bsc":
        return float(np.float32(np.log(np.float32(1.0) - np.float32(noise)) - np.log(np.float32(noise))))
    s = np.float32(noise)
    snr = np.float32(1.0) / (s * s)
    return float(np.float32(2.0) * snr)


def count_bit_errors(results: np.ndarray, bits_packed: np.ndarray) -> np.ndarray:
    x = np.bitwise_xor(results, bits_packed)
    return np.unpackbits(x.view(np.uint8), axis=1).sum(axis=1)
