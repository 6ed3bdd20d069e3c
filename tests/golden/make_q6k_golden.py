"""Golden vectors for the Q6_K dequantisation (ggml type 14): block bytes and the values the `gguf` package's
dequantiser (gguf.quants, an implementation of the public ggml block format independent of this repo) gives for them.

    python tests/golden/make_q6k_golden.py        # writes tests/golden/q6k_golden.npz

The blocks: 12 super-blocks quantised from normal data by the fixtures' quantiser, 12 of arbitrary bytes (every bit
pattern of ql / qh / scales) with finite super-scales.  tests/test_oracle.py checks the C oracle and the fixtures'
dequantiser against the stored values bit for bit, without needing the package.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from gguf import quants, GGMLQuantizationType as T
    from llm.f90_b200 import fixtures as fx
    rng = np.random.default_rng(2024)
    x = (rng.standard_normal((4, 768)) * 0.05).astype(np.float32)
    q = fx.quantize_q6_k(x)                                   # [4, 3 * 210]
    raw = rng.integers(0, 256, (4, 630), dtype=np.uint8)
    for b in range(3):
        raw[:, 210 * b + 208:210 * b + 210] = np.array([0.37 * (b + 1) * (-1) ** b], np.float16).view(np.uint8)
    blocks = np.concatenate([q, raw], axis=0)                 # [8, 630]
    values = quants.dequantize(blocks, T.Q6_K).astype(np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "q6k_golden.npz"), blocks=blocks, values=values)
    print("wrote q6k_golden.npz", blocks.shape, values.shape)


if __name__ == "__main__":
    main()
