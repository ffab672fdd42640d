"""Digits dataset of the reference's golden tests, rebuilt from the committed
bit-packed fixture (oracle/test infrastructure only).

Mirrors the dataset.matrix definitions in TEST/digitos/test.lua:31-67:
  patternSize {16,16}, stepSize {16,16}, numSteps {80,10} / {20,10},
  orderStep {1,0}  -> pattern i is the 16x16 block at row-block i//10,
  column-block i%10, flattened row-major; its class is i%10 (the circular
  one-hot dataset with stepSize -1).
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(_HERE, "..", "tests", "golden", "digits_bits.npy")


def load_digits():
    bits = np.load(FIXTURE)
    img = np.unpackbits(bits, axis=1).astype(np.float32)  # [1600,160], 1 = ink
    blocks = img.reshape(100, 16, 10, 16).transpose(0, 2, 1, 3).reshape(1000, 256)
    labels = np.arange(1000) % 10
    onehot = np.zeros((1000, 10), dtype=np.float32)
    onehot[np.arange(1000), labels] = 1.0
    return (np.ascontiguousarray(blocks[:800]), onehot[:800],
            np.ascontiguousarray(blocks[800:]), onehot[800:])
