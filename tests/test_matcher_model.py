"""CPU model of the arithmetic the tensor-core matcher (csrc/k_match.cu, k_match_imma) rests on, checked against the
oracle's brute-force popcount matcher:

  * with every descriptor bit written as +1 / -1 (s8), the dot product of two 256-bit descriptors is 256 - 2 * Hamming;
  * `expand4` (multiply-mask spread of four bits over four bytes, then 0/1 -> -1/+1) produces exactly those bytes;
  * the dot product may take its terms in any order, so the kernel's bit -> k mapping (lane (g, tig) owns halfwords
    tig, 4 + tig, 8 + tig, 12 + tig of a descriptor) is only a permutation applied to both operands;
  * the epilogue's single multiply-add  acc * -(2^21) + ((256 << 21) | index)  IS the packed key
    (distance << 22 | index) of the XOR / POPC kernel, because 256 - dot is even; min over keys = smallest distance,
    lowest index among equals.

CPU only; the kernel itself is checked on the GPU against the oracle (tests/test_gpu_parity.py, matcher tests)."""
import numpy as np


def expand4(nib: np.ndarray) -> np.ndarray:
    """the device function, on uint32"""
    w = (nib.astype(np.uint64) * 0x00204081) & 0x01010101
    return ((w | ((w ^ 0x01010101) * 0xFF)) & 0xFFFFFFFF).astype(np.uint32)


def test_expand4_is_plus_minus_one_per_bit():
    nib = np.arange(16, dtype=np.uint32)
    out = expand4(nib).view(np.int8).reshape(16, 4)
    want = np.array([[1 if (n >> i) & 1 else -1 for i in range(4)] for n in range(16)], np.int8)
    assert np.array_equal(out, want)


def _pm1(desc: np.ndarray) -> np.ndarray:
    bits = np.unpackbits(desc, axis=1, bitorder="little").astype(np.int32)
    return 2 * bits - 1


def test_dot_product_is_256_minus_twice_hamming_under_any_permutation():
    rng = np.random.default_rng(7)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (96, 32), dtype=np.uint8)
    ham = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2).astype(np.int32)
    perm = rng.permutation(256)  # the kernel's k order is one such permutation, the same for A and B
    dot = _pm1(q)[:, perm] @ _pm1(t)[:, perm].T
    assert np.array_equal(dot, 256 - 2 * ham)
    assert (dot % 2 == 0).all()


def test_packed_key_from_one_multiply_add_and_oracle_agreement(oracle):
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (50, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t[7] = q[3]; t[200] = q[3]          # an exact duplicate: ties must go to the lowest train index
    t[50] = q[10] ^ np.uint8(1)         # distance 1
    dot = (_pm1(q) @ _pm1(t).T).astype(np.int64)
    idx = np.arange(t.shape[0], dtype=np.int64)
    key = (dot * -(1 << 21) + ((256 << 21) | idx)[None, :]) & 0xFFFFFFFF  # the IMAD, modulo 2^32
    ham = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2).astype(np.int64)
    assert np.array_equal(key, (ham << 22) | idx[None, :])
    best = key.min(1)
    o_idx, o_dist, _ = oracle.match_knn(q, t, k=1)
    assert np.array_equal(best & 0x3FFFFF, np.asarray(o_idx).reshape(len(q), -1)[:, 0])
    assert np.array_equal(best >> 22, np.asarray(o_dist).reshape(len(q), -1)[:, 0])
    assert (best[3] & 0x3FFFFF) == 7 and (best[3] >> 22) == 0
