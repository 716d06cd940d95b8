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
    """the device function, on uint32: spread the four bits over the bytes' LSBs, then per byte 255 - 254 * bit"""
    w = (nib.astype(np.uint64) * 0x00204081) & 0x01010101
    return ((w * 0xFFFFFF02 + 0xFFFFFFFF) & 0xFFFFFFFF).astype(np.uint32)


def expand4_first_form(nib: np.ndarray) -> np.ndarray:
    """the five-instruction form the kernels used before (w | (w ^ 0x01010101) * 0xFF)"""
    w = (nib.astype(np.uint64) * 0x00204081) & 0x01010101
    return ((w | ((w ^ 0x01010101) * 0xFF)) & 0xFFFFFFFF).astype(np.uint32)


def test_expand4_is_plus_minus_one_per_bit():
    nib = np.arange(16, dtype=np.uint32)
    out = expand4(nib).view(np.int8).reshape(16, 4)
    want = np.array([[1 if (n >> i) & 1 else -1 for i in range(4)] for n in range(16)], np.int8)
    assert np.array_equal(out, want)
    assert np.array_equal(expand4(nib), expand4_first_form(nib))


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


# ---- the tcgen05 form (k_match_umma / k_expand_train): operand image and epilogue
MU_PLANE, MU_TILE = 2048, 32768


def umma_image(desc: np.ndarray) -> np.ndarray:
    """128 descriptors -> the 32 KB B-tile image as the kernel writes it: thread (row, half) expands halfword c of its half
    into the 16-byte chunk at plane 8 * half + c, row group (row >> 3), row-in-group (row & 7)."""
    assert desc.shape == (128, 32)
    img = np.zeros(MU_TILE, np.uint8)
    hw = desc.view(np.uint16).reshape(128, 16)  # little endian: halfword c = bits 16 c .. 16 c + 15
    for row in range(128):
        for c in range(16):
            h = int(hw[row, c])
            words = expand4(np.array([h & 15, (h >> 4) & 15, (h >> 8) & 15, h >> 12], np.uint32))
            off = c * MU_PLANE + (row >> 3) * 128 + (row & 7) * 16
            img[off:off + 16] = words.view(np.uint8)
    return img


def umma_operand(img: np.ndarray, kstep: int, lbo: int = MU_PLANE, sbo: int = 128) -> np.ndarray:
    """what one MMA reads through the shared-memory descriptor of K step `kstep` (canonical K-major, no swizzle:
    ((8, n), 2) : ((16 B, SBO), LBO) in units of one 16-byte row of a core matrix; start = 2 planes per K step):
    a (128 rows, 32) s8 matrix"""
    out = np.zeros((128, 32), np.int8)
    start = 2 * MU_PLANE * kstep
    for row in range(128):
        for c in range(2):
            off = start + c * lbo + (row >> 3) * sbo + (row & 7) * 16
            out[row, 16 * c:16 * c + 16] = img[off:off + 16].view(np.int8)
    return out


def test_umma_image_and_descriptor_strides_give_the_dot_product():
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (128, 32), dtype=np.uint8)
    qi, ti = umma_image(q), umma_image(t)
    assert len(np.unique(np.concatenate([[c * MU_PLANE + (r >> 3) * 128 + (r & 7) * 16 for r in range(128)] for c in range(16)]))) == 2048
    acc = np.zeros((128, 128), np.int32)
    for j in range(8):  # eight K steps of 32 s8 each
        acc += umma_operand(qi, j).astype(np.int32) @ umma_operand(ti, j).astype(np.int32).T
    ham = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2).astype(np.int32)
    assert np.array_equal(acc, 256 - 2 * ham)
    # every expanded byte is +1 or -1, and the other LBO / SBO assignment reads something else
    assert set(np.unique(qi.view(np.int8)).tolist()) == {-1, 1}
    swapped = sum(umma_operand(qi, j, lbo=128, sbo=MU_PLANE % MU_TILE).astype(np.int32)[:16] @ umma_operand(ti, j).astype(np.int32)[:16].T
                  for j in range(1))
    assert not np.array_equal(swapped, (256 - 2 * ham)[:16, :16])


def _scan_plain(dot_row, k):
    best = [0xFFFFFFFF] * k
    for n, d in enumerate(dot_row):
        key = (int(d) * -(1 << 21) + ((256 << 21) + n)) & 0xFFFFFFFF
        if k == 1:
            best[0] = min(best[0], key)
        else:
            hi = max(key, best[0]); best[0] = min(key, best[0]); best[1] = min(best[1], hi)
    return best


def _scan_chunk_max(dot_row, k, cnt_last=None):
    """the kernel's epilogue: a 32-column chunk is keyed only if its maximum dot product beats that of the K-th best"""
    best = [0xFFFFFFFF] * k
    thr = 256 - 2 * (0xFFFFFFFF >> 22)
    keyed = 0
    n_cols = len(dot_row)
    for c0 in range(0, n_cols, 32):
        chunk = dot_row[c0:c0 + 32]
        if len(chunk) == 32 and int(chunk.max()) <= thr:
            continue
        keyed += 1
        for i, d in enumerate(chunk):
            key = (int(d) * -(1 << 21) + ((256 << 21) + c0 + i)) & 0xFFFFFFFF
            if k == 1:
                best[0] = min(best[0], key)
            else:
                hi = max(key, best[0]); best[0] = min(key, best[0]); best[1] = min(best[1], hi)
        thr = 256 - 2 * (best[k - 1] >> 22)
    return best, keyed


def test_chunk_maximum_epilogue_equals_the_plain_key_scan():
    rng = np.random.default_rng(21)
    q = rng.integers(0, 256, (24, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1000, 32), dtype=np.uint8)  # 31 full chunks + one of 8 columns
    t[600:700] = t[100:200]                 # exact duplicates later in the scan: equal distance, higher index never wins
    t[40] = q[0]; t[900] = q[0]             # distance 0 twice
    t[5] = q[1] ^ np.uint8(3); t[995] = q[1] ^ np.uint8(3)   # the best pair of row 1 spans the first and the partial chunk
    q[2] = t[999]                           # best in the partial chunk
    dot = (_pm1(q) @ _pm1(t).T).astype(np.int64)
    skipped_any = False
    for k in (1, 2):
        for r in range(len(q)):
            plain = _scan_plain(dot[r], k)
            fast, keyed = _scan_chunk_max(dot[r], k)
            assert fast == plain, (k, r)
            skipped_any |= keyed < 20
    assert skipped_any, "the model never skipped a chunk: test too weak"
    assert (_scan_plain(dot[0], 2)[0] & 0x3FFFFF, _scan_plain(dot[0], 2)[1] & 0x3FFFFF) == (40, 900)
