"""Philox4x32-10 counter-based RNG in numpy -- TEST INFRASTRUCTURE ONLY.

This is the CPU twin of ``theanet_b200/csrc/philox.cuh``.  The reference draws its
randomness from Theano ``RandomStreams`` (MT19937 per op; theanet/layer/dropout.py:10-12,
theanet/layer/inlayers.py:72,81,94,101,107,112,141), which cannot be reproduced on a GPU; the
B200 path therefore defines its own counter-based streams, and this file restates them with
integer arithmetic so the oracle can regenerate *exactly* the masks / noise the device used.

Counter / key layout (shared with the device code, see DESIGN.md "Randomness"):
    key  = (seed & 0xffffffff, seed >> 32)        seed = the layer's stream seed
    ctr  = (block, sample, step, purpose)         block = element_index // 4
    element j of a sample uses word (j % 4) of block j // 4.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs may
import this module.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)

# purposes (ctr word 3)
PURPOSE_DROPOUT = 0
PURPOSE_FLIP = 1
PURPOSE_NOISE = 2
PURPOSE_SCALARS = 3
PURPOSE_COLOR = 4
PURPOSE_AUX = 5


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All counter words broadcast together; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        *[np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0          # 32x32 -> 64 bit, no overflow in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK32
        hi1, lo1 = p1 >> S32, p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def _split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def random_words(seed, purpose, step, samples, n_per_sample):
    """uint32 words, shape (len(samples), n_per_sample): word j of sample s."""
    k0, k1 = _split_seed(seed)
    samples = np.asarray(samples, dtype=np.uint64).reshape(-1, 1)
    nblk = (n_per_sample + 3) // 4
    blocks = np.arange(nblk, dtype=np.uint64).reshape(1, -1)
    r = philox4x32_10(blocks, samples, np.uint64(int(step) & 0xFFFFFFFF),
                      np.uint64(purpose), k0, k1)
    words = np.stack(r, axis=-1).reshape(samples.shape[0], nblk * 4)
    return words[:, :n_per_sample]


def bernoulli_threshold(p):
    """uint32 threshold t such that (word < t) has probability ~p; p>=1 handled by the caller."""
    t = int(np.floor(float(p) * 4294967296.0))
    return max(0, min(t, 0xFFFFFFFF))


def bernoulli_mask(seed, purpose, step, samples, n_per_sample, p):
    """1.0 where the event of probability p happens.  Shape (len(samples), n_per_sample) float32."""
    w = random_words(seed, purpose, step, samples, n_per_sample)
    if p >= 1.0:
        return np.ones(w.shape, np.float32)
    return (w < np.uint32(bernoulli_threshold(p))).astype(np.float32)


def uniform01(words):
    """float64 uniform in (0,1) from uint32 words: (w + 0.5) * 2^-32."""
    return (words.astype(np.float64) + 0.5) * (1.0 / 4294967296.0)


def elastic_noise(seed, step, n):
    """n standard normals (float32) by Box-Muller in float64 on word pairs (w0,w1)->(z0,z1)."""
    npair = (n + 1) // 2
    w = random_words(seed, PURPOSE_NOISE, step, [0], 2 * npair)[0]
    u1 = uniform01(w[0::2])
    u2 = uniform01(w[1::2])
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    z = np.empty(2 * npair, np.float64)
    z[0::2] = rad * np.cos(ang)
    z[1::2] = rad * np.sin(ang)
    return z[:n].astype(np.float32)


def elastic_scalars(seed, step):
    """8 float32 uniforms in (0,1): [trans_y, trans_x, origin_y, origin_x, zoom_y, zoom_x, angle, spare]."""
    w = random_words(seed, PURPOSE_SCALARS, step, [0], 8)[0]
    return uniform01(w).astype(np.float32)


def color_uniforms(seed, step, samples, maps):
    """ColorLayer draws: (len(samples), maps, 3) float32 in (-1,1); block = map, words x, y, z."""
    w = random_words(seed, PURPOSE_COLOR, step, samples, 4 * maps).reshape(len(samples), maps, 4)
    return (2.0 * uniform01(w[:, :, :3]) - 1.0).astype(np.float32)


def aux_uniforms(seed, step, samples):
    """LocationInfo mixing weights: one float32 uniform in (0,1) per sample (block 0, word x)."""
    w = random_words(seed, PURPOSE_AUX, step, samples, 1)[:, 0]
    return uniform01(w).astype(np.float32)
