"""Seeded synthetic stand-in for the MineRL Treechop frames (the dataset and
`red-trees/*.npy` are not available offline).

Layout matches what `Handler.load_data` hands to the loops
(reference main.py:113-129): frames uint8 [N,64,64,3] (NHWC), labels float64
[7,N], indices uint16/int [N].  Half of the frames carry a painted "trunk"
(the rewarding object); label rows 1..4 mimic the clipped gamma-discounted
reward of reference main.py:1336-1346.
"""
import numpy as np

TRUNK_RGB = np.array([200, 140, 60], dtype=np.int64)


def synthetic_frames(n, seed=0, with_gt=False):
    """Return (X uint8 [n,64,64,3], Y float64 [7,n], I int32 [n][, GT bool [n,64,64]])."""
    rng = np.random.default_rng(seed)
    X = rng.integers(0, 120, size=(n, 64, 64, 3), dtype=np.uint8)
    has = rng.random(n) < 0.5
    x0 = rng.integers(8, 48, size=n)
    w = rng.integers(6, 12, size=n)
    y0 = rng.integers(0, 20, size=n)
    jit = rng.integers(0, 30, size=(n, 3))
    gt = np.zeros((n, 64, 64), dtype=bool)
    for i in np.nonzero(has)[0]:
        col = (TRUNK_RGB + jit[i]).astype(np.uint8)
        X[i, y0[i]:, x0[i]:x0[i] + w[i], :] = col
        gt[i, y0[i]:, x0[i]:x0[i] + w[i]] = True
    u = rng.random((4, n))
    Y = np.zeros((7, n), dtype=np.float64)
    Y[0] = has
    Y[1:5] = np.where(has[None, :], 0.85 + 0.15 * u, 0.1 * u)
    I = np.arange(n, dtype=np.int32)
    if with_gt:
        return X, Y, I, gt
    return X, Y, I


def sparse_event_labels(n, seed=0, gammas=(0.98, 0.97, 0.96, 0.95), episode=400, p_event=0.01):
    """Alternative label recipe: Bernoulli reward events per pseudo-episode, then the
    backward recursion r[t] = min(r[t] + gamma*r[t+1], 1) of reference main.py:1340-1344."""
    rng = np.random.default_rng(seed)
    raw = (rng.random(n) < p_event).astype(np.float64)
    Y = np.zeros((7, n), dtype=np.float64)
    Y[0] = raw
    for gi, g in enumerate(gammas):
        r = raw.copy()
        for s in range(0, n, episode):
            e = min(n, s + episode)
            for t in range(e - 2, s - 1, -1):
                r[t] = min(r[t] + g * r[t + 1], 1.0)
        Y[1 + gi] = r
    return Y


def perturbed_state(shapes, seed, scale=1.0):
    """Deterministic non-degenerate weights for parity tests (numpy PCG64 is
    platform-independent).  `shapes` = ordered {key: shape}; fan-in scaled
    uniform so activations neither vanish nor explode, then widened by `scale`
    (random-init nets are near-constant)."""
    rng = np.random.default_rng(seed)
    out = {}
    for k, shp in shapes.items():
        shp = tuple(int(s) for s in shp)
        if k.endswith("weight"):
            fan_in = int(np.prod(shp[1:])) if len(shp) > 1 else shp[0]
            b = scale * np.sqrt(3.0 / fan_in)
            out[k] = rng.uniform(-b, b, size=shp).astype(np.float32)
        else:
            out[k] = rng.uniform(-0.1, 0.1, size=shp).astype(np.float32)
    return out


def critic_shapes(chfak=1, bottleneck=32, colorchs=3):
    """state_dict key -> shape for NewCritic (reference nets.py:160-195)."""
    c = chfak
    d = [8 * c, 8 * c, 8 * c, 16 * c]
    nb = bottleneck * c
    return {
        "features.0.weight": (d[0], colorchs, 3, 3), "features.0.bias": (d[0],),
        "features.3.weight": (d[1], d[0], 3, 3), "features.3.bias": (d[1],),
        "features.6.weight": (d[2], d[1], 3, 3), "features.6.bias": (d[2],),
        "features.10.weight": (d[3], d[2], 3, 3), "features.10.bias": (d[3],),
        "features.14.weight": (nb, d[3], 4, 4), "features.14.bias": (nb,),
        "crit.1.weight": (nb, nb), "crit.1.bias": (nb,),
        "crit.4.weight": (1, nb), "crit.4.bias": (1,),
    }


def masker_shapes(chfak=1, bottleneck=32, colorchs=3, masker_channels=16):
    """state_dict key -> shape for UnetDecoder (reference nets.py:452-492)."""
    c = chfak
    e = [8 * c, 8 * c, 8 * c, 16 * c]
    d = [8 * c, 8 * c, 8 * c, 16 * c]
    nb = bottleneck * c
    return {
        "dec_model.0.weight": (d[0], e[0] + d[1], 3, 3), "dec_model.0.bias": (d[0],),
        "dec_model.1.weight": (d[1], e[1] + d[2], 3, 3), "dec_model.1.bias": (d[1],),
        "dec_model.2.weight": (d[2], e[2] + d[3], 3, 3), "dec_model.2.bias": (d[2],),
        "dec_model.3.weight": (d[3], e[3] + nb, 3, 3), "dec_model.3.bias": (d[3],),
        "dec_model.4.weight": (nb, nb, 1, 1), "dec_model.4.bias": (nb,),
        "masker.0.weight": (masker_channels, colorchs + d[0], 3, 3), "masker.0.bias": (masker_channels,),
        "masker.2.weight": (1, masker_channels, 3, 3), "masker.2.bias": (1,),
    }
