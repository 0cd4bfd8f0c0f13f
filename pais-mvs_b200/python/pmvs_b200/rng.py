"""Python statement of csrc/pmvs_rng.h (counter-based replacement of psosolver.cpp:60-68's srand/rand)."""
M64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15


def mix64(z):
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def stream_key(seed, patch_id, run):
    k = mix64((seed + GOLD * ((patch_id + 1) & 0xFFFFFFFF)) & M64)
    return mix64((k + GOLD * ((run + 1) & 0xFFFFFFFF)) & M64)


def rand31(key, ctr):
    return mix64((key + GOLD * (ctr + 1)) & M64) >> 33
