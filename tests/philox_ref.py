"""Philox4x32-10 in Python ints (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
SC'11) and the rank -> guess map of qcf_sweep_random (include/qimcifa_cuda.h).  Test-side twin only."""
M32 = 0xFFFFFFFF


def philox4x32_10(counter, key):
    c = list(counter)
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def random_guess(ctr, seed, periods, period, table_a, table_b):
    r = philox4x32_10([ctr & M32, (ctr >> 32) & M32, 0, 0], (seed & M32, (seed >> 32) & M32))
    m = (((r[1] << 32) | r[0]) * periods) >> 64
    ia = (r[2] * len(table_a)) >> 32
    ib = (r[3] * len(table_b)) >> 32
    return m * period + (table_a[ia] + table_b[ib]) % period
