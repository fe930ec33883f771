"""Philox4x32-10 (csrc/philox.cuh) known-answer tests and the rejection-free samplers'
distributions (they must match the reference's rejection loops: uniform disk / sphere / ball)."""
import ctypes as C

import numpy as np

from tests.emu import pyemu


def philox_py(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def _lib():
    pyemu.build()
    L = C.CDLL(pyemu.PATH)
    L.emu_philox.argtypes = [C.c_uint] * 6 + [C.c_void_p]
    L.emu_samplers.argtypes = [C.c_float] * 3 + [C.c_void_p]
    return L


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    L = _lib()
    for ctr, key, want in kat:
        assert tuple(philox_py(ctr, key)) == want
        out = np.zeros(4, np.uint32)
        L.emu_philox(*ctr, *key, out.ctypes.data)
        assert tuple(int(x) for x in out) == want
    rng = np.random.Generator(np.random.Philox(1))
    for _ in range(50):
        v = [int(x) for x in rng.integers(0, 2 ** 32, 6)]
        out = np.zeros(4, np.uint32)
        L.emu_philox(*v, out.ctypes.data)
        assert [int(x) for x in out] == philox_py(v[:4], v[4:])


def test_samplers_have_the_reference_distributions():
    """uniform disk (E r^2 = 1/2), uniform sphere (|v| = 1, E v = 0, E z^2 = 1/3) and uniform ball
    (E r^2 = 3/5) — what vec3.h:103-130's rejection loops produce."""
    L = _lib()
    rng = np.random.Generator(np.random.Philox(2))
    u = rng.random((20000, 3)).astype(np.float32)
    out = np.zeros((len(u), 9), np.float32)
    for i in range(len(u)):
        L.emu_samplers(float(u[i, 0]), float(u[i, 1]), float(u[i, 2]), out[i].ctypes.data)
    disk, sph, ball = out[:, 0:2], out[:, 2:5], out[:, 5:8]
    assert np.all((disk ** 2).sum(1) < 1.0 + 1e-6) and abs((disk ** 2).sum(1).mean() - 0.5) < 0.01
    assert np.abs(disk.mean(0)).max() < 0.01
    np.testing.assert_allclose((sph ** 2).sum(1), 1.0, atol=2e-6)
    assert np.abs(sph.mean(0)).max() < 0.015 and np.abs((sph ** 2).mean(0) - 1 / 3).max() < 0.01
    r2 = (ball ** 2).sum(1)
    assert np.all(r2 <= 1.0 + 1e-6) and abs(r2.mean() - 0.6) < 0.01 and np.abs(ball.mean(0)).max() < 0.015
