"""Known-answer tests pinning oracle/philox.py (the specification of the device sampler) to the
published Philox4x32-10 vectors (Random123 kat_vectors), and sanity of the draw layout.  CPU only."""
import numpy

from oracle import philox

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, expect in KAT:
        out = philox.philox4x32_10(*[numpy.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(o[0]) for o in out) == expect


def test_philox_vectorised_matches_scalar():
    idx = numpy.arange(100, dtype=numpy.uint64)
    out = philox.philox4x32_10(idx, 0, 7, 3, 123, 456)
    for i in (0, 17, 99):
        one = philox.philox4x32_10(numpy.array([i]), 0, 7, 3, 123, 456)
        assert all(int(a[i]) == int(b[0]) for a, b in zip(out, one))


def test_normals_and_mixture_statistics():
    n = 200000
    z, uc = philox.standard_normals5(numpy.arange(n), step=3, sub=0, seed=42)
    assert numpy.abs(z.mean(axis=0)).max() < 0.01
    assert numpy.abs(z.std(axis=0) - 1).max() < 0.01
    assert numpy.abs(numpy.corrcoef(z.T) - numpy.eye(5)).max() < 0.01
    assert 0 < uc.min() and uc.max() <= 1
    from oracle import mixture
    x, comp = philox.draw_mixture5(mixture.STATE_MEANS, mixture.STATE_COVS, mixture.STATE_WEIGHTS,
                                   numpy.arange(n), 0, 0, 7)
    assert abs((comp == 0).mean() - 0.75) < 0.005
    var = 0.75 * numpy.diag(mixture.STATE_COVS[0]) + 0.25 * numpy.diag(mixture.STATE_COVS[1])
    assert numpy.allclose(x.var(axis=0), var, rtol=0.03)


def test_grouped_layout_statistics_and_independence():
    """The predict kernel's grouped layout: rows of one group share Philox calls but no words."""
    n = 200000
    z, uc = philox.grouped_normals5(numpy.arange(n), step=5, seed=11)
    assert numpy.abs(z.mean(axis=0)).max() < 0.01
    assert numpy.abs(z.std(axis=0) - 1).max() < 0.01
    assert numpy.abs(numpy.corrcoef(z.T) - numpy.eye(5)).max() < 0.01
    flat = z.reshape(n // 4, 20)                                # the 20 normals of every group
    assert numpy.abs(numpy.corrcoef(flat.T) - numpy.eye(20)).max() < 0.02
    assert abs((uc > 0.75).mean() - 0.25) < 0.005
    # a row's draw depends on its global index only
    z2, _ = philox.grouped_normals5(numpy.arange(1001, 1013), step=5, seed=11)
    assert numpy.array_equal(z2, z[1001:1013])
