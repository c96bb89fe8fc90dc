"""GPU tests of the int8 tcgen05 Ozaki-scheme correlation path (precision="ozaki").

The digit-slice products are exact in the tensor core (int8 x int8 -> int32), so the only error is the
fixed-point resolution of the slices: with the default 6 slices the correlations agree with the float64
oracle to ~1e-12 absolute, with 8 slices to the oracle's own rounding level.  Gate here: 1e-10 (6 slices),
far inside the north star's 1e-6, and the whole path must reproduce the FP64 oracle's assignments."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _corr_ozaki(handle, rna, dna, nsl, with_c=True, with_ct=True):
    torch = _torch()
    lib, h = handle.lib, handle.h
    M, G = rna.shape
    N = dna.shape[0]
    ldk = lib.mcd_padded_k_split(G)
    d_r = torch.from_numpy(rna).cuda()
    d_d = torch.from_numpy(dna).cuda()
    a8 = torch.full((nsl, M, ldk), 55, dtype=torch.int8, device="cuda")
    b8 = torch.full((nsl, N, ldk), 55, dtype=torch.int8, device="cuda")
    na = torch.empty(M, dtype=torch.float64, device="cuda")
    nb = torch.empty(N, dtype=torch.float64, device="cuda")
    sa = torch.empty(M, dtype=torch.float64, device="cuda")
    sb = torch.empty(N, dtype=torch.float64, device="cuda")
    handle.check(lib.mcd_standardize_ozaki(h, d_r.data_ptr(), M, G, G, a8.data_ptr(), nsl, sa.data_ptr(), na.data_ptr()))
    handle.check(lib.mcd_standardize_ozaki(h, d_d.data_ptr(), N, G, G, b8.data_ptr(), nsl, sb.data_ptr(), nb.data_ptr()))
    ldc, ldct = N + 2, M + 2
    c = torch.full((M, ldc), 7.0, dtype=torch.float64, device="cuda")
    ct = torch.full((N, ldct), 7.0, dtype=torch.float64, device="cuda")
    handle.check(lib.mcd_corr_ozaki(h, a8.data_ptr(), M, b8.data_ptr(), N, G, ldk, nsl, sa.data_ptr(), sb.data_ptr(),
                                    na.data_ptr(), nb.data_ptr(), c.data_ptr() if with_c else None, ldc,
                                    ct.data_ptr() if with_ct else None, ldct))
    handle.synchronize()
    return c.cpu().numpy(), ct.cpu().numpy(), a8.cpu().numpy(), sa.cpu().numpy(), na.cpu().numpy()


@pytest.mark.parametrize("nsl", [4, 6, 8])
def test_ozaki_digits_reconstruct_unit_rows(handle, nsl):
    from oracle import restatement as R

    rng = np.random.default_rng(3)
    x = np.log1p(rng.poisson(4.0, size=(50, 777)).astype(np.float64))
    x[7] = 1.0  # zero-variance cell: all digits zero
    _, _, a8, sa, na = _corr_ozaki(handle, x, x[:5].copy(), nsl)
    xc, nrm = R.standardise(x)
    unit = np.divide(xc, nrm[:, None], out=np.zeros_like(xc), where=nrm[:, None] > 0)
    assert a8.min() >= -64 and a8.max() <= 63
    q = np.zeros(a8.shape[1:], dtype=np.float64)
    for t in range(nsl):
        q = q * 128.0 + a8[t].astype(np.float64)
    rec = q[:, :777] * 2.0 ** -(7 * nsl - 1) * sa[:, None]
    # resolution: half a unit of the last digit, relative to the row scale
    tol = 2.0 ** -(7 * nsl - 1) * sa.max()
    assert np.abs(rec - unit).max() <= max(tol, 4e-16)
    assert (a8[:, :, 777:] == 0).all()
    assert (a8[:, 7, :] == 0).all()
    assert np.abs(na - nrm).max() <= 1e-12 * nrm.max()
    # the row scaling puts the largest digit vector entry in [0.25, 0.5)
    live = nrm > 0
    top = np.abs(q[live]).max(axis=1) * 2.0 ** -(7 * nsl - 1)
    assert (top >= 0.25 - 1e-9).all() and (top < 0.5).all()


def _digits_of(handle, x_dev, ncells, G, ldx, nsl):
    torch = _torch()
    lib, h = handle.lib, handle.h
    ldk = lib.mcd_padded_k_split(G)
    a8 = torch.full((nsl, ncells, ldk), 55, dtype=torch.int8, device="cuda")
    sa = torch.empty(ncells, dtype=torch.float64, device="cuda")
    na = torch.empty(ncells, dtype=torch.float64, device="cuda")
    handle.check(lib.mcd_standardize_ozaki(h, x_dev.data_ptr(), ncells, G, ldx, a8.data_ptr(), nsl, sa.data_ptr(),
                                           na.data_ptr()))
    handle.synchronize()
    return a8.cpu().numpy(), sa.cpu().numpy(), na.cpu().numpy()


@pytest.mark.parametrize("nsl", [6, 8])
@pytest.mark.parametrize("G", [5, 130, 256, 1000, 3000, 4097, 8192, 9001, 13000, 16385, 20000, 24576])
def test_digit_fast_path_every_launch_shape(handle, G, nsl):
    """K1's digit fast path (4 consecutive genes per thread, 256-bit loads, fma rounding) in each of its launch
    shapes, from aligned and from strided / unaligned input (scalar loads): digits reconstruct the unit rows to half
    a unit of the last digit, padding is zero, and zero-variance rows are exactly zero whatever their constant
    (sum/G need not reproduce the constant: log1p(2) at G = 3000 does not)."""
    from oracle import restatement as R

    torch = _torch()
    rng = np.random.default_rng(G)
    ncells = 9
    x = np.log1p(rng.poisson(3.0, size=(ncells, G)).astype(np.float64)) + rng.random((ncells, G)) * 1e-3
    x[2] = np.log1p(2.0)
    x[5] = 1.0 / 3.0
    x[7] = 0.1
    xc, nrm = R.standardise(x)
    flat = np.array([2, 5, 7])
    nrm[flat] = 0.0  # the oracle's pairwise mean may leave rounding residue on a constant row; the kernel must not
    xc[flat] = 0.0
    unit = np.divide(xc, nrm[:, None], out=np.zeros_like(xc), where=nrm[:, None] > 0)
    for ldx, off in ((G, 0), (G + 3, 1)):  # contiguous (vector loads when aligned) and strided + misaligned
        buf = torch.zeros(ncells * ldx + 8, dtype=torch.float64, device="cuda")
        view = buf[off:off + ncells * ldx].view(ncells, ldx)
        view[:, :G] = torch.from_numpy(x).cuda()
        a8, sa, na = _digits_of(handle, view, ncells, G, ldx, nsl)
        assert a8.min() >= -64 and a8.max() <= 63
        q = np.zeros(a8.shape[1:], dtype=np.float64)
        for t in range(nsl):
            q = q * 128.0 + a8[t].astype(np.float64)
        rec = q[:, :G] * 2.0 ** -(7 * nsl - 1) * sa[:, None]
        tol = 2.0 ** -(7 * nsl - 1) * sa.max()
        assert np.abs(rec - unit).max() <= max(tol, 1e-15), (G, nsl, ldx)  # FP64 rounding of the oracle itself
        assert (a8[:, :, G:] == 0).all()
        assert (a8[:, flat, :] == 0).all()
        assert (na[flat] == 0).all()
        assert np.abs(na - nrm).max() <= 1e-12 * nrm.max()


def test_digit_fast_path_matches_generic_kernel(handle, monkeypatch):
    """Same rows through the fast digit kernel and the generic one (MCD_K1_GENERIC=1): identical scales, and digit
    vectors that differ by at most one unit of the last digit (the fast path rounds c*mul once, in an fma)."""
    torch = _torch()
    rng = np.random.default_rng(11)
    G, ncells, nsl = 5000, 64, 6
    x = torch.from_numpy(np.log1p(rng.poisson(2.0, size=(ncells, G)).astype(np.float64))).cuda()
    a_fast, s_fast, n_fast = _digits_of(handle, x, ncells, G, G, nsl)
    monkeypatch.setenv("MCD_K1_GENERIC", "1")
    a_gen, s_gen, n_gen = _digits_of(handle, x, ncells, G, G, nsl)
    monkeypatch.delenv("MCD_K1_GENERIC")
    assert (s_fast == s_gen).all()
    assert np.abs(n_fast - n_gen).max() <= 1e-13 * n_gen.max()
    qf = np.zeros(a_fast.shape[1:])
    qg = np.zeros(a_gen.shape[1:])
    for t in range(nsl):
        qf = qf * 128.0 + a_fast[t]
        qg = qg * 128.0 + a_gen[t]
    # different summation order of the mean -> the centred values differ by ~1e-16 relative -> a few last-digit units
    assert np.abs(qf - qg).max() <= 8.0


@pytest.mark.parametrize("shape", [(4, 4, 6), (128, 256, 64), (130, 257, 100), (300, 200, 1000), (515, 700, 4099),
                                   (1000, 249, 20000)])
def test_corr_ozaki_kernel(handle, shape):
    from oracle import restatement as R

    M, N, G = shape
    rng = np.random.default_rng(M + 3 * N)
    rna = np.log1p(rng.poisson(4.0, size=(M, G)).astype(np.float64))
    dna = np.log1p(rng.integers(1, 5, size=(N, G)) * (1 + 0.05 * rng.standard_normal((N, G))))
    if N > 2:
        dna[1] = 2.0
    ref = R.correlation_matrix(rna, dna)
    for nsl, tol in ((6, 1e-10), (8, 1e-13), (5, 1e-8)):
        c, ct, _, _, _ = _corr_ozaki(handle, rna, dna, nsl)
        err = np.abs(c[:, :N] - ref).max()
        print("ozaki int8 x%d max |dcorr| for" % nsl, shape, "=", err)
        assert err < tol
        assert (c[:, N:] == 7.0).all() and (ct[:, M:] == 7.0).all()
        assert (ct[:, :M] == c[:, :N].T).all()
        if N > 2:
            assert (c[:, 1] == 0).all()


def test_corr_ozaki_single_output(handle):
    """Only C or only C^T requested: the pass partials then live in the one that exists."""
    from oracle import restatement as R

    rng = np.random.default_rng(11)
    rna = rng.standard_normal((300, 500))
    dna = rng.standard_normal((270, 500))
    ref = R.correlation_matrix(rna, dna)
    c, ct, _, _, _ = _corr_ozaki(handle, rna, dna, 6, with_ct=False)
    assert np.abs(c[:, :270] - ref).max() < 1e-10 and (ct == 7.0).all()
    c, ct, _, _, _ = _corr_ozaki(handle, rna, dna, 6, with_c=False)
    assert np.abs(ct[:, :300] - ref.T).max() < 1e-10 and (c == 7.0).all()


@pytest.mark.parametrize("name", ["C2", "C3", "C4"])
def test_whole_path_ozaki(handle, name):
    from conftest import tie_report
    from macrodna_b200 import synth
    from oracle import restatement as R

    inst = synth.make_config_arrays(name)
    M, G = inst.rna.shape
    N = inst.dna.shape[0]
    corr = np.empty((M, N))
    assign, step, objs, stats = handle.cell2cell(inst.rna, inst.dna, M, N, G, precision="ozaki", corr_out=corr)
    c_ref, a_ref, s_ref, o_ref = R.cell2cell_arrays(inst.rna, inst.dna)
    err = np.abs(corr - c_ref).max()
    print(name, "ozaki: max |dcorr| %.3g" % err)
    assert err < 1e-13  # these instances are below the 8-slice threshold (mcd_ozaki_slices_for)
    ident, rep = tie_report(c_ref, assign, step, a_ref, s_ref, rel=1e-9)
    assert ident, rep
    assert np.abs(objs - o_ref).max() <= 1e-9 * np.abs(o_ref).max()


def test_ozaki_long_rows_and_fp64_fallback(handle):
    """G > 24 576 takes K1's three-sweep variant (digits included); G beyond the exact-int32 bound of the integer
    path (nsl * 4096 * ldk8 < 2^31) silently takes the FP64 tensor pipe instead -- both still device paths."""
    from oracle import restatement as R

    rng = np.random.default_rng(21)
    for G, tol in ((30000, 1e-12), (90000, 1e-12)):
        rna = np.log1p(rng.poisson(3.0, size=(40, G)).astype(np.float64))
        dna = np.log1p(rng.integers(1, 5, size=(17, G)) * (1 + 0.05 * rng.standard_normal((17, G))))
        corr = np.empty((40, 17))
        assign, step, objs, _ = handle.cell2cell(rna, dna, 40, 17, G, precision="ozaki", corr_out=corr)
        c_ref, a_ref, s_ref, o_ref = R.cell2cell_arrays(rna, dna)
        assert np.abs(corr - c_ref).max() < tol
        assert (assign == a_ref).all() and (step == s_ref).all()


def test_ozaki_chunked_host_input_equals_device_input(handle):
    """Host buffers above 768 MB are staged in tile-aligned RNA chunks overlapped with K1 + K2c; the result must be
    bit-identical to the one-shot device-input path (same kernels, same tiles)."""
    torch = _torch()
    from macrodna_b200 import _lib

    rng = np.random.default_rng(2)
    M, N, G = 5300, 300, 20000
    rna = np.log1p(rng.poisson(2.0, size=(M, G)).astype(np.float64))
    dna = np.log1p(rng.integers(1, 5, size=(N, G)) * (1 + 0.05 * rng.standard_normal((N, G))))
    c_host = np.empty((M, N))
    a1, s1, o1, st1 = handle.cell2cell(rna, dna, M, N, G, precision="ozaki", corr_out=c_host)
    d_r, d_d = torch.from_numpy(rna).cuda(), torch.from_numpy(dna).cuda()
    c_dev = np.empty((M, N))
    a2, s2, o2, st2 = handle.cell2cell(d_r.data_ptr(), d_d.data_ptr(), M, N, G, in_space=_lib.MEM_DEVICE,
                                       precision="ozaki", corr_out=c_dev)
    assert (c_host == c_dev).all() and (a1 == a2).all() and (s1 == s2).all() and (o1 == o2).all()
    assert st1.ms_h2d >= 0.0 and st1.kernel_launches > st2.kernel_launches  # more K1 / K2c launches: chunks


def test_ozaki_nonfinite_input_raises(handle):
    rng = np.random.default_rng(4)
    rna = rng.random((20, 300))
    dna = rng.random((9, 300))
    rna[3, 17] = np.nan
    with pytest.raises(ValueError):
        handle.cell2cell(rna, dna, 20, 9, 300, precision="ozaki")
    rna[3, 17] = 0.5
    handle.cell2cell(rna, dna, 20, 9, 300, precision="ozaki")  # the flag was cleared
