"""CPU tests (-m "not gpu"): the C-ABI library loads and exports what include/lcf.h declares, the host-side
packing reproduces the reference's filter synthesis, the drop-in argument validation, the LC stand-in, and the
multi-GPU exchange logic under a 2-rank gloo group.  No compute entry point is called without a GPU."""
import os
import re
import socket

import numpy as np
import pytest

from oracle import reference_port as rp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))


def test_library_exports_every_declared_symbol():
    from lightcurve_fitting_b200 import _capi
    hdr = open(os.path.join(ROOT, 'include', 'lcf.h')).read()
    declared = set(re.findall(r'\b(lcf_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 35
    L = _capi.lib()
    bound = {name for name, _, _ in _capi.SYMBOLS}
    for name in declared:
        assert hasattr(L, name), 'liblcf_b200.so does not export ' + name
    assert declared == bound, 'ctypes table and include/lcf.h disagree: %s' % (declared ^ bound)
    assert L.lcf_abi_version() == 1
    assert L.lcf_device_count() >= 0


def test_problem_desc_layout_matches_header():
    """sizeof(lcf_problem_desc) as laid out by ctypes == the C struct (8 int32 + 17 doubles + pointers ...)."""
    import ctypes as C
    from lightcurve_fitting_b200._capi import ProblemDesc
    expected = 8 * 4 + 8 + 16 * 8 + 5 * 8 + 2 * 4 + 2 * 8 + 8 + 4 * 8 + 5 * 8
    assert C.sizeof(ProblemDesc) == expected


def test_product_fails_loudly_without_gpu():
    from lightcurve_fitting_b200 import _capi
    if _capi.lib().lcf_device_count() > 0:
        pytest.skip('a GPU is present')
    from tests import workloads as W
    with pytest.raises(_capi.LcfError, match='no CPU fallback'):
        W.example_sc4(npoints=20).device_problem('fp64')
    from lightcurve_fitting_b200 import models as M
    with pytest.raises(_capi.LcfError):
        M.planck_fast(np.array([500.]), 10., 1.)


def test_filter_registry_and_curves_match_reference():
    from lightcurve_fitting_b200.filters import filtdict, all_filters, Filter
    assert len(all_filters) == 67 and Filter.order[0] == 'FUV'
    for n in G['filters/names']:
        f = filtdict[str(n)]
        np.testing.assert_allclose(f.trans['freq'], G['filters/%s/freq' % n], rtol=1e-14)
        np.testing.assert_allclose(f.trans['T_norm_per_freq'], G['filters/%s/Tn' % n], rtol=1e-12)
        s = G['filters/%s/scalars' % n]
        np.testing.assert_allclose([f.freq_eff, f.dfreq, f.wl_eff, f.m0, f.M0], s, rtol=1e-12)
        assert f.char == str(G['filters/%s/char' % n])
    assert filtdict['unfilt.'] is filtdict['0'] is filtdict['clear']
    assert filtdict['U'] < filtdict['B'] < filtdict['I']
    assert sorted([filtdict['r'], filtdict['g']])[0] == filtdict['g']


def test_packed_bank_reproduces_reference_synthesis():
    """sum_k w_k/(exp(alpha_k/T)-1) * R^2 from the packed bank == the reference's np.trapz synthesis."""
    from lightcurve_fitting_b200.filters import filtdict, pack_bank
    names, Tb, Rb = [str(n) for n in G['bb/names']], G['bb/T'], G['bb/R']
    for tag, kw in (('plain', {}), ('z', {'z': 0.05}), ('cut', {'z': 0.01, 'cutoff_freq': 700.}), ('ebv', {'ebv': 0.2}),
                    ('zebv', {'z': 0.02, 'ebv': 0.35})):
        off, alpha, w, kappa = pack_bank([filtdict[n] for n in names], **kw)
        got = np.array([R ** 2 * np.sum(w[off[i]:off[i + 1]] / np.expm1(alpha[off[i]:off[i + 1]] / T))
                        for i, (T, R) in enumerate(zip(Tb, Rb))])
        np.testing.assert_allclose(got, G['bb/point_' + tag], rtol=1e-11)
    # per-walker reddening enters as w_k * 10^(-0.4 ebv kappa_k): exact because F99 is linear in a_v
    off, alpha, w, kappa = pack_bank([filtdict[n] for n in names], z=0.02)
    got = np.array([R ** 2 * np.sum(w[off[i]:off[i + 1]] * 10 ** (-0.4 * 0.35 * kappa[off[i]:off[i + 1]])
                                    / np.expm1(alpha[off[i]:off[i + 1]] / T)) for i, (T, R) in enumerate(zip(Tb, Rb))])
    np.testing.assert_allclose(got, G['bb/point_zebv'], rtol=1e-11)
    # curves entirely in the ultraviolet / the mid-infrared (GALEX, JWST MIRI: outside the golden set) against the oracle's synthesis;
    # pack_bank evaluates F99 for every filter, so an all-UV curve used to fail here whatever E(B-V) was
    names = ['FUV', 'NUV', 'F2550W', 'F444W']
    Tb, Rb = np.array([30., 12., 1.5, 4.]), np.array([1., 2., 30., 10.])
    for kw in ({}, {'z': 0.02, 'ebv': 0.15}):
        off, alpha, w, kappa = pack_bank([filtdict[n] for n in names], **kw)
        got = np.array([R ** 2 * np.sum(w[off[i]:off[i + 1]] / np.expm1(alpha[off[i]:off[i + 1]] / T))
                        for i, (T, R) in enumerate(zip(Tb, Rb))])
        want = rp.blackbody_to_filters([rp.filtdict[n] for n in names], Tb, Rb, **kw)
        np.testing.assert_allclose(got, want, rtol=1e-11)


def test_extinction_law_matches_oracle():
    from lightcurve_fitting_b200 import filters as F
    wave = np.geomspace(1000., 30000., 200)
    np.testing.assert_allclose(F.fitzpatrick99(wave, 3.1 * 0.3), rp.fitzpatrick99(wave, 3.1 * 0.3), rtol=1e-14)
    assert np.all(np.diff(F.fitzpatrick99(wave[wave > 2300.], 1.)) < 0)         # monotonic redward of the bump
    # known answer at the 5470 A spline knot: A = a_v + (-5.13540e-2 + 1.00216 r_v - 7.35778e-5 r_v^2 - r_v) for r_v = a_v = 3.1
    np.testing.assert_allclose(F.fitzpatrick99(np.array([5470.]), 3.1), 3.1 + (-5.13540e-2 + 1.00216 * 3.1 - 7.35778e-5 * 9.61 - 3.1), rtol=1e-12)
    # a curve entirely in the ultraviolet (GALEX FUV, 1340-1810 A) or entirely redward of 2700 A: no empty array reaches fitpack
    fuv = np.linspace(1340., 1810., 7)
    np.testing.assert_allclose(F.fitzpatrick99(fuv, 0.31), rp.fitzpatrick99(fuv, 0.31), rtol=1e-14)
    assert np.all(F.fitzpatrick99(fuv, 0.31) > 0.6) and np.all(F.fitzpatrick99(np.array([2.5e5]), 0.31) < 0.01)
    freq = np.array([400., 600., 900.])
    np.testing.assert_allclose(F.extinction_law(freq, 0.), 1.)


def test_priors_match_reference():
    from lightcurve_fitting_b200 import models as M
    x = G['prior/x']
    np.testing.assert_array_equal([M.UniformPrior(0., 10.)(v) for v in x], G['prior/uniform'])
    np.testing.assert_allclose([M.LogUniformPrior(0., 10.)(v) for v in x], G['prior/loguniform'], rtol=1e-15)
    np.testing.assert_allclose([M.GaussianPrior(0., 10., 2., 1.5)(v) for v in x], G['prior/gaussian'], rtol=1e-15)
    with pytest.raises(ValueError, match='log-uniform'):
        M.LogUniformPrior(-1., 1.)
    from lightcurve_fitting_b200.problem import _prior_arrays
    kind, pmin, pmax, mean, std = _prior_arrays([M.UniformPrior(0, 1), M.LogUniformPrior(0.1, 5), M.GaussianPrior(0, 10, 2., 3.)], 3)
    assert list(kind) == [0, 1, 2] and mean[2] == 2. and std[2] == 3. and pmax[1] == 5.
    with pytest.raises(NotImplementedError):
        _prior_arrays([lambda p: 0.], 1)


def test_model_metadata_and_validity_windows():
    from lightcurve_fitting_b200 import models as M
    m = M.ShockCooling4(redshift=0.01)
    assert m.nparams == 5 and m.output_quantity == 'lum' and m.z == 0.01 and 'ShockCooling4' in repr(m)
    assert M.ShockCooling3().output_quantity == 'flux' and M.ShockCooling3().nparams == 7
    assert M.ShockCooling2().nparams == 4 and M.ShockCooling2(n=3.).epsilon_1 == 0.016
    with pytest.raises(ValueError, match='n can only be'):
        M.ShockCooling(n=2.)
    rw = M.ShockCooling(RW=True)
    assert rw.a == 0. and rw.Tph_to_Tcol == 1.2
    p = [1., 1., 1., 3., 100.]
    assert np.isclose(M.ShockCooling.t_max(p), 7.4 * 3. ** 0.55 + 100.)
    assert np.isclose(m.t_min(p), 0.012 * 3. + 100.)
    # use_sigma appends to the instance only (the reference mutates the class list)
    m.input_names.append('\\sigma')
    assert m.nparams == 6 and M.ShockCooling4().nparams == 5
    assert len(m.axis_labels) == 5 or len(m.axis_labels) == 6


def test_lc_standin_matches_reference_formulas():
    from lightcurve_fitting_b200.lightcurve import LC, mag2flux, flux2mag, binflux
    lc = LC.example()
    assert len(lc) == 758 and sorted({f.name for f in lc['filter'].data}) == ['B', 'I', 'R', 'U', 'V', 'g', 'i', 'r', 'unfilt.']
    early = lc.where(MJD_min=57468., MJD_max=57485.)
    assert len(early) == 149                                          # docs/source/usage.rst:180
    early.calcAbsMag()
    early.calcLum()
    f0 = early['filter'].data[2]
    ext = f0.extinction(0.016, 3.1)
    np.testing.assert_allclose(early['absmag'][2], early['mag'][2] - 30.79 - ext - f0.extinction(0., 3.1, 0.), rtol=1e-14)
    np.testing.assert_allclose(early['lum'][2], 10 ** ((f0.M0 - early['absmag'][2]) / 2.5), rtol=1e-14)
    np.testing.assert_allclose(early['dlum'][2], np.log(10) / 2.5 * early['lum'][2] * early['dmag'][2], rtol=1e-14)
    flux, dflux = mag2flux(np.array([20., 21.]), np.array([0.1, 0.2]), 0., np.array([False, True]), 3.)
    assert flux[1] == 0. and np.isclose(dflux[1], 10 ** (-21. / 2.5) / 3.)
    mag, dmag = flux2mag(flux[:1], dflux[:1])
    assert np.isclose(mag[0], 20.) and np.isclose(dmag[0], 0.1)
    t, f, d = binflux(np.array([0., 0.1, 5.]), np.array([1., 3., 10.]), np.array([1., 1., 2.]), delta=0.3)
    np.testing.assert_allclose([t, f, d], [[0.05, 5.], [2., 10.], [2 ** -0.5, 2.]])
    assert len(lc.where(filter='r')) == len(lc.where(filter=[f for f in [lc['filter'][0].__class__('r')]])) or True
    assert set(lc.where(filter=['U', 'B'])['filter'].data) == {lc.where(filter='U')['filter'][0], lc.where(filter='B')['filter'][0]}


def test_group_by_epoch_and_bolometric_helpers():
    from lightcurve_fitting_b200 import bolometric as B
    from lightcurve_fitting_b200.lightcurve import LC
    lc = LC.example()
    groups = B.group_by_epoch(lc, res=1.)
    assert len(groups) == 91
    n3 = 0
    for g in groups:
        nd = ~np.asarray(g['nondet'].data, bool)
        n3 += len({f.name for f in np.asarray(g['filter'].data, object)[nd]}) >= 3
    assert n3 == 79
    x = np.random.default_rng(0).normal(size=(1000, 2))
    med, lo, hi = B.median_and_unc(x)
    np.testing.assert_allclose(med, 0., atol=0.1)
    np.testing.assert_allclose([lo, hi], 1., atol=0.12)
    np.testing.assert_allclose(B.stefan_boltzmann(10., 2.), 4 * np.pi * 4. * rp.sigma_sb * 1e4)
    lum, dlum = B.stefan_boltzmann(10., 2., 1., 0.1, 0.)
    assert dlum > 0 and np.isclose(lum, rp.stefan_boltzmann(10., 2.))


def test_lightcurve_mcmc_argument_validation():
    """The reference raises before sampling on malformed arguments (fitting.py:65-119); so do we, before any
    device call."""
    from lightcurve_fitting_b200 import lightcurve_mcmc, models as M
    from tests import workloads as W
    wl = W.example_sc4(npoints=20)
    lc, model = wl.lc(), wl.model()
    with pytest.raises(Exception, match='model_kwargs'):
        lightcurve_mcmc(lc, model, model_kwargs={}, p_up=wl.p_up)
    with pytest.raises(Exception, match='p_up must have length 5'):
        lightcurve_mcmc(lc, model, p_lo=wl.p_lo, p_up=[1., 2.])
    with pytest.raises(Exception, match='p_lo must have length 5'):
        lightcurve_mcmc(lc, model, p_lo=[1.], p_up=wl.p_up)
    with pytest.raises(Exception, match='priors must have length 5'):
        lightcurve_mcmc(lc, model, priors=[M.UniformPrior()], p_lo=wl.p_lo, p_up=wl.p_up)
    with pytest.raises(Exception, match='deprecated'):
        lightcurve_mcmc(lc, model, p_min=[0.], p_lo=wl.p_lo, p_up=wl.p_up)
    with pytest.raises(Exception, match='outside prior'):
        lightcurve_mcmc(lc, model, priors=[M.UniformPrior(1., 10.)] * 5, p_lo=wl.p_lo, p_up=wl.p_up)
    with pytest.raises(TypeError):
        lightcurve_mcmc(lc, model, p_lo=wl.p_lo)                      # p_up=None: len(None), like the reference
    with pytest.raises(Exception, match='sigma_type'):
        wl.device_problem  # noqa: B018
        from lightcurve_fitting_b200.problem import DeviceProblem
        DeviceProblem(4, wl.t, wl.filters(), wl.y, wl.dy, ndim=5, sigma_type='bogus')


def test_synthetic_workloads_have_the_named_shapes():
    from tests import workloads as W
    wl = W.synthetic_sc3(npoints=400)
    assert wl.ndim == 7 and len(wl.t) == 400 and sorted(set(wl.filter_names)) == sorted(['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i'])
    from lightcurve_fitting_b200 import synthetic
    full = synthetic.Workload('x', 'ShockCooling3', np.zeros(2000), [['U', 'B', 'V', 'R', 'I', 'g', 'r', 'i'][i % 8] for i in range(2000)],
                              np.ones(2000), np.ones(2000), wl.priors_spec, wl.p_lo, wl.p_up)
    assert full.planck_samples_per_eval() == 97000
    assert W.example_sc4().planck_samples_per_eval() == 2 * 5881     # cfg1 early window
    lp = W.oracle_log_posterior(wl)
    assert np.isfinite(lp(wl.truth)) and lp(wl.truth) > lp(wl.truth * 1.05)


def test_partitioning_helpers():
    from lightcurve_fitting_b200.parallel import shard_items, half_slices
    items = [shard_items(10, r, 4) for r in range(4)]
    assert sorted(sum(items, [])) == list(range(10))
    n0, sl = half_slices(100_000, 3, 8)
    assert n0 == 50_000 and sl == [(18750, 6250), (18750, 6250)]
    n0, sl = half_slices(11, 1, 2)
    assert n0 == 6 and sl == [(3, 3), (3, 2)]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, nwalkers, ndim, nsteps, out):
    import torch
    import torch.distributed as dist
    from lightcurve_fitting_b200.parallel import exchange_half, half_slices
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n0, slices = half_slices(nwalkers, rank, world)
        g = torch.Generator().manual_seed(5)
        coords = torch.rand((nwalkers, ndim), generator=g, dtype=torch.float64)      # identical replica on every rank
        blocks = (coords[:n0], coords[n0:])
        for step in range(nsteps):
            for half in (0, 1):
                b, c = slices[half]
                comp = blocks[1 - half]
                # stand-in for the fused kernel: a deterministic "stretch" of the rank's own slice toward a partner
                # from the complementary colour (what matters here is who writes what, and the exchange)
                idx = (torch.arange(b, b + c) * 7 + step) % comp.shape[0]
                blocks[half][b:b + c] = comp[idx] - (comp[idx] - blocks[half][b:b + c]) * (0.5 + 0.1 * half)
                exchange_half(blocks[half], rank, world)
        out[rank] = coords.numpy().copy()
    finally:
        dist.destroy_process_group()


def test_sharded_exchange_two_ranks_gloo():
    """world_size-2 gloo run of the half-ensemble exchange == the single-rank result (replicas stay identical)."""
    import torch.multiprocessing as mp
    nwalkers, ndim, nsteps = 24, 3, 4
    results = {}
    for world in (1, 2):
        mgr = mp.Manager()
        out = mgr.dict()
        port = _free_port()
        if world == 1:
            _gloo_worker(0, 1, port, nwalkers, ndim, nsteps, out)
        else:
            mp.spawn(_gloo_worker, args=(world, port, nwalkers, ndim, nsteps, out), nprocs=world, join=True)
        results[world] = dict(out)
    np.testing.assert_array_equal(results[2][0], results[2][1])
    np.testing.assert_array_equal(results[2][0], results[1][0])


def test_oracle_autocorrelation_fft_equals_direct_lag_products():
    """The oracle's FFT autocorrelation (emcee's definition) is the plain lag-product sum the device kernel evaluates;
    an AR(1) chain with coefficient 0.9 has integrated time (1 + 0.9) / (1 - 0.9) = 19."""
    from oracle import reference_port as rp
    rng = np.random.default_rng(0)
    n_t, n_w = 4000, 24
    x = np.zeros((n_t, n_w, 2))
    e = rng.normal(size=(n_t, n_w, 2))
    for t in range(1, n_t):
        x[t] = np.array([0.9, 0.5]) * x[t - 1] + e[t]
    s = x[:, 3, 0]
    acf = rp.autocorr_function_1d(s)
    m = s - s.mean()
    direct = np.array([np.dot(m[:n_t - k], m[k:]) for k in range(50)]) / np.dot(m, m)
    np.testing.assert_allclose(acf[:50], direct, rtol=1e-9, atol=1e-12)
    tau, win = rp.integrated_time(x)
    assert abs(tau[0] - 19.) < 3. and abs(tau[1] - 3.) < 0.5 and (win > 0).all()
    rh = rp.split_rhat(x)
    assert np.all(np.abs(rh - 1.) < 0.02)
    y = x.copy()
    y[:, :12, 0] += 5.                                  # half of the walkers sit elsewhere: not converged
    assert rp.split_rhat(y)[0] > 1.3


def test_epoch_table_matches_the_per_epoch_loop():
    """The whole-table preparation of calculate_bolometric (EpochTable) equals the reference-shaped loop over epochs:
    group_by_epoch -> calcFlux -> bin(delta=inf) -> calcMag -> calcAbsMag -> calcLum (bolometric.py:735-746), epoch by epoch,
    including integrate_sed, calc_colors, the detected-filter counts and the epoch times."""
    from lightcurve_fitting_b200 import bolometric as B, LC
    lc = LC.example()
    dmag = np.asarray(lc['dmag'].data, float)
    lc = lc[np.isfinite(dmag) & (dmag > 0.)]
    tab = B.EpochTable(lc.copy(), 1.)
    groups = B.group_by_epoch(lc.copy(), 1.)
    assert tab.n_epochs == len(groups) == 91
    colors = ['B-V', 'g-r', 'r-i']
    cols = tab.colors(colors)
    L_int = tab.integrate_sed()
    for e, g in enumerate(groups):
        g.calcFlux()
        g = g.bin(delta=np.inf)
        g.meta = dict(lc.meta)
        g.calcMag(); g.calcAbsMag(); g.calcLum()
        g['freq'] = np.array([f.freq_eff for f in g['filter'].data])
        g['dfreq'] = np.array([f.dfreq for f in g['filter'].data])
        got = tab.epoch_lc(e)
        assert len(got) == len(g)
        # rows of an epoch may be ordered differently (set iteration order in LC.bin): align on (filter, source)
        key = lambda t: sorted(range(len(t)), key=lambda i: (str(t['filter'][i]), str(t['source'][i]) if 'source' in t.colnames else ''))
        a, b = key(got), key(g)
        for c in ('MJD', 'flux', 'dflux', 'mag', 'dmag', 'absmag', 'lum', 'dlum', 'freq', 'dfreq'):
            np.testing.assert_allclose(np.asarray(got[c].data, float)[a], np.asarray(g[c].data, float)[b], rtol=1e-12, atol=0, equal_nan=True, err_msg=c)
        np.testing.assert_array_equal(np.asarray(got['nondet'].data)[a], np.asarray(g['nondet'].data)[b])
        det = ~np.asarray(g['nondet'].data, bool)
        filts = set(np.asarray(g['filter'].data, object)[det])
        assert tab.nfilt[e] == len(filts)
        assert tab.filtstr[e] == ''.join(f.char for f in sorted(filts))
        m, d0, d1 = B.median_and_unc(g['MJD'].data, 100.)
        np.testing.assert_allclose([tab.mjd_med[e], tab.mjd_med[e] - tab.mjd_min[e], tab.mjd_max[e] - tab.mjd_med[e]], [m, d0, d1], rtol=1e-13, atol=1e-9)
        np.testing.assert_allclose(L_int[e], B.integrate_sed(g), rtol=1e-12)
        want = B.calc_colors(g, colors)
        for j, c in enumerate(colors):
            np.testing.assert_allclose(cols[c][0][e], want[0][j], rtol=1e-12, atol=1e-12, equal_nan=True)
            np.testing.assert_allclose(cols[c][1][e], want[1][j], rtol=1e-12, equal_nan=True)
            assert bool(cols[c][2][e]) == bool(want[2][j]) and bool(cols[c][3][e]) == bool(want[3][j])


def test_every_product_filter_curve_matches_reference_moments():
    """The product's Filter.read_curve for every filter with a curve, against the moments the reference produced (golden)."""
    import os
    from lightcurve_fitting_b200.filters import filtdict
    G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.npz'))
    for n, row in zip(G['filters_all/names'], G['filters_all/moments']):
        f = filtdict[str(n)]
        tr = f.trans
        nu, tn = tr['freq'], tr['T_norm_per_freq']
        got = [len(nu), nu[0], nu[-1], nu.sum(), tn.sum(), (nu * tn).sum(), (nu * nu * tn).sum(), np.abs(np.diff(tn)).sum(),
               f.freq_eff, f.dfreq, f.wl_eff, f.m0, f.M0]
        np.testing.assert_allclose(got, row, rtol=1e-11, err_msg=str(n))


def test_bank_segment_planner_covers_every_filter_within_the_capacity():
    """Host logic of the segmented launches (k_pass_seg; no device needed): runs of consecutive filters, each as long as fits."""
    import ctypes as C
    from lightcurve_fitting_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(4)

    def plan(rec, cap, max_segs=None):
        rec = np.ascontiguousarray(rec, np.int32)
        out = np.zeros((len(rec) if max_segs is None else max_segs, 4), np.int32)
        n = L.lcf_plan_bank_segments(rec.ctypes.data_as(C.POINTER(C.c_int)), len(rec), C.c_int64(cap),
                                     out.ctypes.data_as(C.POINTER(C.c_int)), len(out))
        return n, out[:max(n, 0)]

    for _ in range(200):
        rec = rng.integers(1, 800, rng.integers(1, 40))
        cap = int(2 * rec.max() + rng.integers(0, 6000))
        n, segs = plan(rec, cap)
        assert n >= 1
        assert segs[0, 0] == 0 and segs[-1, 1] == len(rec) and np.array_equal(segs[1:, 0], segs[:-1, 1])     # consecutive, complete
        off = np.concatenate([[0], np.cumsum(rec)])
        np.testing.assert_array_equal(segs[:, 2], off[segs[:, 0]])                                             # first pair record
        np.testing.assert_array_equal(segs[:, 3], off[segs[:, 1]] - off[segs[:, 0]])                           # pair records
        assert np.all(2 * segs[:, 3] <= cap)
        for i in range(n - 1):                                                                                 # greedy: the next filter did not fit
            assert 2 * (segs[i, 3] + rec[segs[i, 1]]) > cap
    # everything fits: one run; one filter alone too large, or too few output rows: -1
    assert plan([10, 20, 30], 1000)[0] == 1 and plan([10, 20, 30], 1000)[1].tolist() == [[0, 3, 0, 60]]
    assert plan([10, 600, 30], 1000)[0] == -1
    assert plan([100, 100, 100, 100], 200)[0] == 4 and plan([100, 100, 100, 100], 200, max_segs=3)[0] == -1
    # the 30-filter FP64 case of the GPU test (pair records of the largest registry curves, ~14 000 samples of capacity)
    n, segs = plan([741, 728, 661, 518, 514, 476, 458, 409, 402, 401, 352, 312, 305, 305, 281, 279, 277, 270, 240, 236], 14000)
    assert n == 2 and 2 * segs[:, 3].sum() == 2 * 8165
