#!/usr/bin/env python
"""Generate the golden vectors in this directory by executing the REFERENCE's own source files.

The reference cannot be imported as a package in this image: astropy, asteval, lmfit, emcee,
pathos, corner and matplotlib are absent.  This script therefore puts minimal stand-ins for those
third-party modules on ``sys.path`` (``ref_shims/``, see its README) and loads, unmodified and from
where they lie under ``/root/reference``, exactly the files of the hot path:

    mcmc_dynamics/parameter.py                      mcmc_dynamics/config/*.json
    mcmc_dynamics/utils/coordinates/calc_xy_offset.py, get_amplitude_and_angle.py
    mcmc_dynamics/utils/files/data_reader.py
    mcmc_dynamics/background/gaussian.py, single_stars.py
    mcmc_dynamics/analysis/runner.py, constant.py, model.py

It then builds the reference's model classes on small seeded catalogues, evaluates
``lnprior / lnlike / lnprob`` (and the background callables, ``calc_xy_offset``, ``no_sum``) and
writes inputs and outputs to ``golden_reference.json``.  The committed JSON is what the tests read;
this script only runs where ``/root/reference`` exists:

    python tests/golden/make_golden.py

What the vectors pin: the reference's arithmetic and control flow (its files are executed as they
are).  What they do not pin: real astropy's unit handling, which the stand-in restates.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/mcmc_dynamics'
OUT = os.path.join(HERE, 'golden_reference.json')


def load_reference():
    sys.path.insert(0, os.path.join(HERE, 'ref_shims'))

    def package(name, path):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        mod.__package__ = name
        sys.modules[name] = mod
        return mod

    def load(name, path, search=None):
        spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=search)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        parent, _, leaf = name.rpartition('.')
        if parent:
            setattr(sys.modules[parent], leaf, mod)
        return mod

    root = package('mcmc_dynamics', REF)
    for sub in ('analysis', 'background', 'utils', 'utils/coordinates', 'utils/files'):
        package('mcmc_dynamics.' + sub.replace('/', '.'), os.path.join(REF, sub))
        parent, _, leaf = ('mcmc_dynamics.' + sub.replace('/', '.')).rpartition('.')
        setattr(sys.modules[parent], leaf, sys.modules['mcmc_dynamics.' + sub.replace('/', '.')])
    load('mcmc_dynamics.config', os.path.join(REF, 'config', '__init__.py'), [os.path.join(REF, 'config')])
    parameter = load('mcmc_dynamics.parameter', os.path.join(REF, 'parameter.py'))
    root.Parameters, root.Parameter = parameter.Parameters, parameter.Parameter
    coords = sys.modules['mcmc_dynamics.utils.coordinates']
    coords.calc_xy_offset = load('mcmc_dynamics.utils.coordinates.calc_xy_offset',
                                 os.path.join(REF, 'utils/coordinates/calc_xy_offset.py')).calc_xy_offset
    coords.get_amplitude_and_angle = load('mcmc_dynamics.utils.coordinates.get_amplitude_and_angle',
                                          os.path.join(REF, 'utils/coordinates/get_amplitude_and_angle.py')
                                          ).get_amplitude_and_angle
    files = sys.modules['mcmc_dynamics.utils.files']
    files.DataReader = load('mcmc_dynamics.utils.files.data_reader',
                            os.path.join(REF, 'utils/files/data_reader.py')).DataReader
    bg = sys.modules['mcmc_dynamics.background']
    bg.Gaussian = load('mcmc_dynamics.background.gaussian', os.path.join(REF, 'background/gaussian.py')).Gaussian
    bg.SingleStars = load('mcmc_dynamics.background.single_stars',
                          os.path.join(REF, 'background/single_stars.py')).SingleStars
    load('mcmc_dynamics.analysis.runner', os.path.join(REF, 'analysis/runner.py'))
    constant = load('mcmc_dynamics.analysis.constant', os.path.join(REF, 'analysis/constant.py'))
    model = load('mcmc_dynamics.analysis.model', os.path.join(REF, 'analysis/model.py'))
    from astropy import units as u
    return types.SimpleNamespace(
        u=u, Parameters=parameter.Parameters, DataReader=files.DataReader, Gaussian=bg.Gaussian,
        SingleStars=bg.SingleStars, calc_xy_offset=coords.calc_xy_offset, ConstantFit=constant.ConstantFit,
        ConstantFitGB=constant.ConstantFitGB, ModelFit=model.ModelFit, ModelFitGB=model.ModelFitGB,
        ModelFitConstantBackground=model.ModelFitConstantBackground)


def catalogue(n, seed, contaminated=False):
    """Seeded catalogue from the PRODUCT's synthetic generator (inputs only; they are stored)."""
    sys.path.insert(0, ROOT)
    from mcmc_dynamics_b200 import synthetic
    columns, truth = synthetic.mock_cluster(n, seed=seed, as_reader=False)
    v_bg = None
    if contaminated:
        columns, sample_field = synthetic.add_background(columns, truth, seed=seed + 100)
        v_bg = sample_field(40, seed=seed + 200)
    truth = {k: float(v) for k, v in truth.items() if np.ndim(v) == 0}
    truth.update(v_back=5.0, sigma_back=55.0, f_back=0.3)
    return columns, truth, v_bg


def thetas(truth, names, n, seed, scale=0.2):
    sys.path.insert(0, ROOT)
    from mcmc_dynamics_b200 import synthetic
    return synthetic.initial_ball(truth, names, n, seed=seed, scale=scale)


def main():
    R = load_reference()
    u = R.u
    cases = []

    def reader(columns):
        units = {'ra': u.deg, 'dec': u.deg, 'v': u.km / u.s, 'verr': u.km / u.s}
        return R.DataReader({k: (u.Quantity(v, units[k]) if k in units else u.Quantity(v)) for k, v in columns.items()})

    def run_case(name, cls_name, columns, truth, free_centre=False, background=None, v_bg=None, edits=None,
                 n_theta=6, theta_edits=None, no_sum=False):
        cls = getattr(R, cls_name)
        data = reader(columns)
        bg_obj = None
        bg_desc = None
        if background == 'single_stars':
            bg_obj = R.SingleStars(u.Quantity(v_bg, u.km / u.s))
            bg_desc = {'kind': 'single_stars', 'v_bg': list(map(float, v_bg))}
        elif background == 'gaussian':
            bg_obj = R.Gaussian(5.0 * u.km / u.s, 55.0 * u.km / u.s)
            bg_desc = {'kind': 'gaussian', 'mean': 5.0, 'sigma': 55.0}
        if cls_name == 'ModelFitConstantBackground':
            model = cls(data, background=bg_obj)
        elif bg_obj is not None:
            model = cls(data, background=bg_obj)
        else:
            model = cls(data)
        sets = {}
        if not free_centre:
            sets['ra_center'] = dict(value=truth['ra_center'], fixed=True)
            sets['dec_center'] = dict(value=truth['dec_center'], fixed=True)
        else:
            sets['ra_center'] = dict(value=truth['ra_center'])
            sets['dec_center'] = dict(value=truth['dec_center'])
        for pname, kw in (edits or {}).items():
            sets.setdefault(pname, {}).update(kw)
        for pname, kw in sets.items():
            kw = dict(kw)
            if 'value' in kw:
                kw['value'] = u.Quantity(kw['value'], model.parameters[pname].unit) \
                    if model.parameters[pname].unit is not None else kw['value']
            model.parameters[pname].set(**kw)
        names = model.fitted_parameters
        theta = thetas(truth, names, n_theta, seed=len(cases) + 1)
        for (row, pname, value) in (theta_edits or []):
            theta[row, names.index(pname)] = value
        # the model curves themselves at theta[0], routed exactly as <Model>.lnlike routes them; evaluated first
        # because fetch_parameter_values writes the values into model.parameters (analysis/runner.py:163-176)
        # and the descriptor recorded below must see the state the likelihood calls leave behind
        # (constant.py:137-151, model.py:205-220)
        par0 = model.fetch_parameter_values(theta[0])
        with np.errstate(all='ignore'):
            v_curve = model.rotation_model(**{k: v for k, v in par0.items() if k in model.rotation_parameters.keys()})
            s_curve = model.dispersion_model(**{k: v for k, v in par0.items() if k in model.dispersion_parameters.keys()})
        kms = u.km / u.s
        curves = {'v_los': [float(x) for x in np.asarray(u.Quantity(v_curve).to(kms).value)],
                  'sigma_los': [float(x) for x in np.asarray(u.Quantity(s_curve).to(kms).value)]}
        out = {'lnprior': [], 'lnlike': [], 'lnprob': []}
        with np.errstate(all='ignore'):
            for row in theta:
                lp = model.lnprior(row)
                out['lnprior'].append(float(lp))
                ll = model.lnlike(row)
                out['lnlike'].append(float(getattr(ll, 'value', ll)))
                pr = model.lnprob(row)
                out['lnprob'].append(float(getattr(pr, 'value', pr)))
            per_star = None
            if no_sum:
                per_star = model.lnlike(theta[0], no_sum=True)
                per_star = [float(x) for x in np.asarray(getattr(per_star, 'value', per_star))]
        membership = None
        if hasattr(model, 'calculate_membership_probabilities') and cls_name != 'ConstantFit' and cls_name != 'ModelFit':
            # the reference evaluates the membership at the posterior median of a chain: hand it a
            # chain whose every sample is theta[0] and a minimal stand-in for its indexed best-fit table
            class _BestFit(object):
                def __init__(self, names, values, units):
                    self.columns = ['value'] + list(names)
                    self.loc = {'median': dict([('value', 'median')] + [
                        (n_, (u.Quantity(v_, un) if un is not None else v_)) for n_, v_, un in zip(names, values, units)])}
            units = [model.parameters[n_].unit for n_ in names]
            model.compute_bestfit_values = lambda chain, n_burn: _BestFit(names, theta[0], units)
            if cls_name == 'ConstantFitGB':
                # constant.py:370 hands only the fitted medians on; its helper needs the fixed centre too
                orig = model._calculate_lnlike_cluster_back
                fixed = {n_: u.Quantity(p_.value, p_.unit) for n_, p_ in model.parameters.items() if p_.fixed}
                model._calculate_lnlike_cluster_back = lambda pars: orig(dict(pars, **fixed))
            with np.errstate(all='ignore'):
                pm = model.calculate_membership_probabilities(chain=None, n_burn=0)
            membership = [float(x) for x in np.asarray(getattr(pm, 'value', pm))]
        case = {
            'name': name, 'class': cls_name, 'free_centre': free_centre, 'background': bg_desc,
            'columns': {k: [float(x) for x in v] for k, v in columns.items()},
            'parameter_edits': {k: {kk: (float(vv) if isinstance(vv, (int, float, np.floating)) and not isinstance(vv, bool) else vv)
                                    for kk, vv in kw.items()} for k, kw in sets.items()},
            'fitted_parameters': names, 'theta': theta.tolist(), 'expected': out,
        }
        # what the INTEGRATION.md Option-B binding (tools/reference_binding.py) compiles from this REAL
        # reference object: tests/test_binding_cpu.py checks it against the product's own packing, and
        # tests/test_gpu_binding.py runs the binding on a stand-in with the same attribute surface
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import reference_binding
        case['binding_descriptor'] = reference_binding.summary(reference_binding.describe(model)[0])
        case['model_parameters'] = list(model.MODEL_PARAMETERS)
        case['model_curves_theta0'] = curves
        if model.lnlike_background is not None:
            lbg = model.lnlike_background
            case['lnlike_background'] = [float(x) for x in np.asarray(getattr(lbg, 'value', lbg))]
        if per_star is not None:
            case['lnlike_per_star_theta0'] = per_star
        if membership is not None:
            case['membership_theta0'] = membership
        cases.append(case)
        print('%-44s lnprob[0] = %.12g' % (name, out['lnprob'][0]))

    n = 160
    cols, truth, _ = catalogue(n, seed=31)
    run_case('ConstantFit fixed centre', 'ConstantFit', cols, truth,
             theta_edits=[(5, 'sigma_max', -0.5)])
    run_case('ConstantFit free centre', 'ConstantFit', cols, truth, free_centre=True, n_theta=5)
    run_case('ModelFit fixed centre', 'ModelFit', cols, truth, theta_edits=[(5, 'a', -1.0)])
    run_case('ModelFit free centre', 'ModelFit', cols, truth, free_centre=True, n_theta=5)
    run_case('ModelFit v_sys fixed, bounded', 'ModelFit', cols, truth, n_theta=5,
             edits={'v_sys': dict(value=0.25, fixed=True), 'sigma_max': dict(min=0.0, max=100.0),
                    'r_peak': dict(min=0.0, max=500.0)})
    colsb, truthb, v_bg = catalogue(n, seed=32, contaminated=True)
    run_case('ConstantFit + SingleStars background', 'ConstantFit', colsb, truthb, background='single_stars', v_bg=v_bg)
    run_case('ModelFit + Gaussian background, free centre', 'ModelFit', colsb, truthb, free_centre=True,
             background='gaussian', n_theta=5)
    run_case('ConstantFitGB', 'ConstantFitGB', colsb, truthb)
    run_case('ModelFitGB free centre', 'ModelFitGB', colsb, truthb, free_centre=True, n_theta=5)
    run_case('ModelFitConstantBackground', 'ModelFitConstantBackground', colsb, truthb, background='single_stars',
             v_bg=v_bg, no_sum=True)

    # geometry and the background callables on their own
    dx, dy = R.calc_xy_offset(u.Quantity(cols['ra'], u.deg), u.Quantity(cols['dec'], u.deg),
                              truth['ra_center'] * u.deg, truth['dec_center'] * u.deg)
    extras = {
        'calc_xy_offset': {'ra': [float(x) for x in cols['ra']], 'dec': [float(x) for x in cols['dec']],
                           'ra_center': truth['ra_center'], 'dec_center': truth['dec_center'],
                           'dx_arcmin': [float(x) for x in dx.to(u.arcmin).value],
                           'dy_arcmin': [float(x) for x in dy.to(u.arcmin).value]},
    }
    v = u.Quantity(colsb['v'], u.km / u.s)
    verr = u.Quantity(colsb['verr'], u.km / u.s)
    ss = R.SingleStars(u.Quantity(v_bg, u.km / u.s))
    extras['single_stars'] = {'v_bg': [float(x) for x in v_bg], 'v': [float(x) for x in colsb['v']],
                              'verr': [float(x) for x in colsb['verr']],
                              'lnlike': [float(x) for x in np.asarray(ss(v, verr).value)],
                              'lnlike_sigma_int_3': [float(x) for x in np.asarray(ss(v, verr, sigma_int=3.0 * u.km / u.s).value)]}
    g = R.Gaussian(5.0 * u.km / u.s, 55.0 * u.km / u.s)
    extras['gaussian'] = {'mean': 5.0, 'sigma': 55.0, 'lnlike': [float(x) for x in np.asarray(g(v, verr).value)]}

    # ---- the callers either side of the path: radial binning and chain post-processing -------------
    post = {}
    binning = []
    for nstars, dlogr in ((50, 0.1), (30, 0.2), (200, 0.05)):
        cb, tb, _ = catalogue(700, seed=41)
        rd = reader(cb)
        rd.make_radial_bins(tb['ra_center'] * u.deg, tb['dec_center'] * u.deg, nstars=nstars, dlogr=dlogr)
        binning.append({'n_stars': 700, 'seed': 41, 'nstars': nstars, 'dlogr': dlogr,
                        'labels': [int(x) for x in np.asarray(rd.data['bin'])]})
    post['make_radial_bins'] = binning
    # a synthetic "chain" [walkers, steps, parameters] for a ConstantFit with v_sys fixed
    rng = np.random.default_rng(77)
    cb, tb, _ = catalogue(60, seed=42)
    cf = R.ConstantFit(reader(cb))
    cf.parameters['ra_center'].set(value=u.Quantity(tb['ra_center'], u.deg), fixed=True)
    cf.parameters['dec_center'].set(value=u.Quantity(tb['dec_center'], u.deg), fixed=True)
    cf.parameters['v_sys'].set(value=u.Quantity(1.5, u.km / u.s), fixed=True)
    names = cf.fitted_parameters                         # sigma_max, v_maxx, v_maxy
    chain = np.stack([8.0 + rng.normal(0, 0.7, (12, 40)), -3.0 + rng.normal(0, 1.2, (12, 40)),
                      -0.4 + rng.normal(0, 1.0, (12, 40))], axis=2)
    pct = cf.compute_percentiles(chain, n_burn=10)
    best = cf.compute_bestfit_values(chain, n_burn=10)
    conv = cf.convert_to_parameters(chain, n_burn=10)
    res = cf.compute_theta_vmax(chain, n_burn=10)

    def _val(x):
        return float(getattr(x, 'value', x))
    post['chain_case'] = {
        'fitted_parameters': names, 'chain': chain.tolist(), 'n_burn': 10,
        'percentiles': np.asarray(pct).tolist(),
        'bestfit': {n_: [_val(best.loc[row][n_]) for row in ('median', 'uperr', 'loerr')] for n_ in names},
        'convert_to_parameters': {k: [float(x) for x in np.asarray(getattr(v_, 'value', v_))[:5]] for k, v_ in conv.items()},
        'convert_sizes': {k: int(np.size(v_)) for k, v_ in conv.items()},
        'theta_vmax': {n_: [_val(res.loc[row][n_]) for row in ('median', 'uperr', 'loerr')] for n_ in ('v_max', 'theta_0')},
    }

    # ModelFit.create_profiles on a synthetic chain (model.py:225-317)
    mfp = R.ModelFit(reader(cb))
    mfp.parameters['ra_center'].set(value=u.Quantity(tb['ra_center'], u.deg), fixed=True)
    mfp.parameters['dec_center'].set(value=u.Quantity(tb['dec_center'], u.deg), fixed=True)
    mnames = mfp.fitted_parameters              # v_sys, sigma_max, a, v_maxx, v_maxy, r_peak
    centre = {'v_sys': 0.2, 'sigma_max': 9.0, 'a': 35.0, 'v_maxx': 2.5, 'v_maxy': -3.0, 'r_peak': 70.0}
    mchain = np.stack([centre[n_] * (1.0 + 0.05 * rng.standard_normal((10, 30))) for n_ in mnames], axis=2)
    prof = mfp.create_profiles(mchain, n_burn=5, radii=u.Quantity([1.0, 10.0, 60.0, 200.0], u.arcsec))
    post['create_profiles'] = {
        'fitted_parameters': mnames, 'chain': mchain.tolist(), 'n_burn': 5, 'radii_arcsec': [1.0, 10.0, 60.0, 200.0],
        'columns': {name: [float(x) for x in np.asarray(getattr(col, 'value', col))] for name, col in prof.columns.items()},
    }

    # ---- config format, both directions -----------------------------------------------------------
    # (i) a Parameters object edited and serialised by the REFERENCE; the product must load it
    rp = R.ModelFit.default_parameters()
    rp['ra_center'].set(value=u.Quantity(201.697, u.deg), fixed=True)
    rp['a'].set(value=u.Quantity(0.5, u.arcmin), min=0.0, max=120.0)
    rp['v_sys'].set(value=u.Quantity(232.5, u.km / u.s), fixed=True, lnprior='norm.logpdf(val, loc=232.5, scale=2)')
    rp['sigma_max'].set(initials='rng.lognormal(mean=2.3, sigma=0.5, size=n)')
    post['parameters_json_from_reference'] = rp.dumps()
    post['parameters_expected'] = [[name, float(p.value), None if p.unit is None else str(p.unit), bool(p.fixed),
                                    float(p.min), float(p.max), p.initials, p.lnprior] for name, p in rp.items()]
    # (ii) a string serialised by the PRODUCT; the reference must load it (checked here, at generation time)
    sys.path.insert(0, ROOT)
    from mcmc_dynamics_b200.analysis import ModelFit as ProductModelFit
    pp = ProductModelFit.default_parameters()
    pp['dec_center'].set(value=-47.4799, fixed=True)
    pp['r_peak'].set(value=90.0, min=0.0, max=500.0)
    back = R.Parameters().loads(pp.dumps())
    for name, p in pp.items():
        q = back[name]
        assert float(q.value) == float(p.value) and bool(q.fixed) == bool(p.fixed), name
        assert float(q.min) == float(p.min) and float(q.max) == float(p.max), name
        assert (q.unit is None and p.unit is None) or str(q.unit).replace(' ', '') == str(p.unit).replace(' ', ''), name
        assert q.initials == p.initials, name
    print('product -> reference Parameters JSON round trip ok')

    # default parameter tables as the reference's Parameters class loads them
    tables = {}
    for cls_name in ('ConstantFit', 'ConstantFitGB', 'ModelFit', 'ModelFitGB', 'ModelFitConstantBackground'):
        pars = getattr(R, cls_name).default_parameters()
        tables[cls_name] = [[name, float(p.value), None if p.unit is None else str(p.unit), bool(p.fixed),
                             float(p.min), float(p.max), p.initials] for name, p in pars.items()]
    doc = {'generator': 'tests/golden/make_golden.py', 'reference': 'skamann/mcmc-dynamics (files listed in the generator)',
           'cases': cases, 'extras': extras, 'default_parameters': tables, 'post': post}
    with open(OUT, 'w') as f:
        json.dump(doc, f)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
