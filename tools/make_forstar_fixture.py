#!/usr/bin/env python
"""Independent evaluation of ``Hyperparameters.for_star`` for a few stars -> tests/golden/for_star.json.

TEST INFRASTRUCTURE.  Runs in the build container (reads the reference's data files under
/root/reference; the fixture it writes is what travels).  It imports neither ``gadfly_b200`` nor
``oracle``: the scaling relations are evaluated here a second time, scalar by scalar in mpmath at
40 digits, from the formulas of the reference (file:line cited at every relation), so that the
product's numpy implementation (gadfly_b200/core.py, scale.py, sun.py, feeder.py) is pinned by
numbers it did not produce.  Stars: the Sun, KIC 9333184 (docs/gadfly/synth.rst:44-50), and a
subgiant / a hot dwarf from the ranges of the Huber-2011 table (notebooks/huber2011.ecsv).
"""
import json
import os
import sys

import mpmath as mp

mp.mp.dps = 40
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(HERE, "tests", "golden", "for_star.json")

T_SUN = mp.mpf(5777)          # gadfly/scale.py:20
NUMAX_SUN = mp.mpf(3090)      # gadfly/scale.py:27
DNU_SUN = mp.mpf("135.1")     # gadfly/scale.py:28


def solar_inputs():
    """data/hyperparameters.json: 5 granulation (S0, w0, Q) + 4 per-degree p-mode (S0, Q);
    data/broomhall2009_table2_labeled.ecsv: 81 (nu, degree)."""
    with open(os.path.join(REF, "gadfly", "data", "hyperparameters.json")) as fh:
        hp = json.load(fh)
    gran = [(mp.mpf(repr(p["hyperparameters"]["S0"])), mp.mpf(repr(p["hyperparameters"]["w0"])),
             mp.mpf(repr(p["hyperparameters"]["Q"])))
            for p in hp if p["metadata"]["source"] == "granulation"]
    # gadfly/core.py:146-152: oscillation entries sorted by degree
    osc = sorted((p for p in hp if p["metadata"]["source"] == "oscillation"),
                 key=lambda p: p["metadata"]["degree"])
    S0_ell = [mp.mpf(repr(p["hyperparameters"]["S0"])) for p in osc]
    Q_ell = [mp.mpf(repr(p["hyperparameters"]["Q"])) for p in osc]
    modes = []
    with open(os.path.join(REF, "gadfly", "data", "broomhall2009_table2_labeled.ecsv")) as fh:
        header = False
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            if not header:
                header = True
                continue
            nu, ell = line.split()
            modes.append((mp.mpf(nu), int(ell)))
    return gran, S0_ell, Q_ell, modes


def sho_psd(omega, S0, w0, Q):
    # gadfly/core.py:33-41
    return mp.sqrt(2 / mp.pi) * S0 * w0 ** 4 / ((omega ** 2 - w0 ** 2) ** 2 + omega ** 2 * w0 ** 2 / Q ** 2)


def c_K(T):
    return (T / 5934) ** mp.mpf("0.8")                      # gadfly/scale.py:50-73


def amp_huber(M, T, L):
    # gadfly/scale.py:83-107 with (r, s, t) = (2, 0.886, 1.89), scale.py:40-42
    return L ** mp.mpf("0.886") / (M ** mp.mpf("1.89") * T ** (2 - 1) * c_K(T))


def voigt(x, x0, amplitude_L, fwhm_L, fwhm_G):
    # astropy.modeling Voigt1D as used at gadfly/scale.py:511,538: Faddeeva function w(z)
    s = mp.sqrt(mp.log(2))
    z = (2 * (x - x0) + 1j * fwhm_L) * s / fwhm_G
    w = mp.exp(-z * z) * mp.erfc(-1j * z)
    return mp.re(w) * mp.sqrt(mp.log(2) * mp.pi) / fwhm_G * fwhm_L * amplitude_L


def v_osc_kiefer(freq, numax, dnu):
    # gadfly/scale.py:515-539 (Kiefer et al. 2018)
    w = DNU_SUN / dnu
    sigma, gamma, Sigma = mp.mpf("181.8") / w, mp.mpf("150.9") / w, mp.mpf("611.8") / w
    S, a, b = mp.mpf("-0.1"), mp.mpf(3299) * 10 ** 4, mp.mpf(-581)
    A = 1 / mp.pi * (mp.atan(S * (freq - numax) / Sigma) + mp.mpf("0.5"))
    return A * (b + voigt(freq, numax, a, 2 * gamma, mp.mpf("2.355") * sigma)) * mp.mpf("1e-6")


def p_mode_intensity(T, freq, numax, dnu, wavelength_nm):
    # gadfly/scale.py:579-632: the velocity -> intensity factor cancels in the ratio
    conv = lambda v: mp.mpf("20.1") * (v / (wavelength_nm / 550) / (T / 5777) ** 2)
    return conv(v_osc_kiefer(freq, numax, dnu)) / conv(v_osc_kiefer(numax, numax, dnu))


def for_star(M, R, T, L, alpha=1):
    """gadfly/core.py:107-333, flat SOHO VIRGO bandpass (alpha = 1, mean wavelength 550 nm)."""
    M, R, T, L, alpha = (mp.mpf(repr(v)) for v in (M, R, T, L, alpha))
    gran, S0_ell, Q_ell, modes = solar_inputs()
    out = []
    # core.py:175,186,190: scale factors
    numax = NUMAX_SUN * (M * R ** -2 * (T / T_SUN) ** mp.mpf("-0.5"))            # scale.py:201-226
    gran_amp = (L ** 2 / (M ** 3 * T ** mp.mpf("5.5"))) / (1 / T_SUN ** mp.mpf("5.5"))   # scale.py:458-484
    tau = (L / (M * T ** mp.mpf("3.5"))) / (1 / T_SUN ** mp.mpf("3.5"))          # scale.py:382-406
    for S0, w0, Q in gran:                                                       # core.py:196-228
        w = w0 / tau
        if w > 0:
            out.append((S0 * gran_amp * alpha, w, Q))
    s_dnu = M ** mp.mpf("0.5") * R ** mp.mpf("-1.5")                            # scale.py:176-198
    amp = amp_huber(M, T, L) / amp_huber(1, T_SUN, 1)
    Gamma_scaled = mp.mpf("1.02") * mp.exp((T - T_SUN) / 436)                    # core.py:280
    for nu, ell in modes:                                                        # core.py:236-331
        S0f, Qf = S0_ell[ell], Q_ell[ell]
        w0_sun = 2 * mp.pi * nu
        bg = sum(sho_psd(w0_sun, gS0, gw0, gQ) for gS0, gw0, gQ in gran) * alpha  # core.py:236-241
        nu_s = numax + (nu - NUMAX_SUN) * s_dnu                                  # core.py:243-247
        w0_s = 2 * mp.pi * nu_s
        if not w0_s > 0:                                                         # core.py:250-255
            continue
        fac = p_mode_intensity(T, nu_s, numax, DNU_SUN * s_dnu, mp.mpf(550)) * amp   # core.py:264-277
        Gamma_sun = nu / Qf / 2                                                  # core.py:281
        Q_s = Qf * Gamma_scaled / Gamma_sun                                      # core.py:282
        psd_peak = sho_psd(w0_sun, S0f, w0_sun, Qf)                              # core.py:286-288
        A = 2 * mp.sqrt(4 * mp.pi * nu * psd_peak)                               # core.py:291 (Chaplin 2008)
        height = 2 * A ** 2 / (mp.pi * Gamma_sun) * fac                          # core.py:292-293
        A_s = mp.sqrt(mp.pi * Gamma_scaled * height / 2)                         # core.py:294
        psd_s = (A_s / 2) ** 2 / (4 * mp.pi * nu_s)                              # core.py:295-297
        S0_s = (mp.mpf("0.5") * mp.sqrt(mp.pi / 2) * psd_s / Q_s ** 2) * bg      # core.py:299-310
        if S0_s > 0 and w0_s > 0:                                                # core.py:317-318
            out.append((S0_s, w0_s, Q_s))
    return out


STARS = {
    "Sun": (1.0, 1.0, 5777.0, 1.0),
    "KIC 9333184": (0.9, 10.0, 4919.0, 52.3),        # docs/gadfly/synth.rst:44-50
    "subgiant": (1.25, 2.1, 6050.0, 5.3),
    "hot dwarf": (1.4, 1.5, 6600.0, 3.9),
    "red giant low numax": (1.1, 20.0, 4400.0, 135.0),
}


def main():
    doc = {"_about": "independent mpmath (40 digits) evaluation of Hyperparameters.for_star, flat bandpass; "
                     "tools/make_forstar_fixture.py; reference gadfly/core.py:107-333", "stars": {}}
    for name, (M, R, T, L) in STARS.items():
        terms = for_star(M, R, T, L)
        doc["stars"][name] = {"mass": M, "radius": R, "temperature": T, "luminosity": L,
                              "n_terms": len(terms),
                              "S0": [float(t[0]) for t in terms],
                              "w0": [float(t[1]) for t in terms],
                              "Q": [float(t[2]) for t in terms]}
        print(name, len(terms), "terms")
    with open(OUT, "w") as fh:
        json.dump(doc, fh)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
