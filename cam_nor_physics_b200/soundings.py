"""Synthetic physics_state chunks for the ZM deep-convection path.

The reference ships no input data, so the benchmark and the parity
tests run on synthetic soundings built the way CAM's FV dycore fills `physics_state`
(fv/dp_coupling.F90:554-557 for pmid/pdel, physics/geopotential.F90:218-247 FV branch for
zm/zi).  Everything is a pure function of (seed, global column index, field, level) through a
counter-based splitmix64 generator, so any shard of the global column set can be generated
independently and bit-identically (that is what makes N-GPU vs 1-GPU parity testable).

Array layout is the reference's chunk layout: every field is `[nchunks, nlev, pcols]`
(C order) == Fortran `(pcols, nlev)` chunk arrays back to back, column fastest.
"""
from __future__ import annotations

import dataclasses
import numpy as np

SEED = 20261018

# physconst (shr_const_mod) -- only used to build hydrostatically consistent inputs
_BOLTZ, _AVOGAD = 1.38065e-23, 6.02214e26
_RGAS = _AVOGAD * _BOLTZ
RAIR = _RGAS / 28.966
RH2O = _RGAS / 18.016
ZVIR = RH2O / RAIR - 1.0
EPSILO = 18.016 / 28.966
GRAVIT = 9.80616
PTOP = 226.0

# field ids for the counter-based RNG
_F = {n: i for i, n in enumerate(
    ["ps", "ts", "rhs", "conv", "phis", "pblh", "tpert", "land", "landfrac", "u0", "ushear",
     "v0", "vshear", "unoise", "vnoise", "cld", "tracer_amp", "tracer_exp", "fracis", "tnoise",
     "qnoise", "ts2", "rhs2", "lapse"])}


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _uniform(seed: int, col: np.ndarray, field: int, k=0, sub: int = 0) -> np.ndarray:
    """U[0,1) from (seed, column, field, level, sub-stream)."""
    with np.errstate(over="ignore"):
        col = np.asarray(col, dtype=np.uint64)
        k = np.asarray(k, dtype=np.uint64)
        key = (np.uint64(seed)
               ^ (col << np.uint64(24))
               ^ (np.uint64(field) << np.uint64(12))
               ^ (np.uint64(sub) << np.uint64(20))
               ^ k)
        r = _splitmix64(_splitmix64(key))
    return (r >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _normal(seed, col, field, k=0) -> np.ndarray:
    u1 = _uniform(seed, col, field, k, 0)
    u2 = _uniform(seed, col, field, k, 1)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def gg_svp_water(t):
    """Goff-Gratch saturation vapour pressure over water [Pa] (input construction only)."""
    tb = 373.16
    return 10.0 ** (-7.90298 * (tb / t - 1.0) + 5.02808 * np.log10(tb / t)
                    - 1.3816e-7 * (10.0 ** (11.344 * (1.0 - t / tb)) - 1.0)
                    + 8.1328e-3 * (10.0 ** (-3.49149 * (tb / t - 1.0)) - 1.0)
                    + np.log10(1013.246)) * 100.0


def sigma_interfaces(nlev: int) -> np.ndarray:
    x = np.arange(nlev + 1, dtype=np.float64) / nlev
    return 0.5 * x + 0.25 * (1.0 - np.cos(np.pi * x))


def reference_interface_pressures(nlev: int, ps0: float = 1.0e5) -> np.ndarray:
    return PTOP + (ps0 - PTOP) * sigma_interfaces(nlev)


def limcnv_for(nlev: int) -> int:
    """zm_conv_init's rule (zm_conv_intr.F90:357-368): first interface whose reference
    pressure is >= 40 hPa bounds convection; returns the 1-based interface index."""
    pref = reference_interface_pressures(nlev)
    if pref[0] >= 4.0e3:
        return 1
    for k in range(nlev):
        if pref[k] < 4.0e3 <= pref[k + 1]:
            return k + 1
    return nlev + 1


@dataclasses.dataclass
class Chunks:
    """One rank's chunk set.  All 2-D fields are [nchunks, nlev(+1), pcols] float64."""
    pcols: int
    pver: int
    nchunks: int
    ncol: np.ndarray          # [nchunks] int32
    col0: int                 # global index of the first column
    t: np.ndarray
    q: np.ndarray
    pmid: np.ndarray
    pint: np.ndarray
    pdel: np.ndarray
    zm: np.ndarray
    zi: np.ndarray
    phis: np.ndarray          # [nchunks, pcols]
    pblh: np.ndarray
    tpert: np.ndarray
    landfrac: np.ndarray
    u: np.ndarray
    v: np.ndarray
    cld: np.ndarray
    ztodt: float = 1800.0

    @property
    def ncols_total(self) -> int:
        return int(self.ncol.sum())


def make_chunks(ncols: int, pver: int = 32, pcols: int = 16, p_conv: float = 1.0,
                seed: int = SEED, col0: int = 0) -> Chunks:
    """Generate `ncols` columns starting at global column `col0`, packed into chunks of `pcols`.

    p_conv: Bernoulli probability that a column is a moist tropical (convectively unstable)
    sounding; the rest are cool, dry, stable members."""
    nchunks = (ncols + pcols - 1) // pcols
    npad = nchunks * pcols
    gcol = col0 + np.arange(npad, dtype=np.int64)
    # padded columns replicate the last real column so every lane holds valid physics
    gcol = np.minimum(gcol, col0 + ncols - 1)
    ncol = np.full(nchunks, pcols, dtype=np.int32)
    if ncols % pcols:
        ncol[-1] = ncols % pcols

    ps = 1.0e5 + 1.5e3 * _normal(seed, gcol, _F["ps"])
    is_conv = _uniform(seed, gcol, _F["conv"]) < p_conv
    ts = np.where(is_conv, 300.0 + 1.5 * _normal(seed, gcol, _F["ts"]),
                  272.0 + 16.0 * _uniform(seed, gcol, _F["ts2"]))
    rhs = np.where(is_conv, 0.70 + 0.25 * _uniform(seed, gcol, _F["rhs"]),
                   0.2 + 0.3 * _uniform(seed, gcol, _F["rhs2"]))
    lapse = np.where(is_conv, 0.205 + 0.02 * _uniform(seed, gcol, _F["lapse"]), 0.17)

    sig = sigma_interfaces(pver)                                   # [pver+1]
    pint = PTOP + (ps[None, :] - PTOP) * sig[:, None]              # [pver+1, npad]
    pmid = 0.5 * (pint[:-1] + pint[1:])
    pdel = pint[1:] - pint[:-1]

    kk = np.arange(pver, dtype=np.int64)[:, None]
    pr = pmid / ps[None, :]
    t = np.maximum(ts[None, :] * pr ** lapse[None, :], 195.0)
    t = t + 0.3 * _normal(seed, gcol[None, :], _F["tnoise"], kk)
    rh = rhs[None, :] * pr ** 1.5 + 0.15
    rh = np.clip(rh * (1.0 + 0.05 * _normal(seed, gcol[None, :], _F["qnoise"], kk)), 0.01, 0.98)
    es = gg_svp_water(t)
    qsat = EPSILO * es / (pmid - (1.0 - EPSILO) * es)
    q = np.maximum(rh * qsat, 1.0e-12)

    # hydrostatic heights, geopotential.F90:218-247 (FV branch)
    piln = np.log(pint)
    zi = np.zeros_like(pint)
    zm = np.zeros_like(pmid)
    rog = RAIR / GRAVIT
    for k in range(pver - 1, -1, -1):
        hkl = piln[k + 1] - piln[k]
        hkk = 1.0 - pint[k] * hkl / pdel[k]
        tv = t[k] * (1.0 + ZVIR * q[k])
        zm[k] = zi[k + 1] + rog * tv * hkk
        zi[k] = zi[k + 1] + rog * tv * hkl

    phis = GRAVIT * 2.0e3 * _uniform(seed, gcol, _F["phis"]) ** 3   # mostly low terrain
    pblh = 300.0 + 1200.0 * _uniform(seed, gcol, _F["pblh"])
    tpert = _uniform(seed, gcol, _F["tpert"])
    land = _uniform(seed, gcol, _F["land"])
    landfrac = np.where(land < 0.6, 0.0, np.where(land < 0.9, 1.0, _uniform(seed, gcol, _F["landfrac"])))

    u = (5.0 * _normal(seed, gcol, _F["u0"])[None, :]
         + 20.0 * _normal(seed, gcol, _F["ushear"])[None, :] * (1.0 - pr)
         + 2.0 * _normal(seed, gcol[None, :], _F["unoise"], kk))
    v = (3.0 * _normal(seed, gcol, _F["v0"])[None, :]
         + 10.0 * _normal(seed, gcol, _F["vshear"])[None, :] * (1.0 - pr)
         + 2.0 * _normal(seed, gcol[None, :], _F["vnoise"], kk))
    cld = np.where(pmid > 2.0e4, 0.8 * _uniform(seed, gcol[None, :], _F["cld"], kk), 0.0)

    def c2(a):   # [nlev, npad] -> [nchunks, nlev, pcols]
        return np.ascontiguousarray(a.reshape(a.shape[0], nchunks, pcols).transpose(1, 0, 2))

    def c1(a):
        return np.ascontiguousarray(a.reshape(nchunks, pcols))

    return Chunks(pcols=pcols, pver=pver, nchunks=nchunks, ncol=ncol, col0=col0,
                  t=c2(t), q=c2(q), pmid=c2(pmid), pint=c2(pint), pdel=c2(pdel), zm=c2(zm),
                  zi=c2(zi), phis=c1(phis), pblh=c1(pblh), tpert=c1(tpert), landfrac=c1(landfrac),
                  u=c2(u), v=c2(v), cld=c2(cld))


def make_tracers(ch: Chunks, ncnst: int, seed: int = SEED):
    """Tracer mixing ratios q[nchunks, ncnst, pver, pcols] (constituent 1 = water vapour),
    insoluble fractions fracis (same shape) and dry-layer thickness pdeldry."""
    nchunks, pver, pcols = ch.nchunks, ch.pver, ch.pcols
    gcol = ch.col0 + np.arange(nchunks * pcols, dtype=np.int64)
    gcol = np.minimum(gcol, ch.col0 + ch.ncols_total - 1).reshape(nchunks, 1, pcols)
    kk = np.arange(pver, dtype=np.int64).reshape(1, pver, 1)
    ps = ch.pint[:, -1:, :]
    pr = ch.pmid / ps
    q = np.empty((nchunks, ncnst, pver, pcols))
    fracis = np.empty_like(q)
    q[:, 0] = ch.q
    fracis[:, 0] = 1.0
    for m in range(1, ncnst):
        amp = np.exp(_normal(seed + 7919 * m, gcol, _F["tracer_amp"]))
        expo = -1.0 + 4.0 * _uniform(seed + 7919 * m, np.zeros(1, np.int64), _F["tracer_exp"])
        noise = np.exp(0.3 * _normal(seed + 7919 * m, gcol, _F["tracer_amp"], kk + 1))
        q[:, m] = 1.0e-9 * amp * pr ** expo * noise
        fracis[:, m] = 0.2 + 0.8 * _uniform(seed + 7919 * m, gcol, _F["fracis"], kk)
    pdeldry = ch.pdel * (1.0 - ch.q)
    return q, fracis, pdeldry
