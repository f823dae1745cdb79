"""Shared helpers for the parity tests."""
import numpy as np
from cam_nor_physics_b200 import soundings as S

INT_KEYS = ["lengath", "ideep", "jt", "maxg"]
GATHERED_2D = ["mu", "md", "du", "eu", "ed", "dp"]
GATHERED_1D = ["dsubcld", "jt", "maxg"]
RTOL, ATOL = 1e-10, 1e-14          # BASELINE.json north_star tolerance for r8 outputs


def get_oracle(math, pcols, pver, **overrides):
    from oracle_lib import Oracle
    o = Oracle(math)
    p = o.default_params(pcols, pver, S.limcnv_for(pver))
    for k, v in overrides.items():
        setattr(p, k, v)
    rc = o.convi(p)
    return o, p, rc


def init_cuda(pcols, pver, **overrides):
    from cam_nor_physics_b200 import zm_conv as Z
    p = Z.default_params(pcols, pver, S.limcnv_for(pver))
    for k, v in overrides.items():
        setattr(p, k, v)
    Z.zm_init(p)
    return Z


def state_of(ch):
    return dict(t=ch.t, q=ch.q, u=ch.u, v=ch.v, pmid=ch.pmid, pint=ch.pint, pdel=ch.pdel, zm=ch.zm, zi=ch.zi,
                phis=ch.phis, pblh=ch.pblh, tpert=ch.tpert, landfrac=ch.landfrac, cld=ch.cld)


def gather_mask(lengath, pcols):
    return np.arange(pcols)[None, :] < np.asarray(lengath)[:, None]


def masked(out, key, lengath, pcols):
    """Gathered outputs are only defined for rows < lengath (reference leaves the rest undefined)."""
    m = gather_mask(lengath, pcols)
    a = out[key]
    return a * (m[:, None, :] if a.ndim == 3 else m)


def assert_same(out, ref, keys, pcols, exact=True, what="", scales=None, skip_cols=None):
    """exact: bit for bit.  Otherwise |a-b| <= ATOL + RTOL*|b| (the north-star tolerance); for a field that the
    zm_conv_tend glue forms as a SUM of zm_conv outputs (ptend_s = heat + evaporation + KE dissipation,
    zm_conv_intr.F90:736/803/833) `scales[key]` holds the sum of the terms' magnitudes and replaces |b|: the
    terms keep the tolerance, their cancelling sum cannot.  skip_cols: (chunk, column) pairs left out of the
    comparison (near-threshold columns, reported by the caller)."""
    bad = []
    lengath = ref["lengath"]
    for k in keys:
        a, b = out[k], ref[k]
        if k in GATHERED_2D or k in GATHERED_1D:
            a, b = masked(out, k, lengath, pcols), masked(ref, k, lengath, pcols)
        if skip_cols is not None and len(skip_cols) and a.ndim >= 2 and k not in GATHERED_2D and k not in GATHERED_1D \
                and k != "ideep":
            a, b = a.copy(), b.copy()
            for c, i in skip_cols:
                a[c, ..., i] = 0
                b[c, ..., i] = 0
        if exact or a.dtype.kind == "i":
            if not np.array_equal(a, b):
                d = np.abs(a.astype(float) - b.astype(float))
                bad.append(f"{k}: {int((a != b).sum())} elements differ (max abs {d.max():.3e})")
        else:
            mag = np.abs(b)
            if scales is not None and k in scales:
                mag = np.maximum(mag, scales[k])
            d = np.abs(a - b)
            over = d > ATOL + RTOL * mag
            if over.any() or not np.all(np.isfinite(a) == np.isfinite(b)):
                bad.append(f"{k}: {int(over.sum())} elements over tolerance, max abs {d.max():.3e}, "
                           f"max rel {(d / np.maximum(mag, 1e-300))[d > ATOL].max():.3e}")
    assert not bad, what + " mismatch:\n  " + "\n  ".join(bad)


def near_threshold_columns(cape, capelmt=70.0):
    """Columns whose CAPE lies within 1e-12 relative of the trigger threshold (BASELINE: reported, not failed)."""
    return np.argwhere(np.abs(cape / capelmt - 1.0) <= 1e-12)


CONVR_KEYS = ["lengath", "ideep", "cape", "prec", "jctop", "jcbot", "qtnd", "heat", "mcon", "cme", "eurt", "dlf",
              "pflx", "zdu", "rprd", "ql", "rliq", "rice", "mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt",
              "maxg", "dif", "dnlf", "dnif"]
TEND_KEYS = ["lengath", "ideep", "cape", "prec", "snow", "jctop", "jcbot", "ptend_s", "ptend_q", "ptend_u",
             "ptend_v", "mcon", "cme", "pflx", "zdu", "rliq", "rice", "ql", "rprd", "evapcdp", "flxprec",
             "flxsnow", "dlf", "mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt", "maxg"]


def cuda_convr(Z, ch):
    return Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel,
                      0.5 * ch.ztodt, ch.tpert, ch.landfrac)


def dpdry_gathered(ch, ref, pdeldry):
    """dpdry(i,:) = pdeldry(ideep(i),:)/100 (zm_conv_intr.F90:1014-1017)."""
    dpdry = np.zeros_like(ch.pdel)
    for c in range(ch.nchunks):
        n = int(ref["lengath"][c])
        idx = ref["ideep"][c][:n] - 1
        dpdry[c][:, :n] = pdeldry[c][:, idx] / 100.0
    return dpdry


# ---- fixtures produced by executing the reference's own source text (tests/golden/make_reference_fixtures.py) ----
REFTEXT_CASES = ["config1_L32", "mixed_ragged_L32", "parcel_pbl_L58", "num_cin3_L32", "no_deep_pbl_L32",
                 "not_master_L32", "zm_org_L32", "cam3_L32", "parcel_pbl_numcin5_L32", "pcols24_L26",
                 "single_column_L32", "strong_entrainment_L32", "stress_cold_L32", "stress_near_saturated_L32",
                 "stress_low_pbl_L32", "stress_high_pbl_L32", "cam3_partial_second_pass_L32", "cam3_idle_chunk_L32"]
REFTEXT_CONVR = ["prec", "jctop", "jcbot", "qtnd", "heat", "mcon", "cme", "cape", "eurt", "dlf", "pflx", "zdu", "rprd",
                 "mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt", "maxg", "ideep", "ql", "rliq", "dif", "dnlf",
                 "dnif", "rice"]


def reftext_overrides(g):
    """zm_params overrides that reproduce the namelist of a reference-text fixture."""
    ov = dict(num_cin=int(g["nl_num_cin"]), no_deep_pbl=int(bool(g["nl_no_deep_pbl"])),
              lparcel_pbl=int(bool(g["nl_lparcel_pbl"])), masterproc=int(bool(g["nl_masterproc"])),
              dmpdz=float(g["nl_dmpdz"]), zm_org=int(bool(g["nl_zm_org"])), cam3=int(bool(g["nl_cam3"])),
              tiedke_add=float(g["nl_tiedke_add"]), capelmt=float(g["nl_capelmt"]), tau=float(g["nl_tau"]),
              c0_lnd=float(g["nl_c0_lnd"]), c0_ocn=float(g["nl_c0_ocn"]), ke=float(g["nl_ke"]),
              ke_lnd=float(g["nl_ke_lnd"]), momcu=float(g["nl_momcu"]), momcd=float(g["nl_momcd"]))
    return ov


def reftext_gather(a, n):
    """Gathered outputs are defined for the first lengath entries only."""
    return a[..., :n]
