"""ctypes binding of the CPU oracle (oracle/zm_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class ZmParams(C.Structure):
    """Mirrors zmo_params_t (oracle/zm_oracle.h) and zm_params_t (include/zmconv_b200.h)."""
    _fields_ = [
        ("pcols", C.c_int), ("pver", C.c_int), ("limcnv", C.c_int), ("num_cin", C.c_int),
        ("zm_org", C.c_int), ("microp", C.c_int), ("no_deep_pbl", C.c_int), ("lparcel_pbl", C.c_int),
        ("cam3", C.c_int), ("masterproc", C.c_int),
        ("c0_lnd", C.c_double), ("c0_ocn", C.c_double), ("ke", C.c_double), ("ke_lnd", C.c_double),
        ("momcu", C.c_double), ("momcd", C.c_double), ("tiedke_add", C.c_double),
        ("capelmt", C.c_double), ("dmpdz", C.c_double), ("tau", C.c_double),
        ("cpair", C.c_double), ("epsilo", C.c_double), ("gravit", C.c_double), ("latice", C.c_double),
        ("latvap", C.c_double), ("tmelt", C.c_double), ("rair", C.c_double), ("cpwv", C.c_double),
        ("cpliq", C.c_double), ("rh2o", C.c_double), ("cpvir", C.c_double), ("zvir", C.c_double),
    ]


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


class Oracle:
    """One loaded flavour of the oracle: math='pm' (portable zm_math.h) or 'libm' (glibc)."""

    def __init__(self, math: str = "pm"):
        path = os.path.join(ORACLE_DIR, f"libzm_oracle_{math}.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = C.CDLL(path)
        L = self.lib
        L.zmo_math_backend.restype = C.c_char_p
        L.zmo_entropy.restype = C.c_double
        L.zmo_entropy.argtypes = [C.c_double] * 3
        L.zmo_enthalpy.restype = C.c_double
        L.zmo_enthalpy.argtypes = [C.c_double] * 4
        L.zmo_ientropy.argtypes = [C.c_double] * 4 + [c_dp, c_dp]
        L.zmo_ienthalpy.argtypes = [C.c_double] * 5 + [c_dp, c_dp]
        L.zmo_qsat_hpa.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        L.zmo_qsat_table.argtypes = [C.c_double, C.c_double, c_dp, c_dp]
        self.params = None
        self.math = math

    # -- init -------------------------------------------------------------------------
    def default_params(self, pcols, pver, limcnv) -> ZmParams:
        p = ZmParams()
        self.lib.zmo_params_default(C.byref(p), pcols, pver, limcnv)
        return p

    def convi(self, params: ZmParams) -> int:
        self.params = params
        return self.lib.zmo_convi(C.byref(params))

    def backend(self) -> str:
        return self.lib.zmo_math_backend().decode()

    # -- scalars ----------------------------------------------------------------------
    def entropy(self, t, p, q):
        return self.lib.zmo_entropy(t, p, q)

    def enthalpy(self, t, p, q, z):
        return self.lib.zmo_enthalpy(t, p, q, z)

    def ientropy(self, s, p, qt, tfg):
        t, qs = C.c_double(), C.c_double()
        rc = self.lib.zmo_ientropy(s, p, qt, tfg, C.byref(t), C.byref(qs))
        return rc, t.value, qs.value

    def ienthalpy(self, s, p, z, qt, tfg):
        t, qs = C.c_double(), C.c_double()
        rc = self.lib.zmo_ienthalpy(s, p, z, qt, tfg, C.byref(t), C.byref(qs))
        return rc, t.value, qs.value

    def qsat_hpa(self, t, p):
        es, q = C.c_double(), C.c_double()
        self.lib.zmo_qsat_hpa(t, p, C.byref(es), C.byref(q))
        return es.value, q.value

    def qsat_table(self, t, p):
        es, q = C.c_double(), C.c_double()
        self.lib.zmo_qsat_table(t, p, C.byref(es), C.byref(q))
        return es.value, q.value

    def counters(self):
        out = (C.c_longlong * 10)()
        self.lib.zmo_counters_get(out)
        return list(out)

    def flops(self):
        """3 phases x 8 columns operation count of this thread (zmo_flops_get)."""
        c = (C.c_longlong * 24)()
        self.lib.zmo_flops_get(c)
        return np.array(list(c), dtype=np.int64).reshape(3, 8)

    def counters_reset(self):
        self.lib.zmo_counters_reset()

    # -- zm_convr over a batch of chunks -------------------------------------------------
    def org_fields(self, org, orgt, org2d):
        self.lib.zmo_org_fields(_dp(org), _dp(orgt), _dp(org2d))

    def convr_batch(self, ch, delt=None, nthreads=0, org=None):
        """ch: soundings.Chunks.  Returns dict of outputs, arrays [nchunks, nlev, pcols]."""
        P = self.params
        nch, L, pc = ch.nchunks, P.pver, P.pcols
        delt = 0.5 * ch.ztodt if delt is None else delt
        z2 = lambda n=L: np.zeros((nch, n, pc))
        z1 = lambda: np.zeros((nch, pc))
        o = dict(prec=z1(), jctop=z1(), jcbot=z1(), qtnd=z2(), heat=z2(), mcon=z2(L + 1), cme=z2(),
                 cape=z1(), eurt=z2(), dlf=z2(), pflx=z2(L + 1), zdu=z2(), rprd=z2(), mu=z2(), md=z2(),
                 du=z2(), eu=z2(), ed=z2(), dp=z2(), dsubcld=z1(),
                 jt=np.zeros((nch, pc), np.int32), maxg=np.zeros((nch, pc), np.int32),
                 ideep=np.zeros((nch, pc), np.int32), lengath=np.zeros(nch, np.int32),
                 ql=z2(), rliq=z1(), dif=z2(), dnlf=z2(), dnif=z2(), rice=z1())
        ncol = np.ascontiguousarray(ch.ncol, dtype=np.int32)
        if org is not None:
            org = _f(org)
            o["orgt"], o["org2d"] = np.full_like(org, 7.0), np.zeros_like(org)
            self.org_fields(org, o["orgt"], o["org2d"])
        rc = self.lib.zmo_convr_batch(
            C.c_int(nch), _ip(ncol), _dp(_f(ch.t)), _dp(_f(ch.q)), _dp(o["prec"]), _dp(o["jctop"]),
            _dp(o["jcbot"]), _dp(_f(ch.pblh)), _dp(_f(ch.zm)), _dp(_f(ch.phis)), _dp(_f(ch.zi)),
            _dp(o["qtnd"]), _dp(o["heat"]), _dp(_f(ch.pmid)), _dp(_f(ch.pint)), _dp(_f(ch.pdel)),
            C.c_double(delt), _dp(o["mcon"]), _dp(o["cme"]), _dp(o["cape"]), _dp(o["eurt"]),
            _dp(_f(ch.tpert)), _dp(o["dlf"]), _dp(o["pflx"]), _dp(o["zdu"]), _dp(o["rprd"]),
            _dp(o["mu"]), _dp(o["md"]), _dp(o["du"]), _dp(o["eu"]), _dp(o["ed"]), _dp(o["dp"]),
            _dp(o["dsubcld"]), _ip(o["jt"]), _ip(o["maxg"]), _ip(o["ideep"]), _ip(o["lengath"]),
            _dp(o["ql"]), _dp(o["rliq"]), _dp(_f(ch.landfrac)), _dp(o["dif"]), _dp(o["dnlf"]),
            _dp(o["dnif"]), _dp(o["rice"]), C.c_int(nthreads))
        o["rc"] = rc
        return o

    # -- single-chunk routines (arrays shaped [nlev, pcols] / [ncnst, nlev, pcols]) --------
    def conv_evap(self, ncol, t, pmid, pdel, q, landfrac, prdprec, cldfrc, deltat, prec):
        P = self.params
        L, pc = P.pver, P.pcols
        o = dict(tend_s=np.zeros((L, pc)), tend_s_snwprd=np.zeros((L, pc)),
                 tend_s_snwevmlt=np.zeros((L, pc)), tend_q=np.zeros((L, pc)),
                 prec=np.array(prec, dtype=np.float64).copy(), snow=np.zeros(pc),
                 ntprprd=np.zeros((L, pc)), ntsnprd=np.zeros((L, pc)),
                 flxprec=np.zeros((L + 1, pc)), flxsnow=np.zeros((L + 1, pc)))
        self.lib.zmo_conv_evap(C.c_int(ncol), C.c_int(1), _dp(_f(t)), _dp(_f(pmid)), _dp(_f(pdel)),
                               _dp(_f(q)), _dp(_f(landfrac)), _dp(o["tend_s"]), _dp(o["tend_s_snwprd"]),
                               _dp(o["tend_s_snwevmlt"]), _dp(o["tend_q"]), _dp(_f(prdprec)),
                               _dp(_f(cldfrc)), C.c_double(deltat), _dp(o["prec"]), _dp(o["snow"]),
                               _dp(o["ntprprd"]), _dp(o["ntsnprd"]), _dp(o["flxprec"]), _dp(o["flxsnow"]))
        return o

    def convtran(self, doconvtran, q, mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, lengath,
                 fracis, dpdry, dt, cnst_is_dry):
        P = self.params
        L, pc = P.pver, P.pcols
        ncnst = q.shape[0]
        dqdt = np.zeros((ncnst, L, pc))
        do = np.ascontiguousarray(doconvtran, dtype=np.int32)
        dry = np.ascontiguousarray(cnst_is_dry, dtype=np.int32)
        self.lib.zmo_convtran(C.c_int(1), _ip(do), _dp(_f(q)), C.c_int(ncnst), _dp(_f(mu)), _dp(_f(md)),
                              _dp(_f(du)), _dp(_f(eu)), _dp(_f(ed)), _dp(_f(dp)), _dp(_f(dsubcld)),
                              _ip(np.ascontiguousarray(jt, np.int32)), _ip(np.ascontiguousarray(mx, np.int32)),
                              _ip(np.ascontiguousarray(ideep, np.int32)), C.c_int(1), C.c_int(int(lengath)),
                              C.c_int(0), _dp(_f(fracis)), _dp(dqdt), _dp(_f(dpdry)), C.c_double(dt), _ip(dry))
        return dqdt

    def momtran(self, ncol, domomtran, q, mu, md, du, eu, ed, dp, dsubcld, jt, mx, ideep, lengath, dt):
        P = self.params
        L, pc = P.pver, P.pcols
        ncnst = q.shape[0]
        o = dict(dqdt=np.zeros((ncnst, L, pc)), pguall=np.zeros((ncnst, L, pc)),
                 pgdall=np.zeros((ncnst, L, pc)), icwu=np.zeros((ncnst, L, pc)),
                 icwd=np.zeros((ncnst, L, pc)), seten=np.zeros((L, pc)))
        do = np.ascontiguousarray(domomtran, dtype=np.int32)
        self.lib.zmo_momtran(C.c_int(1), C.c_int(ncol), _ip(do), _dp(_f(q)), C.c_int(ncnst), _dp(_f(mu)),
                             _dp(_f(md)), _dp(_f(du)), _dp(_f(eu)), _dp(_f(ed)), _dp(_f(dp)),
                             _dp(_f(dsubcld)), _ip(np.ascontiguousarray(jt, np.int32)),
                             _ip(np.ascontiguousarray(mx, np.int32)),
                             _ip(np.ascontiguousarray(ideep, np.int32)), C.c_int(1), C.c_int(int(lengath)),
                             C.c_int(0), _dp(o["dqdt"]), _dp(o["pguall"]), _dp(o["pgdall"]), _dp(o["icwu"]),
                             _dp(o["icwd"]), C.c_double(dt), _dp(o["seten"]))
        return o

    def conv_tend_2_batch(self, doconvtran, q, pdeldry, fracis, ztodt, cnst_is_dry, tend, ptend_q=None, nthreads=0):
        """zm_conv_tend_2 (zm_conv_intr.F90:955-1028) over all chunks with the pbuf fields of `tend`
        (a conv_tend_batch result)."""
        q = _f(q)
        nch, pcnst = q.shape[0], q.shape[1]
        dq = np.zeros_like(q) if ptend_q is None else _f(ptend_q).copy()
        i32 = lambda a: _ip(np.ascontiguousarray(a, np.int32))
        self.lib.zmo_conv_tend_2_batch(C.c_int(nch), i32(doconvtran), _dp(q), C.c_int(pcnst), _dp(_f(pdeldry)),
                                       _dp(_f(fracis)), _dp(dq), C.c_double(ztodt), i32(cnst_is_dry),
                                       *[_dp(_f(tend[k])) for k in ("mu", "md", "du", "eu", "ed", "dp", "dsubcld")],
                                       *[i32(tend[k]) for k in ("jt", "maxg", "ideep", "lengath")], C.c_int(nthreads))
        return dq

    def conv_tend_diag(self, ncol, ps, pmid, mu, md, jt, maxg, ideep, lengath):
        """freqzm, mu_out, md_out, pcont, pconb of zm_conv_tend (zm_conv_intr.F90:685-729), one chunk."""
        P = self.params
        L, pc = P.pver, P.pcols
        o = dict(freqzm=np.zeros(pc), mu_out=np.zeros((L, pc)), md_out=np.zeros((L, pc)), pcont=np.zeros(pc),
                 pconb=np.zeros(pc))
        self.lib.zmo_conv_tend_diag(C.c_int(int(ncol)), _dp(_f(ps)), _dp(_f(pmid)), _dp(_f(mu)), _dp(_f(md)),
                                    _ip(np.ascontiguousarray(jt, np.int32)), _ip(np.ascontiguousarray(maxg, np.int32)),
                                    _ip(np.ascontiguousarray(ideep, np.int32)), C.c_int(int(lengath)),
                                    _dp(o["freqzm"]), _dp(o["mu_out"]), _dp(o["md_out"]), _dp(o["pcont"]),
                                    _dp(o["pconb"]))
        return o

    def conv_tend_batch(self, ch, nthreads=0, org=None, convtran1=None):
        """zm_conv_tend sequence on soundings.Chunks; same output names as zm_conv.zm_conv_tend.
        convtran1 = dict(doconvtran, cnst_is_dry, q, fracis[, ptend_q]) adds zm_conv_intr.F90:865-880 -> out["ptend_qc"]."""
        P = self.params
        nch, L, pc = ch.nchunks, P.pver, P.pcols
        out = {}
        if convtran1 is not None:
            q3, f3 = _f(convtran1["q"]), _f(convtran1["fracis"])
            out["ptend_qc"] = np.zeros_like(q3) if convtran1.get("ptend_q") is None else _f(convtran1["ptend_q"]).copy()
            do = np.ascontiguousarray(convtran1["doconvtran"], np.int32)
            dry = np.ascontiguousarray(convtran1["cnst_is_dry"] if convtran1.get("cnst_is_dry") is not None
                                       else np.zeros(q3.shape[1]), np.int32)
            self._tran1_keep = (q3, f3, do, dry)
            self.lib.zmo_convtran1_fields(C.c_int(q3.shape[1]), _ip(do), _ip(dry), _dp(q3), _dp(f3), _dp(out["ptend_qc"]))
        if org is not None:
            org = _f(org)
            out["orgt"], out["org2d"] = np.full_like(org, 7.0), np.zeros_like(org)
            self.org_fields(org, out["orgt"], out["org2d"])
        for k in ["ptend_s", "ptend_q", "ptend_u", "ptend_v", "cme", "zdu", "ql", "rprd", "evapcdp", "dlf",
                  "mu", "md", "du", "eu", "ed", "dp"]:
            out[k] = np.zeros((nch, L, pc))
        for k in ["mcon", "pflx", "flxprec", "flxsnow"]:
            out[k] = np.zeros((nch, L + 1, pc))
        for k in ["rliq", "rice", "jctop", "jcbot", "prec", "snow", "dsubcld", "cape"]:
            out[k] = np.zeros((nch, pc))
        for k in ["jt", "maxg", "ideep"]:
            out[k] = np.zeros((nch, pc), np.int32)
        out["lengath"] = np.zeros(nch, np.int32)
        order = ["ptend_s", "ptend_q", "ptend_u", "ptend_v", "mcon", "cme", "pflx", "zdu", "rliq", "rice",
                 "jctop", "jcbot", "prec", "snow", "ql", "rprd", "evapcdp", "flxprec", "flxsnow", "dlf",
                 "mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt", "maxg", "ideep", "lengath", "cape"]
        ins = [ch.t, ch.q, ch.u, ch.v, ch.pmid, ch.pint, ch.pdel, ch.zm, ch.zi, ch.phis, ch.pblh, ch.tpert,
               ch.landfrac, ch.cld]
        ins = [_f(a) for a in ins]
        ncol = np.ascontiguousarray(ch.ncol, dtype=np.int32)
        args = [C.c_int(nch), _ip(ncol)] + [_dp(a) for a in ins] + [C.c_double(ch.ztodt)]
        for k in order:
            a = out[k]
            args.append(_ip(a) if a.dtype == np.int32 else _dp(a))
        args.append(C.c_int(nthreads))
        out["rc"] = self.lib.zmo_conv_tend_batch(*args)
        return out

    def geopotential_t(self, ncol, dycore_lr, piln, pint, pmid, pdel, rpdel, t, q, rair, gravit, zvir):
        P = self.params
        L, pc = P.pver, P.pcols
        zi, zm = np.zeros((L + 1, pc)), np.zeros((L, pc))
        self.lib.zmo_geopotential_t(C.c_int(ncol), C.c_int(int(dycore_lr)), _dp(_f(piln)), None, _dp(_f(pint)),
                                    _dp(_f(pmid)), _dp(_f(pdel)), _dp(_f(rpdel)), _dp(_f(t)), _dp(_f(q)),
                                    _dp(_f(rair)), C.c_double(gravit), _dp(_f(zvir)), _dp(zi), _dp(zm))
        return zi, zm

    def geopotential_t_gen(self, ncol, dycore_lr, piln, pint, pmid, pdel, rpdel, t, q3, rair, gravit, zvir, species_idx):
        """generalized-Tv branch (geopotential.F90:248-310); q3: (ncnst, pver, pcols), species_idx 1-based."""
        P = self.params
        L, pc = P.pver, P.pcols
        zi, zm = np.zeros((L + 1, pc)), np.zeros((L, pc))
        q3 = _f(q3)
        sp = np.ascontiguousarray(species_idx, dtype=np.int32)
        self.lib.zmo_geopotential_t_gen(C.c_int(ncol), C.c_int(int(dycore_lr)), C.c_int(q3.shape[0]), C.c_int(sp.shape[0]),
                                        _ip(sp), _dp(_f(piln)), None, _dp(_f(pint)), _dp(_f(pmid)), _dp(_f(pdel)),
                                        _dp(_f(rpdel)), _dp(_f(t)), _dp(q3), _dp(_f(rair)), C.c_double(gravit),
                                        _dp(_f(zvir)), _dp(zi), _dp(zm))
        return zi, zm

    def convect_diagnostics(self, ncol, cmfmc, qc, rliq, pmid, rprddp, cnt, cnb):
        P = self.params
        L, pc = P.pver, P.pcols
        o = dict(cmfmc=_f(cmfmc).copy(), qc=_f(qc).copy(), qc2=np.ones((L, pc)), rliq=_f(rliq).copy(),
                 rliq2=np.ones(pc), cnt=_f(cnt).copy(), cnb=_f(cnb).copy(), cmfmc2=np.ones((L + 1, pc)),
                 rprdsh=np.ones((L, pc)), rprdtot=np.zeros((L, pc)), pcnt=np.zeros(pc), pcnb=np.zeros(pc))
        self.lib.zmo_convect_diagnostics(C.c_int(ncol), _dp(o["cmfmc"]), _dp(o["qc"]), _dp(o["qc2"]), _dp(o["rliq"]),
                                         _dp(o["rliq2"]), _dp(_f(pmid)), _dp(_f(rprddp)), _dp(o["cnt"]),
                                         _dp(o["cnb"]), _dp(o["cmfmc2"]), _dp(o["rprdsh"]), _dp(o["rprdtot"]),
                                         _dp(o["pcnt"]), _dp(o["pcnb"]))
        return o
