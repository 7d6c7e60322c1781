"""numpy restatement of the reference's ephemeris evaluation and force model.

TEST INFRASTRUCTURE ONLY -- the product never imports this module.

An independent, vectorised restatement (over particles) of what the reference computes per
particle, written from the formulas, not from the loops.  It is pinned against the reference
build's outputs (tests/golden/*.npz, oracle/_ref) in tests/test_cpu_oracle.py at ~1e-15
relative; it is *not* bit-faithful (different operation order) and is not meant to be: the
bit-level oracle is the reference's own C code.  Its job is to be a second, structurally
different statement of the physics, and to give the variational terms an independent check
through numerical differentiation of the accelerations.

  ephemeris   reference src/spk.c:405-547, src/ascii_ephem.c:27-65, 275-384, src/forces.c:175-263
  forces      reference src/forces.c:266-344 (direct), 435-558 (Earth J2-J4), 644-725 (solar J2),
              774-904 (Marsden), 1059-1107 (potential GR), 1163-1218 (simple GR), 1288-1501 (EIH)
"""
from __future__ import annotations

import struct

import numpy as np

AU_LITERAL = 149597870.7
NPL = 11
NAIF_BY_ASSIST = [10, 1, 2, 399, 301, 4, 5, 6, 7, 8, 9]
DIRECT_ORDER = [10, 4, 5, 1, 9, 8, 3, 2, 7, 6, 0]


def cheb_TS(z, P):
    """Chebyshev T_p(z) and dT_p/dz for p < P."""
    T = np.zeros(P); S = np.zeros(P)
    T[0] = 1.0
    if P > 1:
        T[1] = z; S[1] = 1.0
    for p in range(2, P):
        T[p] = 2.0 * z * T[p - 1] - T[p - 2]
        S[p] = 2.0 * z * S[p - 1] + 2.0 * T[p - 1] - S[p - 2]
    return T, S


class SpkFile:
    """Type-2 SPK/DAF reader (reference src/spk.c:267-402, comment constants :696-766)."""

    def __init__(self, path):
        self.raw = np.fromfile(path, dtype=np.uint8)
        self.words = self.raw[: (self.raw.size // 8) * 8].view("<f8")
        assert bytes(self.raw[:7]) == b"DAF/SPK"
        fward = struct.unpack_from("<i", self.raw, 76)[0]
        self.targets = {}       # code -> dict(beg, end, segs=[(one, two)])
        self.order = []
        rec = fward
        while rec > 0:
            base = (rec - 1) * 1024
            nxt, _, nsum = struct.unpack_from("<ddd", self.raw, base)
            for s in range(int(nsum)):
                beg, end, tar, cen, ref, typ, one, two = struct.unpack_from("<ddiiiiii", self.raw, base + 24 + 40 * s)
                jb, je = 2451545.0 + beg / 86400.0, 2451545.0 + end / 86400.0
                if tar not in self.targets:
                    self.targets[tar] = dict(beg=jb, end=je, res=je - jb, segs=[], cen=cen)
                    self.order.append(tar)
                self.targets[tar]["segs"].append((one, two))
                self.targets[tar]["end"] = je
            rec = int(nxt)
        # constants from the comment area
        text = b""
        for r in range(2, fward):
            chunk = bytes(self.raw[(r - 1) * 1024: r * 1024]).rstrip(b"\0\4")
            text += chunk
        self.constants = {}
        seen = False
        for line in text.replace(b"\0", b"\n").decode("ascii", "replace").split("\n"):
            if "Initial conditions and constants used for integration:" in line:
                seen = True
            if not seen:
                continue
            parts = line.replace("D", "e").replace("d", "e").split()
            if len(parts) >= 2:
                try:
                    self.constants[parts[0]] = float(parts[1])
                except ValueError:
                    pass

    def posvel(self, code, jd_ref, t):
        """Position [km] and velocity [km/s] of target `code` at jd_ref + t (TDB).  The offset from the
        record mid-point is formed as (jd_ref - mid) + t, like the reference (src/spk.c:514), so that
        the 2.4e6-day Julian date does not eat the precision of t."""
        jd = jd_ref + t
        tg = self.targets[code]
        n = min(int((jd - tg["beg"]) / tg["res"]), len(tg["segs"]) - 1)
        one, two = tg["segs"][n]
        init, intlen, rsize, nrec = self.words[two - 4: two]
        R = int(rsize); P = (R - 2) // 3
        b = min(int((jd - (2451545.0 + init / 86400.0)) / (intlen / 86400.0)), int(nrec) - 1)
        rec = self.words[one - 1 + b * R: one - 1 + (b + 1) * R]
        mid, radius = rec[0], rec[1]
        z = ((jd_ref - (2451545.0 + mid / 86400.0)) + t) / (radius / 86400.0)
        T, S = cheb_TS(z, P)
        c = rec[2:].reshape(3, P)
        return c @ T, (c @ S) / radius


class De440File:
    """DE binary reader (reference src/ascii_ephem.c:105-252)."""

    COLS = dict(MER=0, VEN=1, EMB=2, MAR=3, JUP=4, SAT=5, URA=6, NEP=7, PLU=8, LUN=9, SUN=10)

    def __init__(self, path):
        self.raw = np.fromfile(path, dtype=np.uint8)
        beg, end, inc, ncon, au, emrat = struct.unpack_from("<dddidd", self.raw, 0x0A5C)
        self.beg, self.end, self.inc, self.au, self.emrat = beg, end, inc, au, emrat
        tri = []
        pos = 0x0A5C + 44
        for _ in range(12):
            tri.append(struct.unpack_from("<iii", self.raw, pos)); pos += 12
        pos += 4
        tri.append(struct.unpack_from("<iii", self.raw, pos)); pos += 12
        pos += 6 * (ncon - 400)
        for _ in range(2):
            tri.append(struct.unpack_from("<iii", self.raw, pos)); pos += 12
        self.tri = tri
        ncm = [3] * 15; ncm[11] = 2; ncm[14] = 1
        self.rec_words = 2 + sum(t[1] * t[2] * m for t, m in zip(tri, ncm))
        self.words = self.raw[: (self.raw.size // 8) * 8].view("<f8")
        names = [bytes(self.raw[0xFC + 6 * i: 0xFC + 6 * i + 6]).decode("ascii") for i in range(400)]
        names += [bytes(self.raw[0x0B28 + 6 * i: 0x0B28 + 6 * i + 6]).decode("ascii") for i in range(ncon - 400)]
        vals = self.words[self.rec_words: self.rec_words + ncon]
        self.constants = {n.strip(): float(v) for n, v in zip(names, vals)}

    def column(self, col, jd_ref, t_rel):
        jd = jd_ref + t_rel
        off, ncf, niv = self.tri[col]
        blk = min(int((jd - self.beg) / self.inc), self.words.size // self.rec_words - 3)
        rec = self.words[(blk + 2) * self.rec_words: (blk + 3) * self.rec_words]
        t = ((jd_ref - self.beg - blk * self.inc) + t_rel) / self.inc * niv
        b = min(int(t), niv - 1)
        z = 2.0 * (t - b) - 1.0
        T, S = cheb_TS(z, ncf)
        c = rec[off - 1 + b * 3 * ncf: off - 1 + (b + 1) * 3 * ncf].reshape(3, ncf)
        return c @ T, (c @ S) * (2.0 * niv / self.inc / 86400.0)


class Ephemeris:
    """assist_all_ephem for all bodies: GM, barycentric position [AU] and velocity [AU/day]."""

    def __init__(self, planets_path, asteroids_path):
        self.ast = SpkFile(asteroids_path)
        with open(planets_path, "rb") as f:
            magic = f.read(8)
        if magic == b"DAF/SPK ":
            self.spk = SpkFile(planets_path); self.de = None
            self.const = self.spk.constants
            self.au = self.const["AU"]
        else:
            self.de = De440File(planets_path); self.spk = None
            self.const = self.de.constants
            self.au = self.de.au
        c = self.const
        emrat = c["EMRAT"]
        gmb = c["GMB"]
        self.gm = np.array([c["GMS"], c["GM1"], c["GM2"], gmb * (emrat / (1. + emrat)), gmb * (1. / (1. + emrat)),
                            c["GM4"], c["GM5"], c["GM6"], c["GM7"], c["GM8"], c["GM9"]] +
                           [c.get("MA%04d" % (code - 2000000), 0.0) for code in self.ast.order])
        self.emrat = emrat
        self.nbodies = NPL + len(self.ast.order)
        self.c_au_day = c["CLIGHT"] / self.au * 86400.0

    def states(self, t, jd_ref=2451545.0):
        jd = jd_ref + t
        pos = np.zeros((self.nbodies, 3)); vel = np.full((self.nbodies, 3), np.nan)
        if self.spk is not None:
            emb_p, emb_v = self.spk.posvel(3, jd_ref, t)
            for b, code in enumerate(NAIF_BY_ASSIST):
                p, v = self.spk.posvel(code, jd_ref, t)
                if code in (301, 399):
                    p = p + emb_p; v = v + emb_v
                pos[b] = p / self.au; vel[b] = v / (self.au / 86400.0)
        else:
            emb_p, emb_v = self.de.column(2, jd_ref, t)
            lun_p, lun_v = self.de.column(9, jd_ref, t)
            cols = [10, 0, 1, None, None, 3, 4, 5, 6, 7, 8]
            for b, col in enumerate(cols):
                if col is None:
                    f = -1.0 / (1.0 + self.emrat) if b == 3 else self.emrat / (1.0 + self.emrat)
                    p, v = emb_p + f * lun_p, emb_v + f * lun_v
                else:
                    p, v = self.de.column(col, jd_ref, t)
                pos[b] = p / self.au; vel[b] = v / (self.au / 86400.0)
        for m, code in enumerate(self.ast.order):
            p, _ = self.ast.posvel(code, jd_ref, t)
            pos[NPL + m] = p / AU_LITERAL + pos[0]
        return self.gm, pos, vel


def _rot(ra_deg, dec_deg):
    """Rows of the rotation to the body-equatorial frame (reference src/forces.c:504-511)."""
    a, d = np.deg2rad(ra_deg), np.deg2rad(dec_deg)
    ca, sa, cd, sd = np.cos(a), np.sin(a), np.cos(d), np.sin(d)
    return np.array([[-sa, ca, 0.0], [-ca * sd, -sa * sd, cd], [ca * cd, sa * cd, sd]])


def _zonal(gm, Jn, R, d, Rm, orders):
    """Zonal-harmonic acceleration in the inertial frame; Rm rotates into the body frame."""
    p = d @ Rm.T
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    r2 = np.sum(p * p, axis=1); r = np.sqrt(r2)
    c2 = z * z / r2
    ax = np.zeros_like(x); ay = np.zeros_like(x); az = np.zeros_like(x)
    if 2 in orders:
        f = 3.0 * Jn[2] * R ** 2 / (2.0 * r2 * r2 * r)
        ax += gm * f * (5 * c2 - 1) * x; ay += gm * f * (5 * c2 - 1) * y; az += gm * f * (5 * c2 - 3) * z
    if 3 in orders:
        f = 5.0 * Jn[3] * R ** 3 / (2.0 * r2 * r2 * r)
        ax += -gm * f / r2 * (3 - 7 * c2) * x * z; ay += -gm * f / r2 * (3 - 7 * c2) * y * z
        az += -gm * f * (6 * c2 - 7 * c2 * c2 - 0.6)
    if 4 in orders:
        f = 5.0 * Jn[4] * R ** 4 / (8.0 * r2 * r2 * r2 * r)
        g4 = 63 * c2 * c2 - 42 * c2 + 3
        ax += gm * f * g4 * x; ay += gm * f * g4 * y; az += gm * f * (g4 + 12 - 28 * c2) * z
    return np.stack([ax, ay, az], axis=1) @ Rm


def accelerations(eph: Ephemeris, t, x, v, params=None, forces=0x7F, gr_eih_sources=1,
                  ng=(1.0, 0.0, 2.0, 5.093, 1.0)):
    """Acceleration [n][3] of real particles at barycentric x, v [n][3] (no variational part)."""
    gm, bp, bv = eph.states(t)
    c = eph.const
    n = x.shape[0]
    a = np.zeros((n, 3))
    over_c2 = 1.0 / eph.c_au_day ** 2
    if forces & 0x08 and params is not None:                       # Marsden, src/forces.c:774-904
        alpha, nk, nm, nn, r0 = ng
        d = x - bp[0]; dv = v - bv[0]
        r = np.linalg.norm(d, axis=1)
        g = alpha * (r / r0) ** (-nm) * (1.0 + (r / r0) ** nn) ** (-nk)
        h = np.cross(d, dv); tvec = np.cross(h, d)
        term = (params[:, 0:1] * d / r[:, None] + params[:, 1:2] * tvec / np.linalg.norm(tvec, axis=1)[:, None] +
                params[:, 2:3] * h / np.linalg.norm(h, axis=1)[:, None]) * g[:, None]
        active = np.any(params != 0.0, axis=1)
        a += np.where(active[:, None], term, 0.0)
    if forces & 0x10:                                              # Earth J2-J4, pole reset to (0, 90) deg
        a += _zonal(gm[3], {2: c["J2E"], 3: c["J3E"], 4: c["J4E"]}, c["RE"] / eph.au, x - bp[3], _rot(0.0, 90.0), (2, 3, 4))
    if forces & 0x20:                                              # solar J2, pole (286.13, 63.87) deg
        a += _zonal(gm[0], {2: c["J2SUN"]}, c["ASUN"] / eph.au, x - bp[0], _rot(286.13, 63.87), (2,))
    if forces & 0x40:                                              # EIH, beta = gamma = 1, src/forces.c:1319-1501
        term0 = np.zeros(n)
        for k in range(NPL):
            term0 += gm[k] / np.linalg.norm(x - bp[k], axis=1)
        vi2 = np.sum(v * v, axis=1)
        t7 = np.zeros((n, 3)); t8 = np.zeros((n, 3))
        for j in range(gr_eih_sources):
            dij = x - bp[j]; rij = np.linalg.norm(dij, axis=1)
            pref = gm[j] / rij ** 3
            aj = np.zeros(3); term1 = 0.0
            for k in range(NPL):
                if k == j:
                    continue
                djk = bp[j] - bp[k]; rjk = np.linalg.norm(djk)
                term1 += gm[k] / rjk
                aj -= gm[k] * djk / rjk ** 3
            vj = bv[j]
            rdv = dij @ vj
            factor = (-4.0 * over_c2 * term0 - over_c2 * term1 + over_c2 * vi2 + 2.0 * over_c2 * (vj @ vj)
                      - 4.0 * over_c2 * (v @ vj) - 1.5 * over_c2 * (rdv / rij) ** 2 - 0.5 * over_c2 * (dij @ aj))
            a += -(pref * factor)[:, None] * dij
            f = np.sum(dij * (4.0 * v - 3.0 * vj), axis=1)
            t7 += (pref * f)[:, None] * (v - vj)
            t8 += (gm[j] / rij * 3.5)[:, None] * aj[None, :]
        a += (t7 + t8) * over_c2
    if forces & 0x100:                                             # Nobili & Roxburgh, src/forces.c:1103
        d = x - bp[0]; r2 = np.sum(d * d, axis=1)
        a += (-6.0 * gm[0] ** 2 / (eph.c_au_day ** 2 * r2 * r2))[:, None] * d
    if forces & 0x80:                                              # Damour & Deruelle, src/forces.c:1211-1218
        d = x - bp[0]; dv = v - bv[0]; r = np.linalg.norm(d, axis=1)
        A = 4.0 * gm[0] / r - np.sum(dv * dv, axis=1); B = 4.0 * np.sum(d * dv, axis=1)
        a += (gm[0] / (r ** 3 * eph.c_au_day ** 2))[:, None] * (A[:, None] * d + B[:, None] * dv)
    if forces & 0x07:                                              # direct terms, asteroids first
        seq = list(range(NPL, eph.nbodies)) + DIRECT_ORDER
        for i in seq:
            if i == 0 and not forces & 0x01: continue
            if 0 < i < NPL and not forces & 0x02: continue
            if i >= NPL and not forces & 0x04: continue
            d = x - bp[i]; r = np.linalg.norm(d, axis=1)
            a -= (gm[i] / r ** 3)[:, None] * d
    return a


def variational_by_differences(eph, t, x, v, dx, dv, params=None, dparams=None, eps=1e-6, **kw):
    """d(acceleration) for a first-order variation (dx, dv[, dA]) by central differences."""
    p1 = None if params is None else params + (0.0 if dparams is None else eps * dparams)
    p2 = None if params is None else params - (0.0 if dparams is None else eps * dparams)
    ap = accelerations(eph, t, x + eps * dx, v + eps * dv, p1, **kw)
    am = accelerations(eph, t, x - eps * dx, v - eps * dv, p2, **kw)
    return (ap - am) / (2.0 * eps)


def interpolate_simulation(x0, v0, a0, br, dt_last_done, h):
    """Dense output between two snapshots, reference src/assist.c:682-752, from the IAS15 series itself rather than
    from the reference's running products: with s = dt_last_done * h,

        x(h) = x0 + s v0 + s^2 (a0/2 + h b0/6 + h^2 b1/12 + h^3 b2/20 + h^4 b3/30 + h^5 b4/42 + h^6 b5/56 + h^7 b6/72)
        v(h) = v0 + s (a0 + h b0/2 + h^2 b1/3 + h^3 b2/4 + h^4 b3/5 + h^5 b4/6 + h^6 b5/7 + h^7 b6/8)

    x0, v0: state at the start of the step (the earlier snapshot), a0, br[7]: start acceleration and b coefficients of
    the step (the later snapshot).  Arrays of any common shape; not bit-faithful (different operation order)."""
    x0, v0, a0 = (np.asarray(q, dtype=np.float64) for q in (x0, v0, a0))
    br = np.asarray(br, dtype=np.float64)
    s = dt_last_done * h
    px = a0 / 2.0
    pv = a0.copy()
    hp = 1.0
    for k, (dx, dv) in enumerate(zip((6., 12., 20., 30., 42., 56., 72.), (2., 3., 4., 5., 6., 7., 8.))):
        hp = hp * h
        px = px + hp * br[k] / dx
        pv = pv + hp * br[k] / dv
    return x0 + s * v0 + s * s * px, v0 + s * pv
