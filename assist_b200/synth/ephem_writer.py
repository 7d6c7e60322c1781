"""Synthetic ephemeris files in the REAL on-disk formats ASSIST reads.

The JPL files (de440.bsp, linux_p1550p2650.440, sb441-n16.bsp) are not available
offline, so tests and benchmarks use files written here.  The byte layouts are the
ones the reference parses:

  * DE binary ".440"  -- reference src/ascii_ephem.c:105-252 (header walk),
    :27-65 (coefficient indexing), dev_tools/ephem_slicer/main.c:86-155
  * SPK/DAF ".bsp" type-2 -- reference src/spk.c:214-402 (file/summary records),
    :49-110 and :696-766 (comment-area constants), :432-445 (segment trailer)

Bodies move on fixed Keplerian ellipses (J2000 mean elements); the Moon on a
geocentric ellipse; the Sun is placed so that the barycentre stays at the origin.
Each record holds a Chebyshev interpolant of that motion.  The VALUES are free
parameters -- every consumer (reference build, oracle, CUDA path) reads the same
files -- only the FORMAT has to be exact.
"""
from __future__ import annotations

import os
import struct
import numpy as np

AU_KM = 149597870.7
JD_J2000 = 2451545.0
JD_BEG = 2441000.5
JD_END = 2465000.5
REC_DAYS = 32.0

CONSTANTS = {
    "AU": AU_KM,
    "CLIGHT": 299792.458,
    "EMRAT": 81.3005682214972,
    "J2E": 1.08262539e-3,
    "J3E": -2.53241e-6,
    "J4E": -1.619898e-6,
    "J2SUN": 2.1961391516529825e-7,
    "RE": 6378.1366,
    "ASUN": 696000.0,
    "GMS": 2.9591220828411956e-04,
    "GM1": 4.9125001948893182e-11,
    "GM2": 7.2434523326441187e-10,
    "GMB": 8.9970113929473466e-10,
    "GM4": 9.5495488297258119e-11,
    "GM5": 2.8253458252257917e-07,
    "GM6": 8.4597059933762903e-08,
    "GM7": 1.2920265649682399e-08,
    "GM8": 1.5243573478851939e-08,
    "GM9": 2.1750964648933581e-12,
}

# name, a [AU], e, i, L, long.peri, long.node [deg]  (ecliptic J2000 mean elements)
PLANET_ELEMENTS = [
    ("MER", 0.38709927, 0.20563593, 7.00497902, 252.25032350, 77.45779628, 48.33076593, "GM1"),
    ("VEN", 0.72333566, 0.00677672, 3.39467605, 181.97909950, 131.60246718, 76.67984255, "GM2"),
    ("EMB", 1.00000261, 0.01671123, -0.00001531, 100.46457166, 102.93768193, 0.0, "GMB"),
    ("MAR", 1.52371034, 0.09339410, 1.84969142, -4.55343205, -23.94362959, 49.55953891, "GM4"),
    ("JUP", 5.20288700, 0.04838624, 1.30439695, 34.39644051, 14.72847983, 100.47390909, "GM5"),
    ("SAT", 9.53667594, 0.05386179, 2.48599187, 49.95424423, 92.59887831, 113.66242448, "GM6"),
    ("URA", 19.18916464, 0.04725744, 0.77263783, 313.23810451, 170.95427630, 74.01692503, "GM7"),
    ("NEP", 30.06992276, 0.00859048, 1.77004347, -55.12002969, 44.96476227, 131.78422574, "GM8"),
    ("PLU", 39.48211675, 0.24882730, 17.14001206, 238.92903833, 224.06891629, 110.30393684, "GM9"),
]

# The 16 perturbers of sb441-n16 in the order that defines ASSIST body ids 11..26
# (reference assist/ephem.py:20-35).  number, a, e, i, node, argperi, M0 [deg], GM
ASTEROIDS = [
    (107, 3.487, 0.066, 10.00, 172.6, 306.8, 160.0, 1.4e-15),   # Camilla
    (1,   2.767, 0.079, 10.59,  80.3,  73.6,  60.0, 1.3964518123081070e-13),  # Ceres
    (65,  3.428, 0.112,  3.56, 155.6, 102.4, 250.0, 1.5e-15),   # Cybele
    (511, 3.164, 0.188, 15.94, 107.6, 337.2,  20.0, 4.3e-15),   # Davida
    (15,  2.644, 0.187, 11.75, 292.9,  98.5, 110.0, 4.5e-15),   # Eunomia
    (31,  3.155, 0.217, 26.30,  31.0,  61.5, 300.0, 2.4e-15),   # Euphrosyne
    (52,  3.095, 0.111,  7.48, 128.6, 343.4, 200.0, 3.6e-15),   # Europa
    (10,  3.142, 0.112,  3.83, 283.2, 312.3, 140.0, 1.2542530761640810e-14),  # Hygiea
    (704, 3.056, 0.155, 17.31, 280.3,  94.8,  80.0, 5.2e-15),   # Interamnia
    (7,   2.386, 0.230,  5.52, 259.6, 145.3, 330.0, 2.0e-15),   # Iris
    (3,   2.669, 0.257, 12.99, 169.9, 248.1,  45.0, 4.2e-15),   # Juno
    (2,   2.773, 0.230, 34.83, 173.0, 310.2, 270.0, 3.0471146330043200e-14),  # Pallas
    (16,  2.923, 0.134,  3.10, 150.0, 229.2, 190.0, 3.4e-15),   # Psyche
    (87,  3.483, 0.094, 10.88,  73.0, 263.6, 120.0, 2.2e-15),   # Sylvia
    (88,  2.769, 0.162,  5.21, 276.5,  36.6, 225.0, 1.8e-15),   # Thisbe
    (4,   2.362, 0.089,  7.14, 103.8, 150.7, 350.0, 3.8548000225257904e-14),  # Vesta
]

OBLIQUITY = np.deg2rad(23.43928)

# .440 column table (DE440): name, ncf, niv, ncm
ASCII_COLUMNS = [
    ("MER", 14, 4, 3), ("VEN", 10, 2, 3), ("EMB", 13, 2, 3), ("MAR", 11, 1, 3),
    ("JUP", 8, 1, 3), ("SAT", 7, 1, 3), ("URA", 6, 1, 3), ("NEP", 6, 1, 3),
    ("PLU", 6, 1, 3), ("LUN", 13, 8, 3), ("SUN", 11, 2, 3), ("NUT", 10, 4, 2),
    ("LIB", 10, 4, 3), ("MAN", 0, 0, 3), ("TDB", 0, 0, 1),
]

# planets .bsp: NAIF code, model key, interval [d], P (coefficients per component)
SPK_PLANET_TARGETS = [
    (1, "MER", 8.0, 14), (2, "VEN", 16.0, 10), (3, "EMB", 16.0, 13), (4, "MAR", 32.0, 11),
    (5, "JUP", 32.0, 8), (6, "SAT", 32.0, 7), (7, "URA", 32.0, 6), (8, "NEP", 32.0, 6),
    (9, "PLU", 32.0, 6), (10, "SUN", 16.0, 11), (301, "MOON_EMB", 4.0, 13), (399, "EARTH_EMB", 4.0, 13),
]


# --------------------------------------------------------------------------
# analytic model (positions in km, equatorial J2000, time = JD TDB)
# --------------------------------------------------------------------------

def _kepler_E(M, e):
    M = np.mod(M + np.pi, 2 * np.pi) - np.pi
    E = M + e * np.sin(M)
    for _ in range(60):
        dE = (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
        E = E - dE
        if np.max(np.abs(dE)) < 1e-15:
            break
    return E


def _ellipse_xyz(a, e, inc, node, argp, M):
    """Position on a Kepler ellipse (ecliptic frame), angles in rad; returns (3, n)."""
    E = _kepler_E(M, e)
    xp = a * (np.cos(E) - e)
    yp = a * np.sqrt(1 - e * e) * np.sin(E)
    cw, sw = np.cos(argp), np.sin(argp)
    cO, sO = np.cos(node), np.sin(node)
    ci, si = np.cos(inc), np.sin(inc)
    x = (cO * cw - sO * sw * ci) * xp + (-cO * sw - sO * cw * ci) * yp
    y = (sO * cw + cO * sw * ci) * xp + (-sO * sw + cO * cw * ci) * yp
    z = (sw * si) * xp + (cw * si) * yp
    return np.stack([x, y, z])


def _ecl2equ(v):
    ce, se = np.cos(OBLIQUITY), np.sin(OBLIQUITY)
    return np.stack([v[0], ce * v[1] - se * v[2], se * v[1] + ce * v[2]])


class SolarSystemModel:
    """Closed-form positions (km) of every body the files describe."""

    def __init__(self, n_extra_asteroids=0):
        self.gms = CONSTANTS["GMS"]
        self.emrat = CONSTANTS["EMRAT"]
        self.planets = {}
        for name, a, e, i, L, lp, node, gmkey in PLANET_ELEMENTS:
            gm = CONSTANTS[gmkey]
            n = np.sqrt((self.gms + gm) / a ** 3)            # rad/day
            self.planets[name] = dict(a=a, e=e, i=np.deg2rad(i), node=np.deg2rad(node),
                                      argp=np.deg2rad(lp - node), M0=np.deg2rad(L - lp), n=n, gm=gm)
        gmtot = self.gms + sum(p["gm"] for p in self.planets.values())
        self.gmtot = gmtot
        # geocentric Moon
        self.moon = dict(a=384400.0 / AU_KM, e=0.0549, i=np.deg2rad(5.145), node=np.deg2rad(125.08),
                         argp=np.deg2rad(318.15), M0=np.deg2rad(135.27), n=2 * np.pi / 27.321661)
        self.asteroids = []
        for num, a, e, i, node, argp, M0, gm in ASTEROIDS:
            n = np.sqrt(self.gms / a ** 3)
            self.asteroids.append(dict(num=num, a=a, e=e, i=np.deg2rad(i), node=np.deg2rad(node),
                                       argp=np.deg2rad(argp), M0=np.deg2rad(M0), n=n, gm=gm))

        # further perturbers of an sb441-n373 style kernel: seeded main-belt ellipses, numbers 1001 ...
        rng = np.random.default_rng(20261737)
        for k in range(n_extra_asteroids):
            a = rng.uniform(2.2, 3.4)
            self.asteroids.append(dict(num=1001 + k, a=a, e=rng.uniform(0.02, 0.25), i=np.deg2rad(rng.uniform(0.5, 25.0)),
                                       node=rng.uniform(0, 2 * np.pi), argp=rng.uniform(0, 2 * np.pi), M0=rng.uniform(0, 2 * np.pi),
                                       n=np.sqrt(self.gms / a ** 3), gm=10.0 ** rng.uniform(-16.5, -14.0)))

    def _helio(self, el, jd):
        M = el["M0"] + el["n"] * (jd - JD_J2000)
        return _ecl2equ(_ellipse_xyz(el["a"], el["e"], el["i"], el["node"], el["argp"], M))

    def sun_bary(self, jd):
        s = np.zeros((3, np.size(jd)))
        for p in self.planets.values():
            s -= p["gm"] / self.gmtot * self._helio(p, jd)
        return s * AU_KM

    def moon_geo(self, jd):
        return self._helio(self.moon, jd) * AU_KM

    def position(self, key, jd):
        jd = np.atleast_1d(np.asarray(jd, dtype=np.float64))
        if key == "SUN":
            return self.sun_bary(jd)
        if key == "LUN":            # .440 column: geocentric Moon
            return self.moon_geo(jd)
        if key == "MOON_EMB":       # SPK 301 w.r.t. EMB
            return self.moon_geo(jd) * (self.emrat / (1.0 + self.emrat))
        if key == "EARTH_EMB":      # SPK 399 w.r.t. EMB
            return -self.moon_geo(jd) / (1.0 + self.emrat)
        if key in self.planets:
            return self.sun_bary(jd) + self._helio(self.planets[key], jd) * AU_KM
        if key.startswith("AST"):
            return self._helio(self.asteroids[int(key[3:])], jd) * AU_KM   # heliocentric
        if key in ("NUT", "LIB"):
            # smooth filler so those columns hold plausible, unused data
            ph = (jd - JD_J2000) / 6798.0 * 2 * np.pi
            return np.stack([1e-4 * np.sin(ph), 1e-4 * np.cos(ph), 1e-5 * np.sin(2 * ph)])
        raise KeyError(key)


def cheb_fit(fun, t0, t1, ncoef):
    """Chebyshev interpolant of fun on [t0,t1] (arrays of interval bounds).

    fun(jd)->(ncomp, n).  Returns coefficients (nint, ncomp, ncoef)."""
    t0 = np.asarray(t0, dtype=np.float64)
    t1 = np.asarray(t1, dtype=np.float64)
    nint = t0.size
    k = np.arange(ncoef)
    xk = np.cos(np.pi * (k + 0.5) / ncoef)                       # nodes in [-1,1]
    mid = 0.5 * (t0 + t1)
    rad = 0.5 * (t1 - t0)
    tt = (mid[:, None] + rad[:, None] * xk[None, :]).ravel()
    f = fun(tt)                                                  # (ncomp, nint*ncoef)
    ncomp = f.shape[0]
    f = f.reshape(ncomp, nint, ncoef)
    j = np.arange(ncoef)
    Tjk = np.cos(np.outer(j, np.pi * (k + 0.5) / ncoef))         # T_j(x_k)
    c = (2.0 / ncoef) * np.einsum("cik,jk->icj", f, Tjk)
    c[:, :, 0] *= 0.5
    return c


# --------------------------------------------------------------------------
# DE binary (.440) writer
# --------------------------------------------------------------------------

def write_de440(path, model=None, jd_beg=JD_BEG, jd_end=JD_END, extra_constants=4):
    """Write a DE-binary planets file (format of linux_p1550p2650.440)."""
    model = model or SolarSystemModel()
    nrec = int(round((jd_end - jd_beg) / REC_DAYS))
    assert jd_beg + nrec * REC_DAYS == jd_end
    offs = []
    off = 3                                             # 1-based, after the two JD words
    for name, ncf, niv, ncm in ASCII_COLUMNS:
        offs.append(off)
        off += ncf * niv * ncm
    nwords = off - 1
    rec = 8 * nwords
    assert nwords == 1018, nwords

    # constants table: the 20 the reader needs + MAxxxx + fillers; a few beyond #400
    names, values = [], []
    for k, v in CONSTANTS.items():
        names.append(k); values.append(v)
    for a in model.asteroids:
        names.append("MA%04d" % a["num"]); values.append(a["gm"])
    ncon = 400 + extra_constants
    filler = 0
    # move ASUN past index 400 to exercise the second name block
    tail_names, tail_values = [], []
    if extra_constants > 0:
        idx = names.index("ASUN")
        tail_names.append(names.pop(idx)); tail_values.append(values.pop(idx))
    while len(names) < 400:
        names.append("XX%04d" % filler); values.append(0.0); filler += 1
    while len(tail_names) < extra_constants:
        tail_names.append("YY%04d" % filler); tail_values.append(0.0); filler += 1
    names += tail_names; values += tail_values
    assert len(names) == ncon

    header = bytearray(rec)
    titles = ["SYNTHETIC JPL-FORMAT EPHEMERIS (assist-b200), KEPLERIAN ELLIPSES",
              "Start Epoch: JED= %.1f" % jd_beg, "Final Epoch: JED= %.1f" % jd_end]
    for i, t in enumerate(titles):
        header[84 * i:84 * (i + 1)] = t.ljust(84).encode("ascii")[:84]
    for i in range(400):
        header[0xFC + 6 * i:0xFC + 6 * (i + 1)] = names[i].ljust(6).encode("ascii")
    pos = 0x0A5C
    struct.pack_into("<dddi", header, pos, jd_beg, jd_end, REC_DAYS, ncon); pos += 28
    struct.pack_into("<dd", header, pos, CONSTANTS["AU"], CONSTANTS["EMRAT"]); pos += 16
    for p in range(12):
        struct.pack_into("<iii", header, pos, offs[p], ASCII_COLUMNS[p][1], ASCII_COLUMNS[p][2]); pos += 12
    struct.pack_into("<i", header, pos, 440); pos += 4
    struct.pack_into("<iii", header, pos, offs[12], ASCII_COLUMNS[12][1], ASCII_COLUMNS[12][2]); pos += 12
    assert pos == 0x0B28
    for i in range(400, ncon):
        header[pos:pos + 6] = names[i].ljust(6).encode("ascii"); pos += 6
    for p in (13, 14):
        struct.pack_into("<iii", header, pos, offs[p], ASCII_COLUMNS[p][1], ASCII_COLUMNS[p][2]); pos += 12

    constrec = bytearray(rec)
    struct.pack_into("<%dd" % ncon, constrec, 0, *values)

    data = np.zeros((nrec, nwords), dtype="<f8")
    rb = jd_beg + REC_DAYS * np.arange(nrec)
    data[:, 0] = rb
    data[:, 1] = rb + REC_DAYS
    for (name, ncf, niv, ncm), o in zip(ASCII_COLUMNS, offs):
        if ncf == 0:
            continue
        sub = REC_DAYS / niv
        t0 = (rb[:, None] + sub * np.arange(niv)[None, :]).ravel()
        c = cheb_fit(lambda jd: model.position(name, jd)[:ncm], t0, t0 + sub, ncf)
        # layout inside a record: [sub-interval][component][coefficient]
        data[:, o - 1:o - 1 + ncf * niv * ncm] = c.reshape(nrec, niv * ncm * ncf)
    with open(path, "wb") as f:
        f.write(header); f.write(constrec); f.write(data.tobytes())
    return path


# --------------------------------------------------------------------------
# SPK / DAF (.bsp) writer
# --------------------------------------------------------------------------

def _fortran_d(v):
    s = "%.18E" % v
    return s.replace("E", "D")


def _comment_records(lines):
    text = "\0".join(lines) + "\0"
    # a line separator must not be the last character of a 1000-char chunk,
    # otherwise the reader's end-of-record stripping glues two lines together
    out = []
    i = 0
    chunks = []
    while i < len(text):
        chunk = text[i:i + 1000]
        if len(chunk) == 1000 and chunk[-1] == "\0":
            text = text[:i + 999] + " " + text[i + 999:]
            chunk = text[i:i + 1000]
        chunks.append(chunk)
        i += 1000
    for n, chunk in enumerate(chunks):
        recb = bytearray(1024)
        b = chunk.encode("ascii")
        recb[:len(b)] = b
        if n == len(chunks) - 1:
            recb[len(b)] = 4                      # EOT ends the comment area
        out.append(bytes(recb))
    return out


def _write_daf(path, segments, comment_lines, ifname):
    """segments: list of (target, center, jd_beg, jd_end, interval_days, coeffs(nrec,3,P))."""
    crecs = _comment_records(comment_lines)
    fward = 2 + len(crecs)
    nsumrec = (len(segments) + 24) // 25
    # records: 1 file | comments | (summary, name) * nsumrec | data
    first_data_rec = fward + 2 * nsumrec
    word = (first_data_rec - 1) * 128 + 1            # 1-based double-word address
    blobs, summaries = [], []
    for tar, cen, jb, je, intlen, c in segments:
        nrec, _, P = c.shape
        rsize = 2 + 3 * P
        arr = np.zeros((nrec, rsize), dtype="<f8")
        init = (jb - JD_J2000) * 86400.0
        il = intlen * 86400.0
        arr[:, 0] = init + il * (np.arange(nrec) + 0.5)   # MID
        arr[:, 1] = il / 2.0                              # RADIUS
        arr[:, 2:] = c.reshape(nrec, 3 * P)
        trailer = np.array([init, il, float(rsize), float(nrec)], dtype="<f8")
        blob = arr.tobytes() + trailer.tobytes()
        nwords = len(blob) // 8
        one, two = word, word + nwords - 1
        summaries.append(struct.pack("<ddiiiiii", init, (je - JD_J2000) * 86400.0, tar, cen, 1, 2, one, two))
        blobs.append(blob)
        word = two + 1
    filerec = bytearray(1024)
    filerec[0:8] = b"DAF/SPK "
    struct.pack_into("<ii", filerec, 8, 2, 6)
    filerec[16:76] = ifname.ljust(60).encode("ascii")[:60]
    bward = fward + 2 * (nsumrec - 1)
    struct.pack_into("<iii", filerec, 76, fward, bward, word)
    filerec[88:96] = b"LTL-IEEE"
    with open(path, "wb") as f:
        f.write(filerec)
        for r in crecs:
            f.write(r)
        for s in range(nsumrec):
            chunk = summaries[25 * s:25 * (s + 1)]
            nxt = float(fward + 2 * (s + 1)) if s + 1 < nsumrec else 0.0
            prv = float(fward + 2 * (s - 1)) if s > 0 else 0.0
            srec = bytearray(1024)
            struct.pack_into("<ddd", srec, 0, nxt, prv, float(len(chunk)))
            for i, sm in enumerate(chunk):
                srec[24 + 40 * i:24 + 40 * (i + 1)] = sm
            f.write(srec)
            f.write(b" " * 1024)                     # name record
        data = b"".join(blobs)
        f.write(data)
        pad = (-len(data)) % 1024
        f.write(b"\0" * pad)
    return path


def _constant_comment_lines(model, with_asteroid_masses=True):
    lines = ["; synthetic SPK kernel written by assist_b200.synth.ephem_writer",
             "; bodies follow fixed Keplerian ellipses; format only is JPL's", "",
             "Initial conditions and constants used for integration:", ""]
    for k, v in CONSTANTS.items():
        lines.append("%-8s%s" % (k, _fortran_d(v)))
    if with_asteroid_masses:
        for a in model.asteroids:
            lines.append("%-8s%s" % ("MA%04d" % a["num"], _fortran_d(a["gm"])))
    return lines


def write_planets_bsp(path, model=None, jd_beg=JD_BEG, jd_end=JD_END, nseg=2):
    """Planet SPK: targets 10,1..9 w.r.t. SSB, 301/399 w.r.t. EMB; `nseg` equal segments each."""
    model = model or SolarSystemModel()
    span = (jd_end - jd_beg) / nseg
    segments = []
    for tar, key, intlen, P in SPK_PLANET_TARGETS:
        for s in range(nseg):
            jb = jd_beg + s * span
            nrec = int(round(span / intlen))
            assert jb + nrec * intlen == jb + span
            t0 = jb + intlen * np.arange(nrec)
            c = cheb_fit(lambda jd: model.position(key, jd), t0, t0 + intlen, P)
            cen = 3 if tar in (301, 399) else 0
            segments.append((tar, cen, jb, jb + span, intlen, c))
    return _write_daf(path, segments, _constant_comment_lines(model), "SYNTH-DE440")


def write_asteroids_bsp(path, model=None, jd_beg=JD_BEG, jd_end=JD_END, nseg=2, P=16, intlen=32.0):
    """Small-body SPK: 16 heliocentric targets 2000000+n in sb441-n16 order."""
    model = model or SolarSystemModel()
    span = (jd_end - jd_beg) / nseg
    segments = []
    for i, a in enumerate(model.asteroids):
        for s in range(nseg):
            jb = jd_beg + s * span
            nrec = int(round(span / intlen))
            t0 = jb + intlen * np.arange(nrec)
            c = cheb_fit(lambda jd: model.position("AST%d" % i, jd), t0, t0 + intlen, P)
            segments.append((2000000 + a["num"], 10, jb, jb + span, intlen, c))
    return _write_daf(path, segments, ["; synthetic sb441-n16 style kernel (assist-b200)"], "SYNTH-SB16")


def write_extended(outdir, n_asteroids=40, jd_beg=JD_BEG, jd_end=JD_END):
    """An sb441-n373 style pair: a small-body kernel with `n_asteroids` targets (the 16 of sb441-n16 first) and a
    planets kernel whose comment area carries the MAxxxx masses of all of them (same planet records as
    synth_planets.bsp).  Returns dict of paths.  Idempotent."""
    os.makedirs(outdir, exist_ok=True)
    paths = {"planets_bsp": os.path.join(outdir, "synth_planets_n%d.bsp" % n_asteroids),
             "asteroids_bsp": os.path.join(outdir, "synth_sb%d.bsp" % n_asteroids)}
    if all(os.path.exists(p) for p in paths.values()):
        return paths
    model = SolarSystemModel(n_extra_asteroids=n_asteroids - len(ASTEROIDS))
    write_planets_bsp(paths["planets_bsp"] + ".tmp", model, jd_beg, jd_end)
    write_asteroids_bsp(paths["asteroids_bsp"] + ".tmp", model, jd_beg, jd_end)
    for p in paths.values():
        os.replace(p + ".tmp", p)
    return paths


def write_all(outdir, jd_beg=JD_BEG, jd_end=JD_END):
    """Write the three files; returns dict of paths.  Idempotent (skips existing)."""
    os.makedirs(outdir, exist_ok=True)
    paths = {"de440": os.path.join(outdir, "synth_planets.440"),
             "planets_bsp": os.path.join(outdir, "synth_planets.bsp"),
             "asteroids_bsp": os.path.join(outdir, "synth_sb16.bsp")}
    if all(os.path.exists(p) for p in paths.values()):
        return paths
    model = SolarSystemModel()
    write_de440(paths["de440"] + ".tmp", model, jd_beg, jd_end)
    write_planets_bsp(paths["planets_bsp"] + ".tmp", model, jd_beg, jd_end)
    write_asteroids_bsp(paths["asteroids_bsp"] + ".tmp", model, jd_beg, jd_end)
    for p in paths.values():
        os.replace(p + ".tmp", p)
    return paths


if __name__ == "__main__":
    import sys
    print(write_all(sys.argv[1] if len(sys.argv) > 1 else "data"))
