"""CPU tests of the PRODUCT's host side: the C-ABI library loads, exports every declared
symbol, parses the ephemeris containers exactly like the reference does, keeps the ABI
struct layouts, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest

import cases
import refharness as rh
from conftest import ROOT, planets_path
from assist_b200 import batch as ab
from assist_b200.cstructs import Ephem, Extras, Particle, Simulation
from assist_b200.synth import ephem_writer, populations


def _declared_functions():
    names = set()
    for hdr in ("assist.h", "assist_gpu.h", "assist_ephem_files.h", "rebound.h"):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        txt = re.sub(r"//.*", "", txt)
        for m in re.finditer(r"\b((?:assist|reb)_[a-z0-9_]+)\s*\(", txt):
            names.add(m.group(1))
    return sorted(names)


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in _declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    # data symbols read by the reference's Python loader (assist/_libassist.py:46-67)
    assert ctypes.c_char_p.in_dll(lib, "assist_version_str").value == b"1.2.0"
    assert ctypes.c_char_p.in_dll(lib, "assist_build_str").value
    assert ctypes.c_char_p.in_dll(lib, "assist_githash_str").value
    n = ctypes.c_int.in_dll(lib, "assist_error_messages_N").value
    msgs = (ctypes.c_char_p * n).in_dll(lib, "assist_error_messages")
    assert msgs[5].startswith(b"The requested time is outside the coverage")


def test_reference_python_loader_accepts_the_library(lib):
    """The reference's own assist/_libassist.py must load our .so unchanged (north star).
    It is re-stated here line by line because /root/reference does not exist on the GPU box."""
    from assist_b200 import _lib
    cl = ctypes.cdll.LoadLibrary(_lib.library_path())
    for name in ("assist_version_str", "assist_build_str", "assist_githash_str"):
        assert ctypes.c_char_p.in_dll(cl, name).value.decode("ascii")
    e_N = ctypes.c_int.in_dll(cl, "assist_error_messages_N").value
    assert (ctypes.c_char_p * e_N).in_dll(cl, "assist_error_messages")[0].decode("ascii").startswith("No error")


def test_struct_layouts_are_the_reference_abi():
    assert ctypes.sizeof(Ephem) == 208 and ctypes.sizeof(Extras) == 112 and ctypes.sizeof(Particle) == 128
    assert Ephem.spk_target_index.offset == 48 and Ephem.spk_emb_index.offset == 92 and Ephem.AU.offset == 96
    assert Ephem.over_c_squared.offset == 200
    assert Extras.particle_params.offset == 48 and Extras.forces.offset == 60 and Extras.gr_eih_sources.offset == 64
    assert Extras.alpha.offset == 72 and Extras.r0.offset == 104


def test_format_detection(lib, paths, tmp_path):
    """reference unit_tests/spk_detection + format_detection: header samples of real files."""
    bsp = b"DAF/SPK " + (2).to_bytes(4, "little") + (6).to_bytes(4, "little") + b"NIO2SPK ".ljust(48)
    titles = ("JPL Planetary Ephemeris DE440/LE440".ljust(84) +
              "Start Epoch: JED=  2287184.5  1549-DEC-21 00:00:00".ljust(84) +
              "Final Epoch: JED=  2688976.5  2650-JAN-25 00:00:00".ljust(84))
    names = "DENUM LENUM TDATEFTDATEBJDEPOCCENTERCLIGHTBETA  GMS   GM1   GM2   "
    de = (titles + names).encode("ascii")[:301]
    ascii_src = b"KSIZE=  2036    NCOEFF=  1018\nGROUP   1010\n" * 8
    cases_ = [(bsp, 0), (de, 2), (ascii_src, 3), (b"", 3)]
    for i, (blob, want) in enumerate(cases_):
        p = tmp_path / ("f%d" % i)
        p.write_bytes(blob)
        fd = os.open(p, os.O_RDONLY)
        try:
            assert lib.assist_detect_ephemeris_file_format(fd) == want
        finally:
            os.close(fd)
    for key, want in (("planets_bsp", 0), ("de440", 2), ("asteroids_bsp", 0)):
        fd = os.open(paths[key], os.O_RDONLY)
        assert lib.assist_detect_ephemeris_file_format(fd) == want
        os.close(fd)


def test_loaders_reject_bad_files(lib, tmp_path):
    p = tmp_path / "junk.bsp"
    p.write_bytes(b"NOT A DAF FILE" * 100)
    lib.assist_spk_init.restype = ctypes.c_void_p
    lib.assist_spk_init.argtypes = [ctypes.c_char_p]
    assert lib.assist_spk_init(str(p).encode()) is None
    assert lib.assist_spk_init(b"/nonexistent/file.bsp") is None
    assert not lib.assist_ephem_create(str(p).encode(), None)
    e = Ephem()
    assert lib.assist_ephem_init(ctypes.byref(e), b"/nonexistent/de440.bsp", None) == 1   # ASSIST_ERROR_EPHEM_FILE


def test_host_parsers_match_the_reference(lib, ref, paths, fmt):
    """Constants, GMs, target tables and time bounds parsed by the product equal the reference's."""
    mine = lib.assist_ephem_create(planets_path(paths, fmt).encode(), paths["asteroids_bsp"].encode())
    theirs = ref.assist_ephem_create(planets_path(paths, fmt).encode(), paths["asteroids_bsp"].encode())
    assert mine and theirs
    a, b = mine.contents, theirs.contents
    for f in ("jd_ref", "planets_source", "spk_emb_index", "AU", "EMRAT", "J2E", "J3E", "J4E", "J2SUN", "RE", "CLIGHT",
              "ASUN", "Re_eq", "Rs_eq", "c_AU_per_day", "c_squared", "over_c_squared"):
        if fmt == "440" and f == "spk_emb_index":
            continue
        assert getattr(a, f) == getattr(b, f), f
    if fmt == "bsp":
        assert list(a.spk_target_index) == list(b.spk_target_index)
    tb1, te1, tb2, te2 = (ctypes.c_double() for _ in range(4))
    lib.assist_ephem_time_bounds(mine, tb1, te1)
    ref.assist_ephem_time_bounds(theirs, tb2, te2)
    assert (tb1.value, te1.value) == (tb2.value, te2.value) == (-10544.5, 13455.5)
    lib.assist_ephem_free(mine)
    ref.assist_ephem_free(theirs)


def test_attach_installs_the_reference_defaults(lib, paths):
    """assist_attach / assist_init defaults, reference src/assist.c:408-447."""
    eph = lib.assist_ephem_create(paths["planets_bsp"].encode(), paths["asteroids_bsp"].encode())
    r = lib.reb_simulation_create()
    ax = lib.assist_attach(r, eph)
    a, s = ax.contents, r.contents
    assert a.forces == 0x7F and a.gr_eih_sources == 1 and a.geocentric == 0
    assert (a.alpha, a.nk, a.nm, a.nn, a.r0) == (1.0, 0.0, 2.0, 5.093, 1.0)
    assert s.integrator == 0 and s.gravity == 0 and s.force_is_velocity_dependent == 1 and s.ri_ias15.adaptive_mode == 1
    assert not lib.assist_attach(None, eph)
    idx0 = None
    lib.reb_simulation_add(r, Particle(x=1.0))
    idx0 = lib.reb_simulation_add_variation_1st_order(r, 0)
    assert idx0 == 1 and s.N == 2 and s.N_var == 1 and s.var_config[0].testparticle == 0 and s.var_config[0].index == 1
    lib.assist_free(ax)
    lib.reb_simulation_free(r)
    lib.assist_ephem_free(eph)


def test_compute_calls_fail_loudly_without_a_gpu(lib, have_gpu, paths):
    if have_gpu:
        pytest.skip("a GPU is present")
    eph = ab.EphemHandle(paths["planets_bsp"], paths["asteroids_bsp"])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        eph.eval([0.0])
    with pytest.raises(RuntimeError, match="no CUDA device"):
        ab.Batch(eph, 4)
    # through the REBOUND-style API the error surfaces as an error status, never as a CPU result
    r = lib.reb_simulation_create()
    ax = lib.assist_attach(r, eph.ptr)
    r.contents.t = cases.T0
    lib.reb_simulation_add(r, Particle(x=2.0, vy=0.01))
    status = lib.reb_simulation_integrate(r, cases.T0 + 10.0)
    assert status == 1 and r.contents.t == cases.T0 and r.contents.particles[0].x == 2.0
    err = ctypes.c_int(0)
    lib.assist_get_particle_with_error(eph.ptr, 0, cases.T0, err)
    assert err.value == 6
    lib.assist_free(ax)
    lib.reb_simulation_free(r)


def test_synthetic_inputs_are_deterministic(tmp_path):
    p1 = ephem_writer.write_all(str(tmp_path / "a"))
    p2 = ephem_writer.write_all(str(tmp_path / "b"))
    for k in p1:
        assert open(p1[k], "rb").read() == open(p2[k], "rb").read()
    a = populations.neo_mba_mix(1000, seed=3)
    assert np.array_equal(a, populations.neo_mba_mix(1000, seed=3))
    assert not np.array_equal(a, populations.neo_mba_mix(1000, seed=4))
    c, prm = populations.comets(100)
    assert c.shape == (100, 6) and prm.shape == (100, 3) and np.isfinite(c).all()


def test_packed_spk_copy_is_a_pure_rearrangement(lib, paths):
    """The device copy of an SPK kernel (gpu_api.cu, upload_packed_spk) against an independent reader of the file:
    records 16-byte aligned, [_jul(MID), RADIUS, (x y z) per term, zero terms up to a multiple of four], nothing else
    changed."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_oracle
    lib.assist_spk_init.restype = ctypes.c_void_p
    lib.assist_spk_init.argtypes = [ctypes.c_char_p]
    lib.assist_spk_free.argtypes = [ctypes.c_void_p]
    lib.assist_gpu_spk_pack_host.restype = ctypes.c_int
    lib.assist_gpu_spk_pack_host.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.POINTER(ctypes.c_double)),
                                            ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    for key in ("planets_bsp", "asteroids_bsp"):
        f = np_oracle.SpkFile(paths[key])
        h = lib.assist_spk_init(paths[key].encode())
        assert h
        out = ctypes.POINTER(ctypes.c_double)()
        words = ctypes.c_size_t(0)
        nt = len(f.order)
        off = (ctypes.c_longlong * (4 * nt))()
        assert lib.assist_gpu_spk_pack_host(h, ctypes.byref(out), ctypes.byref(words), off, 4 * nt) == 0
        packed = np.ctypeslib.as_array(out, shape=(words.value,)).copy()
        libc.free(out)
        lib.assist_spk_free(h)
        expect_words = 0
        for m, code in enumerate(f.order):
            for s, (one, two) in enumerate(f.targets[code]["segs"]):
                init, intlen, rsize, nrec = f.words[two - 4: two]
                R, nrec = int(rsize), int(nrec)
                P = (R - 2) // 3
                Rp = 2 + 3 * ((P + 3) & ~3)
                o = off[4 * m + s]
                assert o % 2 == 0 and o == expect_words
                src = f.words[one - 1: one - 1 + nrec * R].reshape(nrec, R)
                dst = packed[o: o + nrec * Rp].reshape(nrec, Rp)
                assert np.array_equal(dst[:, 0], 2451545.0 + src[:, 0] / 86400.0)
                assert np.array_equal(dst[:, 1], src[:, 1])
                coef = src[:, 2:].reshape(nrec, 3, P)
                assert np.array_equal(dst[:, 2: 2 + 3 * P].reshape(nrec, P, 3), coef.transpose(0, 2, 1))
                assert (dst[:, 2 + 3 * P:] == 0).all()
                expect_words += nrec * Rp
        assert words.value == expect_words + 2 and (packed[expect_words:] == 0).all()


def test_reference_programs_compile_and_link_unmodified():
    """Every C program of the reference (unit_tests/*/problem.c, examples/*/problem.c) compiles UNMODIFIED against
    include/ and links with the product library, the two SimulationArchive programs included.
    Covers the header names spk.h / ascii_ephem.h / forces.h / tools.h and the evaluator
    entry points of reference src/spk.h:113-122, src/ascii_ephem.h:15-22."""
    import ref_programs
    if not os.path.isdir(ref_programs.REF):
        pytest.skip("needs the reference tree")
    res = ref_programs.build_all()
    assert len(res) >= 34
    failed = {k for k, v in res.items() if not v[0]}
    assert failed == ref_programs.NEED_ARCHIVE, {k: res[k][1] for k in failed}


def test_snapshot_file_round_trip_without_a_gpu(lib, paths, tmp_path):
    """The snapshot file (include/rebound.h: reb_simulation_save_to_file / reb_simulationarchive_create_from_file /
    reb_simulation_create_from_simulationarchive_with_messages) is host code: particles, times and the variational
    configuration come back as written; a file that is not a snapshot file is refused; outside the stored range
    assist_create_interpolated_simulation returns NULL as the reference does (src/assist.c:602-610)."""
    eph = lib.assist_ephem_create(paths["planets_bsp"].encode(), paths["asteroids_bsp"].encode())
    r = lib.reb_simulation_create()
    ax = lib.assist_attach(r, eph)
    fn = str(tmp_path / "snap.bin").encode()
    lib.reb_simulation_add(r, Particle(x=1.0, y=2.0, z=3.0, vx=0.1, vy=0.2, vz=0.3))
    assert lib.reb_simulation_add_variation_1st_order(r, 0) == 1
    for k in range(3):
        r.contents.t = 100.0 + k
        r.contents.dt = 0.5 + k
        r.contents.particles[0].x = 1.0 + k
        lib.reb_simulation_save_to_file(r, fn)
    lib.assist_free(ax)
    lib.reb_simulation_free(r)
    sa = lib.reb_simulationarchive_create_from_file(fn)
    assert sa and sa.contents.nblobs == 3 and [sa.contents.t[i] for i in range(3)] == [100.0, 101.0, 102.0]
    r2 = lib.reb_simulation_create()
    lib.reb_simulation_create_from_simulationarchive_with_messages(r2, sa, 1, None)
    c = r2.contents
    assert (c.t, c.dt, c.N, c.N_var, c.N_var_config) == (101.0, 1.5, 2, 1, 1)
    assert (c.particles[0].x, c.particles[0].vz) == (2.0, 0.3)
    assert c.var_config[0].index == 1 and c.var_config[0].testparticle == 0
    assert c.ri_ias15.x0[0] == 2.0 and c.ri_ias15.v0[2] == 0.3 and c.ri_ias15.br.p6[5] == 0.0
    # a restored simulation written out again carries its step data along
    c.ri_ias15.a0[1] = 0.25
    c.ri_ias15.br.p3[4] = -0.5
    fn2 = str(tmp_path / "snap2.bin").encode()
    lib.reb_simulation_save_to_file(r2, fn2)
    lib.reb_simulation_free(r2)
    sa2 = lib.reb_simulationarchive_create_from_file(fn2)
    r3 = lib.reb_simulation_create()
    lib.reb_simulation_create_from_simulationarchive_with_messages(r3, sa2, 0, None)
    assert r3.contents.ri_ias15.a0[1] == 0.25 and r3.contents.ri_ias15.br.p3[4] == -0.5 and r3.contents.t == 101.0
    lib.reb_simulation_free(r3)
    lib.reb_simulationarchive_free(sa2)
    assert not lib.assist_create_interpolated_simulation(sa, 100.5)      # inside the first interval: no accelerations yet
    assert not lib.assist_create_interpolated_simulation(sa, 102.0)
    lib.reb_simulationarchive_free(sa)
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not a snapshot file" * 4)
    assert not lib.reb_simulationarchive_create_from_file(str(bad).encode())
    assert not lib.reb_simulationarchive_create_from_file(b"/nonexistent/file.bin")
    lib.assist_ephem_free(eph)
