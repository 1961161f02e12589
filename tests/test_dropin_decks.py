"""The drop-in seam end to end: the reference's OWN deck binaries (unmodified sources, built by oracle/Makefile into
oracle/_ref/) run with LD_PRELOAD=libvpic_b200.so, so every advance_p / sort_p / load_interpolator_array /
clear|reduce|unload_accumulator_array call inside the reference host program lands on the GPU.

 * the five legacy known-answer decks (test/integrated/legacy: accel, cyclo, inbndj, interpe, outbndj) must still
   print "pass" — exact E interpolation, exact acceleration, gyration, charge conservation in and across cells;
 * sample/harris (C1 of BASELINE.json) as shipped: its `energies` history with the GPU path must match the history of
   the same binary run on the CPU reference path within fp32 tolerance.
"""
import os
import shutil
import subprocess
import tempfile
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
LIB = os.path.join(ROOT, "vpic_b200", "libvpic_b200.so")


def _need(binary):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    path = os.path.join(REF, binary)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs /root/reference at build time)")
    return path


# VPB_MODE_AUTO with every array of a page or more tracked by page protection (the production threshold is 32 MB):
# the decks' host code reads and writes particles, fields and accumulators between calls, the reference's field
# solver and divergence cleaning run on the host, dumps go through fwrite — all of it has to fault the data back.
AUTO_ENV = {"VPIC_B200_MODE": "auto", "VPIC_B200_LAZY_MIN": "4096", "VPIC_B200_LAZY_CHUNK": "16384"}
MODES = {"coherent": {"VPIC_B200_MODE": "coherent"}, "auto": AUTO_ENV, "auto_pinned": dict(AUTO_ENV, VPIC_B200_PIN="1")}


def _run(path, args, preload, cwd, timeout=900, extra_env=None):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = LIB
        env.update(extra_env or {})
    r = subprocess.run([path] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("mode", ["coherent", "auto"])
@pytest.mark.parametrize("deck", ["accel", "cyclo", "inbndj", "interpe", "outbndj"])
def test_reference_kat_deck_passes_on_gpu_path(deck, mode):
    path = _need(f"{deck}.scalar")
    with tempfile.TemporaryDirectory() as d:
        rc, out = _run(path, ["1", "1"], True, d, extra_env=MODES[mode])
    assert rc == 0, out[-2000:]
    assert "pass" in out and "FAIL" not in out and "fail" not in out.replace("fail 0", ""), out[-2000:]


@pytest.mark.parametrize("binary,gold,tpp,mode", [("3d_test", "energies_gold.3d_test", 1, "coherent"),
                                                  ("3d_test_threaded", "energies_gold.3d_test_threaded", 8, "coherent"),
                                                  ("3d_test_threaded", "energies_gold.3d_test_threaded", 8, "auto"),
                                                  ("3d_test_threaded", "energies_gold.3d_test_threaded", 8, "auto_pinned"),
                                                  ("weibel_driver", "energies_gold.weibel_driver", 1, "coherent")])
def test_reference_golden_energy_test_passes_on_gpu_path(binary, gold, tpp, mode):
    """test/unit/energy_comparison — the reference's own golden-vector tests of the whole step loop (3d_test: 16^3
    cells, 16 ppc, 2 species, 50 steps; its --tpp 8 variant, whose particle load differs; weibel_driver: 700 steps)
    compare themselves against energies_gold.* with the reference's tolerances.  Here the unmodified test binaries
    run with the hot path on the GPU."""
    path = _need(f"{binary}.scalar")
    goldp = os.path.join(REF, gold)
    for preload in (False, True):
        with tempfile.TemporaryDirectory() as d:
            shutil.copy(goldp, d)
            if not preload and mode != "coherent":
                continue                                      # the CPU leg is the same for every mode
            rc, out = _run(path, ["--tpp", str(tpp)], preload, d, extra_env=MODES[mode])
            assert rc == 0 and "All tests passed" in out, out[-2000:]
            if preload:
                mine = np.loadtxt(os.path.join(d, "energies"), comments="%")
                ref = np.loadtxt(goldp, comments="%")
                n = min(len(mine), len(ref))
                # much tighter than the reference's own 1 % / 3 % / 30 %: kinetic energies to 1e-5
                assert np.abs(mine[:n, -2:] - ref[:n, -2:]).max() / np.abs(ref[:n, -2:]).max() < 1e-5


_cpu_history = {}


@pytest.mark.parametrize("deck,mode", [("simple", "coherent"), ("reconnection_test", "coherent"),
                                       ("simple", "auto"), ("reconnection_test", "auto")])
def test_reference_to_completion_deck_runs_on_gpu_path(deck, mode):
    """test/integrated/to_completion ("does not die" in the reference's own CTest): whole vpic_simulation::advance loop,
    two species, sorts, divergence cleaning on the host, dumps — with the hot path on the GPU through LD_PRELOAD.
    reconnection_test also writes an energies history, which must match the CPU run of the same binary."""
    path = _need(f"{deck}.scalar")
    hist = {}
    if deck in _cpu_history:
        hist["cpu"] = _cpu_history[deck]
    for tag, preload in (("cpu", False), ("gpu", True)):
        if tag in hist:
            continue
        d = tempfile.mkdtemp(prefix=f"{deck}_{tag}_")
        try:
            rc, out = _run(path, ["--tpp", "1"], preload, d, extra_env=MODES[mode])
            assert rc == 0 and "normal exit" in out, out[-2000:]
            en = os.path.join(d, "rundata", "energies")
            if os.path.exists(en):
                rows = [ln.split() for ln in open(en) if ln.strip() and not ln.startswith("%")]
                hist[tag] = np.array([[float(x) for x in r] for r in rows if len(r) > 3])
        finally:
            shutil.rmtree(d, ignore_errors=True)
    if "cpu" in hist:
        _cpu_history[deck] = hist["cpu"]
        a, b = hist["cpu"], hist["gpu"]
        assert a.shape == b.shape and len(a) > 0
        tot_a, tot_b = a[:, 1:].sum(axis=1), b[:, 1:].sum(axis=1)
        np.testing.assert_allclose(tot_b, tot_a, rtol=1e-4)


@pytest.mark.parametrize("mode", ["coherent", "auto_pinned"])
def test_harris_energy_history_matches_reference(mode):
    path = _need("harris.scalar")
    hist = {}
    if "harris" in _cpu_history:
        hist["cpu"] = _cpu_history["harris"]
    for tag, preload in (("cpu", False), ("gpu", True)):
        if tag in hist:
            continue
        d = tempfile.mkdtemp(prefix=f"harris_{tag}_")
        try:
            rc, out = _run(path, ["--tpp", "1"], preload, d, timeout=1800, extra_env=MODES[mode])
            assert rc == 0, out[-3000:]
            rows = [ln.split() for ln in open(os.path.join(d, "energies")) if ln.strip() and not ln.startswith("%")]
            hist[tag] = np.array([[float(x) for x in r] for r in rows if len(r) > 3])
        finally:
            shutil.rmtree(d, ignore_errors=True)
    _cpu_history["harris"] = hist["cpu"]
    a, b = hist["cpu"], hist["gpu"]
    assert a.shape == b.shape and a.shape[0] >= 5, (a.shape, b.shape)
    assert np.array_equal(a[:, 0], b[:, 0])                        # same steps
    # columns: step ex ey ez bx by bz electron ion.  Tolerances: the reference's own golden test allows 1 % on
    # particle energy and 3 % on B energy (test/unit/energy_comparison/3d_test.cc:330-351); here both runs share
    # the scalar arithmetic, so only the fp32 deposit order differs — but 484 steps of a tearing-unstable current sheet
    # amplify that: 5e-4 relative on the dominant terms, 2e-4 of the total energy on every term (run-to-run spread
    # of the GPU path itself, whose atomic order is not deterministic, is ~1e-4).
    tot_a, tot_b = a[:, 1:].sum(axis=1), b[:, 1:].sum(axis=1)
    np.testing.assert_allclose(tot_b, tot_a, rtol=1e-5)
    total = np.abs(tot_a).max()
    for col in range(1, a.shape[1]):
        scale, diff = np.abs(a[:, col]).max(), np.abs(a[:, col] - b[:, col]).max()
        assert diff <= 2e-4 * total, (col, diff, total)            # every component, against the total energy
        if scale >= 1e-2 * total:                                  # dominant components (main B, kinetic energies)
            assert diff / scale < 5e-4, (col, diff, scale)
        elif scale > 0:                                            # noise-driven components: same order of magnitude
            assert diff / scale < 1e-1, (col, diff, scale)


def _trace(out):
    line = [ln for ln in out.splitlines() if ln.startswith("vpic_b200 trace")][-1]
    return {k: int(v) for k, v in (tok.split("=") for tok in line.split()[2:])}


@pytest.mark.parametrize("fields", ["device", "reference"])
def test_preloaded_deck_uses_the_device_kernels_it_claims(fields):
    """VPIC_B200_TRACE=1 reports how often each entry point ran on the device.  Under LD_PRELOAD the reference's own
    field-kernel symbols (advance_b, vacuum_advance_e, clear_jf, synchronize_jf, vacuum_energy_f) are interposed as
    well, so an unmodified deck advances its fields on the GPU; VPIC_B200_FIELDS=0 hands them back to the reference
    through dlsym(RTLD_NEXT) — and the deck's answers do not change either way."""
    path = _need("reconnection_test.scalar")
    env = {"VPIC_B200_TRACE": "1"}
    if fields == "reference":
        env["VPIC_B200_FIELDS"] = "0"
    d = tempfile.mkdtemp(prefix="trace_")
    try:
        rc, out = _run(path, ["--tpp", "1"], True, d, extra_env=env)
        assert rc == 0 and "normal exit" in out, out[-2000:]
        t = _trace(out)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    assert t["advance_p"] > 0 and t["sort_p"] > 0 and t["load_interpolator_array"] > 0 and t["unload_accumulator_array"] > 0
    if fields == "device":
        assert t["advance_b"] >= 2 * t["advance_e"] > 0 and t["clear_jf"] > 0 and t["synchronize_jf"] > 0
        assert t["divergence_cleaning_kernels"] > 0                  # the deck cleans div E and div B at intervals
        assert t["field_kernel_fallback_to_reference"] == 0
    else:
        assert t["advance_b"] == t["advance_e"] == t["clear_jf"] == t["divergence_cleaning_kernels"] == 0
        assert t["field_kernel_fallback_to_reference"] > 0


def _tail_floats(path, count):
    raw = np.fromfile(path, dtype=np.uint8)
    return raw[len(raw) - 4 * count:].view(np.float32)


@pytest.mark.parametrize("mode", ["coherent", "auto"])
def test_dump_deck_files_match_reference(mode):
    """test/integrated/to_completion/dump.deck: two steps that dump energies, fields, hydro moments of both species,
    particles and a checkpoint on every step — every way the reference's host code reads the arrays the device owns
    (fwrite of whole arrays, accumulate_hydro_p, center_p on a staging buffer, checkpt).  The files written with the
    hot path on the GPU must match the CPU run's: same names and sizes, payloads within fp32 deposit-order tolerance;
    then the reference's restart test (--restore checkpt_test.1) has to run to completion on the GPU path too."""
    path = _need("dump.scalar")
    nv = 10 * 10 * 3                                             # 8 x 8 x 1 cells plus ghosts
    runs = {}
    for tag, preload in (("cpu", False), ("gpu", True)):
        d = tempfile.mkdtemp(prefix=f"dump_{tag}_")
        rc, out = _run(path, ["--tpp", "1"], preload, d, extra_env=dict(MODES[mode], VPIC_B200_TRACE="1"))
        assert rc == 0 and "normal exit" in out, out[-2000:]
        runs[tag] = (d, out)
    try:
        cpu, gpu = runs["cpu"][0], runs["gpu"][0]
        names = sorted(os.listdir(cpu))
        assert names == sorted(os.listdir(gpu))
        for n in names:
            if n.startswith("checkpt"):
                continue          # a checkpoint names the field kernels by symbol and library, which differ by design
            assert os.path.getsize(os.path.join(cpu, n)) == os.path.getsize(os.path.join(gpu, n)), n
        t = _trace(runs["gpu"][1])
        assert t["hydro_kernels"] > 0 and t["advance_p"] > 0
        for step in (1, 2):
            for sp in ("e", "i"):
                a = _tail_floats(os.path.join(cpu, f"{sp}hydro.{step}.0"), nv * 16).reshape(nv, 16)[:, :14]
                b = _tail_floats(os.path.join(gpu, f"{sp}hydro.{step}.0"), nv * 16).reshape(nv, 16)[:, :14]
                assert np.abs(a).max() > 0
                for col in range(14):
                    scale = np.abs(a[:, col]).max()
                    assert np.abs(a[:, col] - b[:, col]).max() <= 2e-5 * scale + 1e-30, (sp, step, col)
            fa_ = _tail_floats(os.path.join(cpu, f"fields.{step}.0"), nv * 20).reshape(nv, 20)
            fb_ = _tail_floats(os.path.join(gpu, f"fields.{step}.0"), nv * 20).reshape(nv, 20)
            for lo, hi in ((0, 3), (4, 7), (12, 15)):            # e, cb, jf
                scale = np.abs(fa_[:, lo:hi]).max()
                assert np.abs(fa_[:, lo:hi] - fb_[:, lo:hi]).max() <= 5e-5 * scale + 1e-30, (step, lo)
            for sp in ("e", "i"):
                pa_ = np.fromfile(os.path.join(cpu, f"{sp}particle.{step}.0"), dtype=np.uint8)
                pb_ = np.fromfile(os.path.join(gpu, f"{sp}particle.{step}.0"), dtype=np.uint8)
                npart = 8 * 8 * 8                                # 0.5 * nppc * cells per species
                qa = pa_[len(pa_) - 32 * npart:].view(np.float32).reshape(npart, 8)
                qb = pb_[len(pb_) - 32 * npart:].view(np.float32).reshape(npart, 8)
                assert np.array_equal(qa[:, 3].view(np.int32), qb[:, 3].view(np.int32))       # same voxels, same order
                np.testing.assert_allclose(qb[:, [0, 1, 2, 4, 5, 6, 7]], qa[:, [0, 1, 2, 4, 5, 6, 7]], rtol=0, atol=2e-5)
        # the reference's restart test, on the GPU path, from the checkpoint the GPU run wrote
        rc, out = _run(path, ["--tpp", "1", "--restore", os.path.join(gpu, "checkpt_test.1")], True, gpu, extra_env=MODES[mode])
        assert rc == 0 and "normal exit" in out, out[-2000:]
    finally:
        for d, _ in runs.values():
            shutil.rmtree(d, ignore_errors=True)


def test_reference_grid_heating_rate_on_gpu_path():
    """test/unit/grid_heating with the hot path on the GPU: the numerical heating rate over 27 000 steps must stay
    within the 5 standard deviations the reference's own checker allows around its authors' mean.  (15 s on a B200,
    profiles/r02b_grid_heating_gpu.log.)"""
    _need("gridHeatingTestElec.scalar")
    import test_oracle_vs_ref as T
    out, verdict = T._grid_heating({"LD_PRELOAD": LIB, "VPIC_B200_TRACE": "1"}, timeout=3000)
    assert "Electron heating rate test PASS" in verdict, verdict[-1500:]
    t = _trace(out)
    assert t["advance_p"] > 20000 and t["field_kernel_fallback_to_reference"] == 0


def test_lpi_2d_deck_matches_reference_and_reports_its_forwards():
    """BASELINE.json configs[2]: sample/lpi_2d_F6_test (96 x 1 x 549 cells, electrons + He, absorbing field walls on
    four sides, maxwellian_reflux particle walls, a laser injected into fa->f by host code every step,
    num_comm_round = 6, divergence cleaning every 200 / 20 steps) through the drop-in seam, against the same binary on
    the CPU.  Both runs stop after 245 steps (the deck dumps fields and hydro every 244); particles per cell 128 by
    default, the shipped 512 with VPIC_B200_LONG_TESTS=1.

    Parity: the field and electron-hydro dumps of step 244 agree within fp32 deposit-order tolerance.
    Honesty: the custom particle boundary (host function pointers, host RNG) is NOT on the device — boundary_p is
    forwarded to the reference's CPU code; the test asserts the once-per-symbol warning, reads the forward count and
    the bytes that faulted back per step from the trace, and checks that VPIC_B200_STRICT=1 refuses to run."""
    path = _need("lpi_2d_F6_test.scalar")
    nppc = "512" if os.environ.get("VPIC_B200_LONG_TESTS") == "1" else "128"
    steps = 245
    base = {"VPIC_LPI_STEPS": str(steps), "VPIC_LPI_NPPC": nppc}
    runs = {}
    try:
        # "cpu4": the same CPU binary with another thread count — another summation order of the same deposits, the
        # yardstick for how far two legitimate runs of this deck drift apart in 245 steps
        for tag, preload, tpp in (("cpu", False, "8"), ("cpu4", False, "4"), ("gpu", True, "8")):
            d = tempfile.mkdtemp(prefix=f"lpi_{tag}_")
            env = dict(os.environ, **base)
            if preload:
                env.update({"LD_PRELOAD": LIB, "VPIC_B200_TRACE": "1"})
            r = subprocess.run([path, "--tpp", tpp], cwd=d, env=env, capture_output=True, text=True, timeout=3000)
            out = r.stdout + r.stderr
            assert r.returncode == 0 and "normal exit" in out, out[-3000:]
            runs[tag] = (d, out)
        cpu, gpu, cpu4 = runs["cpu"][0], runs["gpu"][0], runs["cpu4"][0]
        out = runs["gpu"][1]
        # the forward is loud, once, and names its reason
        warn = [ln for ln in out.splitlines() if "is not served on the device" in ln]
        assert len([w for w in warn if w.strip().startswith("boundary_p")]) == 1, warn
        assert "custom particle boundary handlers" in " ".join(warn)
        t = _trace(out)
        assert t["advance_p"] == 2 * steps and t["field_kernel_fallback_to_reference"] == 0
        assert t["advance_b"] == 2 * steps and t["advance_e"] == steps      # absorbing (Higdon) walls run on the device
        per_step = t["lazy_fault_bytes"] / steps
        print(f"lpi_2d_F6_test on the GPU path: {t['lazy_faults']} page-fault fetches, {per_step / 1e6:.1f} MB faulted "
              f"back per step (boundary_p forwarded to the host: maxwellian_reflux walks movers on the CPU)")
        for rel, cols in ((f"field/T.{steps - 1}/fields.{steps - 1}.0", 20), (f"ehydro/T.{steps - 1}/e_hydro.{steps - 1}.0", 16)):
            fa_, fb_ = os.path.join(cpu, rel), os.path.join(gpu, rel)
            assert os.path.getsize(fa_) == os.path.getsize(fb_), rel
        # the banded dumps end with the payload, variable-major: compare every band against its own scale
        for rel, nband in ((f"field/T.{steps - 1}/fields.{steps - 1}.0", 6), (f"ehydro/T.{steps - 1}/e_hydro.{steps - 1}.0", 4)):
            size = os.path.getsize(os.path.join(cpu, rel))
            a = np.fromfile(os.path.join(cpu, rel), dtype=np.uint8)
            b = np.fromfile(os.path.join(gpu, rel), dtype=np.uint8)
            assert a.size == b.size == size
            c = np.fromfile(os.path.join(cpu4, rel), dtype=np.uint8)
            nfl = (size // 4) * 3 // 4                            # the last three quarters of the file are all payload
            fa_, fb_, fc_ = (x[size - 4 * nfl:].view(np.float32).astype(np.float64) for x in (a, b, c))
            assert np.isfinite(fa_).all() and np.isfinite(fb_).all()
            # 245 steps of reordered fp32 deposits; one particle that reaches a wall a step earlier or later also shifts
            # the host RNG stream of the reflux walls, so later re-injections differ particle by particle.  The bound is
            # therefore statistical, band by band (the dumps are variable-major), and relative to the distance between
            # two CPU runs of the same binary that differ only in their thread count.
            worst = 0.0
            for k, (xa, xb, xc) in enumerate(zip(*(np.array_split(x, nband * 3) for x in (fa_, fb_, fc_)))):
                norm = max(np.sqrt(np.mean(xa ** 2)), 1e-30)
                d_gpu = np.sqrt(np.mean((xa - xb) ** 2)) / norm
                d_cpu = np.sqrt(np.mean((xa - xc) ** 2)) / norm
                worst = max(worst, d_gpu)
                # noise-dominated bands (E_z, j of a thermal plasma) decorrelate between ANY two runs once the reflux
                # RNG streams have shifted — the CPU pair shows how much; laser-dominated bands stay tight
                assert d_gpu <= 4 * d_cpu + 1e-4, (rel, k, d_gpu, d_cpu)
                print(f"  {rel} band piece {k}: GPU-CPU {d_gpu:.2e}  CPU(8 threads)-CPU(4 threads) {d_cpu:.2e}")
            print(f"{rel}: largest rms distance GPU path vs CPU {worst:.2e} (relative to the band's rms)")
        # strict mode: the same run must refuse to use the CPU implementation
        d = tempfile.mkdtemp(prefix="lpi_strict_")
        runs["strict"] = (d, "")
        env = dict(os.environ, LD_PRELOAD=LIB, VPIC_B200_STRICT="1", VPIC_LPI_STEPS="3", VPIC_LPI_NPPC="8")
        r = subprocess.run([path, "--tpp", "1"], cwd=d, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode != 0 and "VPIC_B200_STRICT=1 forbids" in (r.stdout + r.stderr)
    finally:
        for d, _ in runs.values():
            shutil.rmtree(d, ignore_errors=True)


SHIMRUN = os.path.join(ROOT, "oracle", "mpi_shim", "shimrun")


def _run_ranks(n, path, args, preload, cwd, extra_env=None, timeout=1800):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = LIB
        env.update(extra_env or {})
    env["SHIMRUN_ALL"] = "0"
    r = subprocess.run([SHIMRUN, "-n", str(n), path] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    logs = "".join(open(os.path.join(cwd, f)).read() for f in sorted(os.listdir(cwd)) if f.startswith("shimrun."))
    return r.returncode, r.stdout + r.stderr, logs


def test_pcomm_deck_passes_on_eight_ranks_with_the_device_boundary_p():
    """The reference's own migration test (pcomm.deck, 8 MPI ranks, 2 x 2 x 2) with every rank's hot path on the GPU:
    boundary_p packs and injects on the device, the records travel through the host program's mp_* ports.  All eight
    ranks share the one GPU of the test box.  The deck checks the arrivals itself: exact voxel, offsets within 11 ulp."""
    path = _need("pcomm.scalar")
    with tempfile.TemporaryDirectory() as d:
        rc, out, logs = _run_ranks(8, path, ["1", "1"], True, d, extra_env={"VPIC_B200_TRACE": "1", "VPIC_B200_STRICT": "1"})
    assert rc == 0 and "pass" in out and "FAIL" not in out + logs, (out + logs)[-3000:]
    t = _trace(out)
    assert t["boundary_p_species_on_device"] > 0 and t["advance_p"] > 0


@pytest.mark.parametrize("mode", ["coherent", "auto"])
def test_harris_two_ranks_on_the_gpu_match_the_two_rank_reference(mode):
    """sample/harris on TWO MPI ranks (its own 1 x nproc x 1 slab topology) with the hot path of both ranks on the GPU,
    against the same binary on two CPU ranks: same decomposition, same per-rank seeds, so the energy histories agree to
    the fp32 deposit-order tolerance of the one-rank test.  VPIC_B200_STRICT=1: nothing may fall back to a CPU kernel —
    migration, current and tang-B halos, and the divergence-cleaning halos all run through the device kernels."""
    path = _need("harris.scalar")
    hist = {}
    for tag, preload in (("cpu", False), ("gpu", True)):
        d = tempfile.mkdtemp(prefix=f"harris2_{tag}_")
        try:
            rc, out, logs = _run_ranks(2, path, ["--tpp", "1"], preload, d,
                                       extra_env=dict(MODES[mode], VPIC_B200_TRACE="1", VPIC_B200_STRICT="1"))
            assert rc == 0 and "normal exit" in out, (out + logs)[-3000:]
            hist[tag] = np.loadtxt(os.path.join(d, "energies"), comments="%")
            if preload:
                t = _trace(out)
                assert t["boundary_p_species_on_device"] > 0 and t["field_kernel_fallback_to_reference"] == 0
                assert "is not served on the device" not in out + logs
        finally:
            shutil.rmtree(d, ignore_errors=True)
    a, b = hist["cpu"], hist["gpu"]
    assert a.shape == b.shape and a.shape[0] > 10
    tot_a, tot_b = a[:, 1:].sum(axis=1), b[:, 1:].sum(axis=1)
    assert np.abs(tot_a - tot_b).max() <= 1e-5 * np.abs(tot_a).max()
    for col in range(1, a.shape[1]):
        scale = np.abs(a[:, col]).max()
        if scale > 1e-3 * np.abs(tot_a).max():                      # the components that carry the energy
            # two ranks: the shared planes add one more reordering of fp32 sums per step than the one-rank run has
            assert np.abs(a[:, col] - b[:, col]).max() <= 2e-3 * scale, col
