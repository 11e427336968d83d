"""CPU tests of host-side logic that needs no GPU: the reference's post-processing scripts
run unchanged through scripts/run_reference_scripts.py, the threshold rule of
scripts/thresholds_md.py follows include/doGemm.hh:418-490, and the stream-K segment
arithmetic of csrc/dgemm_tma.cu covers every k-tile exactly once with consistent
contributor counts."""
import importlib.util
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------- stream-K arithmetic
def _streamk_plan(tiles, ktiles, P):
    """Mirror of dgemm_tma_dispatch's decision (csrc/dgemm_tma.cu)."""
    rem, waves = tiles % P, -(-tiles // P)
    seg = rem * ktiles // P
    if rem > 0 and rem * 10 <= P * 9 and seg >= 8 and waves <= 8:
        return rem, ktiles // seg + 2
    return 0, 0


@pytest.mark.parametrize("tiles,ktiles", [(256, 64), (256, 16), (182, 35), (255, 129), (100, 64), (149, 40),
                                          (1183, 9), (300, 1000), (75, 8), (64, 32), (50, 40), (36, 64)])
def test_streamk_segments_cover_every_ktile_once(tiles, ktiles):
    P = 148
    sk_tiles, sk_slots = _streamk_plan(tiles, ktiles, P)
    if sk_tiles == 0:
        pytest.skip("shape does not take the stream-K schedule")
    units = sk_tiles * ktiles
    begin = lambda b: b * units // P                       # sk_begin
    cta_of = lambda u: ((u + 1) * P - 1) // units          # sk_cta_of
    seen = [0] * units
    contrib = {}
    for b in range(P):
        u, uend = begin(b), begin(b + 1)
        assert uend > u, "empty segment: contributor ranges would be wrong"
        while u < uend:
            t, kt0 = divmod(u, ktiles)
            kt1 = min(ktiles, kt0 + (uend - u))
            for x in range(t * ktiles + kt0, t * ktiles + kt1):
                seen[x] += 1
            if not (kt0 == 0 and kt1 == ktiles):
                first = cta_of(t * ktiles)
                last = cta_of((t + 1) * ktiles - 1)
                assert first <= b <= last
                contrib.setdefault(t, set()).add(b - first)
                assert last - first + 1 <= sk_slots
            u += kt1 - kt0
    assert all(c == 1 for c in seen)
    for t, ords in contrib.items():
        n = cta_of((t + 1) * ktiles - 1) - cta_of(t * ktiles) + 1
        assert ords == set(range(n)), (t, ords, n)  # the counter reaches ncontrib exactly


def test_cta_of_is_inverse_of_begin():
    P, units = 148, 6912
    begin = lambda b: b * units // P
    cta_of = lambda u: ((u + 1) * P - 1) // units
    for b in range(P):
        for u in (begin(b), begin(b + 1) - 1):
            assert cta_of(u) == b


# ---------------------------------------------------------------- threshold rule
def test_threshold_rule_matches_reference_semantics():
    thr = _load(os.path.join(ROOT, "scripts", "thresholds_md.py"), "thresholds_md").threshold
    order = [(d, d, d) for d in range(1, 9)]
    cpu = dict(zip(order, [5, 5, 5, 5, 5, 5, 5, 5]))
    # GPU wins from size 3 on
    assert thr(order, cpu, dict(zip(order, [1, 2, 6, 7, 8, 9, 9, 9]))) == (3, 3, 3)
    # one CPU win in between does not reset (the previous GPU result still beats the CPU) ...
    assert thr(order, cpu, dict(zip(order, [1, 6, 4, 7, 8, 9, 9, 9]))) == (2, 2, 2)
    # ... two in a row do (doGemm.hh:427-429), and the threshold is found again later
    assert thr(order, cpu, dict(zip(order, [1, 6, 4, 4, 8, 9, 9, 9]))) == (5, 5, 5)
    assert thr(order, cpu, dict(zip(order, [1, 2, 3, 4, 4, 4, 4, 4]))) is None


# ---------------------------------------------------------------- reference scripts, unchanged
@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container)")
def test_reference_postprocessing_scripts_run_unchanged(tmp_path):
    blob_cpu = os.path.join(ROOT, "oracle", "_ref", "gpu-blob-cpu")
    if not os.path.exists(blob_cpu):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    run = tmp_path / "run"
    run.mkdir()
    # the harness resolves libconsume.so through a relative rpath (reference Makefile:240)
    os.makedirs(run / "src" / "Consume")
    os.symlink(os.path.join(ROOT, "oracle", "_ref", "libconsume.so"), run / "src" / "Consume" / "libconsume.so")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    subprocess.check_call([blob_cpu, "-i", "2", "-s", "1", "-d", "32", "--no_gpu", "-o", str(run / "csv")],
                          cwd=run, env=env, stdout=subprocess.DEVNULL)
    csvs = sorted(os.listdir(run / "csv"))
    assert len(csvs) == 28
    # give every file GPU rows too (a copy of the CPU rows under the three GPU device names), so
    # both scripts go down their GPU paths
    for f in csvs:
        p = run / "csv" / f
        lines = p.read_text().splitlines()
        rows = []
        for ln in lines[1:]:  # harness order: cpu, once, always, unified per problem size
            rows.append(ln)
            rows += [ln.replace("cpu,", dev + ",", 1) for dev in ("gpu_offloadOnce", "gpu_offloadAlways", "gpu_unified")]
        lines, extra = lines[:1], rows
        p.write_text("\n".join(lines + extra) + "\n")
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_scripts.py"), str(run / "csv"),
                        "--out", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    graphs = [d for d in os.listdir(out) if d.startswith("Graphs_")]
    assert len(graphs) == 1
    pngs = os.listdir(out / graphs[0])
    assert len(pngs) == 28 and all(p.endswith(".png") for p in pngs)
    with open(out / graphs[0] / pngs[0], "rb") as fh:
        assert fh.read(8) == b"\x89PNG\r\n\x1a\n"
    table = (out / "offload_thresholds.txt").read_text()
    assert table.count("rc=0") == 28 and "GPU (Offload Once)" in table
