"""CPU tests: the C-ABI library loads and exports every symbol include/dlsa_b200.h declares (no compute calls),
host-side layout logic against the reference goldens, and the no-CPU-fallback rule."""
import os
import re
import subprocess

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import ROOT, load_golden

HEADER = os.path.join(ROOT, "include", "dlsa_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dlsa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from dlsa_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 17
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT\s+(dlsa_[a-z0-9_]+)", out))
    assert set(syms) <= exported, set(syms) - exported
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)  # binding mirrors the header 1:1
    assert _lib.lib.dlsa_abi_version() == 1
    assert _lib.launch_count() >= 0
    # size queries are pure host functions
    assert _lib.lib.dlsa_irls_pass_workspace_bytes(1024, 50, 97657) == 0  # one CTA per partition: no partials
    assert _lib.lib.dlsa_irls_pass_workspace_bytes(16, 20, 6250) > 0
    assert _lib.lib.dlsa_lars_path_workspace_bytes(50, 400) >= 50 * 50 * 8


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ["dlsa/models.py:110-113", "dlsa/models.py:114,130", "dlsa/R/wlse_func.R:55-62", "dlsa/dlsa.py:30-34",
                 "dlsa/dlsa.py:48", "dlsa/lsa.py:90-212", "dlsa/models.py:151-225", "dlsa/models.py:6-40"]:
        assert cite in src, cite


def test_argument_errors_do_not_need_a_gpu():
    from dlsa_b200 import _lib
    rc = _lib.lib.dlsa_irls_pass_f64(None, 4, None, None, 1, 4, 0, 0, None, None, None, None, None, 10, None, 0, None)
    assert rc == -1 and b"null pointer" in _lib.lib.dlsa_last_error()
    rc = _lib.lib.dlsa_lars_path_f64(None, None, 0, 0, 1.0, 0, 1e-16, 0, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(_lib.DlsaLibraryError):
        _lib.check(rc, "lars")


def test_no_cpu_fallback():
    from dlsa_b200 import engine
    X = torch.zeros(8, 3, dtype=torch.float64)
    with pytest.raises(ValueError, match="no CPU path"):
        engine.fit_logistic_partitions(X, torch.zeros(8, dtype=torch.float64), torch.tensor([0, 8]))
    # the product never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "dlsa_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_dummy_bookkeeping_matches_reference_golden():
    from dlsa_b200 import dummies
    g = load_golden("dummy_layout.npz")
    df = pd.DataFrame({"Month": g["in_Month"], "Carrier": g["in_Carrier"]})
    counts = dummies.dummy_factors_counts(df, ["Month", "Carrier"])
    assert [str(k) for k in counts["Month"].keys()] == list(g["ref_count_keys_Month"])
    assert list(counts["Carrier"].keys()) == list(g["ref_count_keys_Carrier"])
    info = dummies.select_dummy_factors(counts, list(g["in_keep_top"]), "000_OTHERS")
    assert info["factor_selected"]["Carrier"] == list(g["ref_selected_Carrier"])
    assert info["factor_dropped"]["Carrier"] == list(g["ref_dropped_Carrier"])
    assert info["factor_selected_names"]["Carrier"] == list(g["ref_selected_names_Carrier"])
    assert info["factor_selected_names"]["Month"] == list(g["ref_selected_names_Month"])
    cols = dummies.design_columns(["partition_id", "y", "znum", "anum", "Month", "Carrier"], "y", info,
                                  list(g["in_baselines"]), True)
    assert ["par_id", "coef", "Sig_invMcoef"] + cols == list(g["ref_frame_columns"])
    schema = dummies.convert_schema(["znum", "anum", "Month", "Carrier"], info, True, list(g["in_baselines"]))
    assert [n for n, _ in schema] == cols and all(t == "double" for _, t in schema)
    merged = dummies.cumsum_dicts(counts, counts)
    assert merged["Carrier"]["A"] == 2 * counts["Carrier"]["A"] and dummies.cumsum_dicts({}, counts) is counts


def test_sdummies_level_rule():
    """dlsa/sdummies.py:41-83 (Spark absent -> parity unpinned): keep while cum share <= keep_top OR count >= 10%."""
    from dlsa_b200 import dummies
    counts = {"Carrier": {"A": 50, "B": 20, "C": 12, "D": 10, "E": 5, "F": 3}}
    info = dummies.select_sdummy_levels(counts, [0.8])
    assert info["factor_selected"]["Carrier"] == ["A", "B", "C", "D"]  # D kept by the 10% rule (cum = 0.92 > 0.8)
    assert sorted(info["factor_dropped"]["Carrier"]) == ["E", "F"]
    assert info["factor_selected_names"]["Carrier"] == ["A", "B", "C", "D", "000_OTHERS"]
    assert dummies.select_sdummy_levels(counts, [0.8], dropLast=True)["factor_selected_names"]["Carrier"][-1] == "D"


def test_synth_host_restatement_is_deterministic():
    from dlsa_b200 import synth
    from oracle import synth_host
    u = synth_host.host_uniform(7, np.arange(5, dtype=np.uint64)[:, None], np.arange(3, dtype=np.uint64)[None, :])
    assert u.shape == (5, 3) and np.all((u >= 0) & (u < 1))
    assert np.array_equal(u, synth_host.host_uniform(7, np.arange(5, dtype=np.uint64)[:, None], np.arange(3, dtype=np.uint64)[None, :]))
    big = synth_host.host_uniform(1, np.arange(200000, dtype=np.uint64), np.uint64(0))
    assert abs(big.mean() - 0.5) < 5e-3 and abs(big.var() - 1 / 12) < 2e-3
    off = synth.equal_partition_offsets(100_000_000, 1024)
    sizes = (off[1:] - off[:-1])
    assert int(off[-1]) == 100_000_000 and int(sizes.max()) == 97657 and int(sizes.min()) == 97656
    assert np.array_equal(synth.true_beta(50), np.r_[np.ones(20), np.zeros(30)])


def test_bench_contract_reference_arm_and_no_gpu_failure():
    """bench.py --impl reference runs the reference's CPU path (oracle/ref_cpu.py) and prints ONE JSON line with the
    contract's keys; without a GPU the product arm must fail loudly instead of falling back to anything."""
    import json
    import sys
    bench = os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, bench, "--impl", "reference", "--steps", "1", "--warmup", "0", "--rows", "32000",
                          "--parts", "16", "--cpu-parts", "4", "--p", "6"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "dlsa_rows_per_sec" and d["unit"] == "rows/s" and d["value"] > 0
    from oracle import ref_shim
    assert d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["config"]["workload"].startswith("c2: simulated logistic") and "sample" in d["config"]
    if not torch.cuda.is_available():
        out = subprocess.run([sys.executable, bench, "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                             timeout=300)
        assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_tiled_schedule_covers_every_tile_exactly_once():
    """The tiled fused pass gives every 8x8 tile of the lower triangle of H to exactly one (sibling CTA, warp, slot) —
    that is what makes it reduction-free and bitwise reproducible.  Checked on the CPU for every p the kernel accepts,
    together with the ring geometry (>= 3 stages, a multiple of the 3 link warps, within 227 KB of shared memory)."""
    import ctypes as C
    from dlsa_b200 import _lib
    sched = (C.c_uint32 * (16 * 12 * 3))()
    info = (C.c_int * 8)()
    assert _lib.lib.dlsa_irls_tiled_schedule(56, 0, sched, info) != 0       # p <= 56 belongs to the small kernel
    assert _lib.lib.dlsa_irls_tiled_schedule(513, 0, sched, info) != 0
    for px in list(range(57, 513)):
        for icpt in ((0, 1) if px % 37 == 0 or px in (57, 181, 511) else (0,)):
            p = px + icpt
            if p > 512:
                continue
            assert _lib.lib.dlsa_irls_tiled_schedule(px, icpt, sched, info) == 0, (px, icpt)
            NT, TG, CH, stages, pre, Sd, smem, slots = list(info)
            assert NT == (p + 7) // 8 and 1 <= TG <= 16 and CH in (16, 32)
            assert stages >= 3 and stages % 3 == 0 and smem <= 227 * 1024
            assert Sd % 16 == 4 and Sd >= px + 1 and Sd >= 32 * ((NT + 3) // 4)
            assert pre == (1 if TG >= 2 else 0)
            owner = np.zeros((NT, NT), dtype=np.int32)
            g_owner = np.zeros((NT + 3) // 4, dtype=np.int32)
            per_warp_partials = {}
            for idx in range(TG * 12 * 3):
                e = sched[idx]
                mask, m, J, g = e & 0xFF, (e >> 8) & 0xFF, (e >> 16) & 0xFF, (e >> 24) & 1
                if mask == 0:
                    assert e == 0
                    continue
                if mask != 0xFF:
                    per_warp_partials[idx // 3] = per_warp_partials.get(idx // 3, 0) + 1
                for ii in range(2):
                    for jj in range(4):
                        if mask & (1 << (4 * ii + jj)):
                            i, j = 2 * m + ii, 4 * J + jj
                            assert i < NT and j <= i, (p, i, j)
                            owner[i, j] += 1
                g_owner[J] += g
            assert all(sched[idx] == 0 for idx in range(TG * 12 * 3, 16 * 12 * 3))
            lower = np.tril(np.ones((NT, NT), dtype=np.int32))
            assert np.array_equal(owner, lower), p                 # every lower-triangle tile exactly once, nothing else
            assert np.all(g_owner == 1), p                        # every column group's gradient has one owner
            assert max(per_warp_partials.values(), default=0) <= 1  # the body handles one partial block per warp


def test_categorical_plan_gives_every_level_one_home():
    """K1c's host-side layout (csrc/irls_cat.cu::cat_plan through the host-only dlsa_irls_cat_plan): every level of every
    categorical lives either in exactly one DMMA tile slot or in its own 64-byte half row of a scatter table; no two of
    those overlap, sinks included; the whole CTA fits 227 KB; pair-histogram cells = sum of (L+1)(L'+1)."""
    import ctypes as C
    from dlsa_b200 import _lib
    rng = np.random.default_rng(4)
    shapes = [(7, [11, 6, 7, 75, 75]), (0, [5]), (2, [3, 4]), (3, [8] * 8), (7, [20] * 8), (6, [40, 60, 90]),
              (4, [16, 9, 16, 16]), (1, [1, 8, 2, 17]), (5, [1, 1, 30, 100]), (0, [192]), (7, [24] * 8)]
    while len(shapes) < 150:
        n_cat = int(rng.integers(1, 9))
        lev = rng.integers(1, 60, size=n_cat)
        while lev.sum() > 192:
            lev[np.argmax(lev)] //= 2
        shapes.append((int(rng.integers(0, 8)), [int(x) for x in lev]))
    info, cat, tile = (C.c_int * 12)(), (C.c_int * 32)(), (C.c_int * 144)()
    for n_num, lev in shapes:
        n_cat = len(lev)
        arr = (C.c_int * n_cat)(*lev)
        for K, rows in ((120, 1_000_000), (3, 500), (1024, 60_000)):
            assert _lib.lib.dlsa_irls_cat_plan(K, n_num, n_cat, arr, rows, info, cat, tile) == 0
            S, warps, share, n_wide, rounds, ntiles, tab_bytes, stage, front, smem, npairs, Lpad = list(info)
            assert 1 <= S <= 16 and 4 <= warps <= 16 and smem <= 227 * 1024 and Lpad >= sum(lev) and Lpad % 8 == 0
            assert rounds == (n_wide + 1) // 2 and 0 <= ntiles <= 6
            assert stage == 256 * n_num + 128 * n_cat + 256 and front >= 2 * stage + 2560 + 160 * n_cat
            sets = (warps + 1) // 2 if share else warps
            assert smem >= warps * front + sets * tab_bytes + 8 * (8 + Lpad)
            want_pairs = sum((lev[a] + 1) * (lev[b] + 1) for a in range(n_cat) for b in range(a + 1, n_cat))
            assert npairs == (want_pairs if n_cat > 1 else lev[0] + 1)
            cats = np.array(list(cat)[:4 * n_cat]).reshape(n_cat, 4)
            slots = np.array(list(tile)).reshape(48, 3)
            assert sorted(cats[:, 1]) == list(range(n_cat))                 # one scratch row per categorical
            wide = [c for c in range(n_cat) if cats[c, 0] >= 0]
            assert len(wide) == n_wide and sorted(cats[wide, 0]) == list(range(n_wide))
            assert all(cats[c, 1] == cats[c, 0] for c in wide)               # wide slots come first in the scratch
            used = np.zeros(max(tab_bytes, 1), dtype=np.int32)
            for c in wide:
                for code in range(lev[c]):
                    used[cats[c, 2] + 128 * code: cats[c, 2] + 128 * code + 64] += 1
                used[cats[c, 3]: cats[c, 3] + 64] += 1                       # its sink half row
                assert cats[c, 2] % 64 == 0 and cats[c, 3] % 64 == 0
                assert cats[c, 3] + 64 <= tab_bytes
            assert used.max(initial=0) <= 1
            # narrow categoricals: every level has exactly one tile slot, slots are packed densely (categoricals mixed)
            lo = np.concatenate([[0], np.cumsum(lev)])
            home = np.zeros(sum(lev), dtype=np.int32)
            n_used = 0
            for sl in range(48):
                row, code, level = slots[sl]
                if code < 0:
                    assert code == -2 and sl >= n_used
                    continue
                assert sl == n_used and sl < 8 * ntiles
                n_used += 1
                c = int(np.flatnonzero(cats[:, 1] == row)[0])
                assert cats[c, 0] < 0 and 0 <= code < lev[c] <= 16 and level == lo[c] + code
                home[level] += 1
            assert ntiles == (n_used + 7) // 8
            for c in range(n_cat):
                assert (home[lo[c]: lo[c + 1]] == (0 if cats[c, 0] >= 0 else 1)).all()
    bad = (C.c_int * 9)(*([3] * 9))
    assert _lib.lib.dlsa_irls_cat_plan(4, 8, 2, bad, 100, info, cat, tile) != 0    # 8 numeric columns
    assert _lib.lib.dlsa_irls_cat_plan(4, 2, 9, bad, 100, info, cat, tile) != 0    # 9 categorical columns


def test_tiled_slices_follow_the_cost_model():
    """Slices per partition of the tiled pass, read back from the workspace size (K * S partial H buffers, none for
    S = 1): short partitions in a launch that already fills the SMs are NOT sliced (the wave-fill rule of round 1 cut the
    2500-row partitions of C5 five ways); few long partitions are sliced to fill the machine."""
    from dlsa_b200 import _lib
    lib = _lib.lib

    def slices(K, p, rows):
        b = lib.dlsa_irls_pass_workspace_bytes(K, p, rows)
        per, pre = (p * p + p + 1) * 8, 3 * K * rows * 8
        for S in range(1, 33):
            base = 0 if S == 1 else K * S * per
            if b in (base, base + pre):
                return S
        raise AssertionError((K, p, rows, b))

    assert slices(1024, 500, 2500) == 1 and slices(8192, 500, 2500) == 1 and slices(1024, 100, 2500) == 1
    assert slices(15, 182, 1_000_000) >= 8          # 15 partitions x 3 sibling CTAs cannot fill 148 SMs
    assert slices(32, 100, 976_563) >= 4 and slices(1, 100, 1_000_000) == 32
    for K, p, rows in [(3, 64, 100), (7, 300, 40), (148, 200, 5000)]:
        assert 1 <= slices(K, p, rows) <= 32
