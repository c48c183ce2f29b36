"""CPU check of the kernel header's compile-time tables (window masks, target bit, hash multipliers) against a
from-scratch restatement of the reference's lattice geometry (mod.rs:80-82 directions, mod.rs:186-232 back / mid /
front neighbourhoods).  The tool is host code compiled by nvcc; nothing runs on a GPU."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1), (-1, -1), (1, 1)]          # mod.rs:80-82, (row, col) steps


@pytest.fixture(scope="module")
def tables(tmp_path_factory):
    from evosops_b200 import _build
    exe = str(tmp_path_factory.mktemp("tables") / "dump_tables")
    subprocess.check_call([_build.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "tools", "dump_tables.cu")])
    return json.loads(subprocess.check_output([exe], text=True))


def nbrs(cell):
    return {(cell[0] + dr, cell[1] + dc) for dr, dc in DIRS}


def test_window_masks_follow_the_lattice_geometry(tables):
    ws = tables["ws"]
    assert ws >= 5 and 5 * ws <= 32                           # five rows of five cells fit a 32-bit window

    def bit(cell):                                            # window centred on the source (0, 0): bit = ws*(r+2) + (c+2)
        assert -2 <= cell[0] <= 2 and -2 <= cell[1] <= 2
        return 1 << (ws * (cell[0] + 2) + (cell[1] + 2))

    for d, (dr, dc) in enumerate(DIRS):
        t = tables["dirs"][d]
        s, tgt = (0, 0), (dr, dc)
        mid = nbrs(s) & nbrs(tgt)                             # common neighbours
        back = nbrs(s) - mid - {tgt}
        front = nbrs(tgt) - mid - {s}
        assert (len(back), len(mid), len(front)) == (3, 2, 3)  # phenotype [4][3][4] (mod.rs:25-36)
        assert (t["dr"], t["dc"]) == (dr, dc)
        assert t["delta"] == dr + 256 * dc                    # packed (row | col << 8) step
        assert 1 << t["tbit"] == bit(tgt)
        assert t["back"] == sum(bit(c) for c in back)
        assert t["mid"] == sum(bit(c) for c in mid)
        assert t["front"] == sum(bit(c) for c in front)
        assert t["all"] == t["back"] | t["mid"] | t["front"]


def test_window_hash_is_a_bijection_per_direction(tables):
    for t in tables["dirs"]:
        cells = [1 << k for k in range(32) if (t["all"] >> k) & 1]
        assert len(cells) == 8
        seen = set()
        for pattern in range(256):
            x = sum(c for k, c in enumerate(cells) if (pattern >> k) & 1)
            seen.add(((x * t["magic"]) & 0xFFFFFFFF) >> 24)
        assert seen == set(range(256))


def test_window_patch_flips_exactly_the_cells_inside_the_window(tables):
    # a committed move flips its source and target cell; a lane's 5x5 window must change in exactly those of the two
    # cells that lie inside it — for every relative position and direction
    ws = tables["ws"]
    assert len(tables["patch"]) == 21 * 21 * 6               # offsets -9..9 plus far ones (borrows of the packed subtraction)
    for orow, ocol, d, mask in tables["patch"]:
        want = 0
        for cell in ((orow, ocol), (orow + DIRS[d][0], ocol + DIRS[d][1])):
            if -2 <= cell[0] <= 2 and -2 <= cell[1] <= 2:
                want ^= 1 << (ws * (cell[0] + 2) + (cell[1] + 2))
        assert mask == want, (orow, ocol, d, mask, want)


def test_patch_lookup_table_and_index_arithmetic(tables):
    # the parallel commit rounds fetch "what lane f's move flips in lane l's window" from a table by one packed
    # subtraction, one validity mask and one byte dot product: for every relative position (incl. far ones, where the packed
    # subtraction borrows) and direction the result must be exactly the two flipped cells clipped to the 5x5 window, bit 31
    # iff the mover's source is lane l's source (same particle), bit 30 iff its target is, and nothing for a lane that is
    # not earlier in the batch
    ws = tables["ws"]
    assert len(tables["plut"]) == 21 * 21 * 6 * 3
    for orow, ocol, d, lane_f, lane_l, word in tables["plut"]:
        want = 0
        if lane_f < lane_l:
            src, tgt = (orow, ocol), (orow + DIRS[d][0], ocol + DIRS[d][1])
            for cell in (src, tgt):
                if -2 <= cell[0] <= 2 and -2 <= cell[1] <= 2:
                    want ^= 1 << (ws * (cell[0] + 2) + (cell[1] + 2))
            if src == (0, 0):
                want |= 1 << 31
            if tgt == (0, 0):
                want |= 1 << 30
        assert word == want, (orow, ocol, d, lane_f, lane_l, hex(word), hex(want))


def test_adaptive_rounds_controller(tables):
    # The adaptive kernels pick a chain's round form from a running mean of its commits per 32-step batch (in 1/64, decay
    # 1/8).  With a constant commit count c the mean settles at ~64 c; with none it must fall below every "calm" threshold
    # (integer decay stops at 7/64: a threshold at or below that would never be crossed — a bug this test pins); the forms
    # are ordered side by side < serial single batches < parallel rounds in commit rate; aggregation has no serial
    # single-batch phase, separation no side-by-side phase; and a chain that leaves the side-by-side phase (hot1 commits
    # in one iteration of wide_n batches, or hot2 in two) commits more per batch than the rate that brought it there.
    WIDE, ONE_SERIAL, ONE_PARALLEL = 0, 1, 2
    AGG, SEP = 0, 1
    for a in tables["adapt"]:
        steady = {c: (ema, form) for c, ema, form in a["steady"]}
        for c, (ema, _) in steady.items():
            assert abs(ema - 64 * c) <= 7, (a["beh"], c, ema)
        assert steady[0][0] <= 7
        if a["beh"] == SEP:
            assert a["calm"] == 0 and all(form != WIDE for _, form in steady.values())
            assert steady[0][1] == ONE_SERIAL
        else:
            assert a["calm"] > 7 and steady[0][1] == WIDE
            # leaving the side-by-side phase needs a higher commit rate than entering it
            assert a["hot1"] / a["wide_n"] > a["calm"] / 64.0 and a["hot2"] / (2 * a["wide_n"]) > a["calm"] / 64.0
        if a["beh"] == AGG:
            assert a["mid"] == 0 and all(form != ONE_SERIAL for _, form in steady.values())
        else:
            assert a["mid"] > a["calm"]
        forms = [steady[c][1] for c in sorted(steady)]
        assert forms == sorted(forms) and forms[-1] == ONE_PARALLEL      # monotone in the commit rate, hot chains: parallel rounds
